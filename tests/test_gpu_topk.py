"""GPU parity tests of batched top-k scoring (BASELINE.json config #5) through the C-ABI mfb200_topk.
Oracle: score = mf_predict (mf/mf.cpp:4295-4314), order = score descending, item id ascending,
restated in oracle/mf_oracle.cpp (orc_topk).  Bar: index lists AND scores bit-exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import mfb200  # noqa: E402
import orc  # noqa: E402

pytestmark = pytest.mark.gpu


def factors(m, n, k, seed, nan_users=(), nan_items=(), scale=0.35):
    rng = np.random.RandomState(seed)
    P = (rng.rand(m, k).astype(np.float32) * scale + rng.randn(m, k).astype(np.float32) * 0.1)
    Q = (rng.rand(n, k).astype(np.float32) * scale + rng.randn(n, k).astype(np.float32) * 0.1)
    P[list(nan_users)] = np.nan  # rows never seen in training, mf/mf.cpp:996-999
    Q[list(nan_items)] = np.nan
    return P, Q


def check(P, Q, b, users, topk):
    idx, sc = mfb200.topk(P, Q, b, users, topk)
    idx_o, sc_o = orc.oracle_topk(P, Q, b, users, topk)
    assert np.array_equal(idx, idx_o), (np.argwhere(idx != idx_o)[:5], idx[idx != idx_o][:5], idx_o[idx != idx_o][:5])
    assert np.array_equal(sc.view(np.uint32), sc_o.view(np.uint32))


def test_topk_small_item_set_exact_path():
    """n <= 2048: every item re-scored exactly; k not a multiple of 4; NaN user, NaN items, out-of-range user."""
    P, Q = factors(300, 500, 37, 1, nan_users=[5], nan_items=[0, 17, 499])
    users = np.array([0, 5, 299, 7, 7, 300, -1, 123], np.int32)
    check(P, Q, 3.25, users, 10)
    check(P, Q, -1.5, users, 1)
    check(P[:, :8].copy(), Q[:3, :8].copy(), 0.5, users[:3], 5)  # fewer items than topk: padded with -1


@pytest.mark.parametrize("shape", [(1000, 6000, 128, 10), (700, 5000, 40, 5), (400, 40000, 128, 100), (257, 9000, 64, 16),
                                   (500, 5000, 36, 7), (300, 4000, 21, 12)])
def test_topk_tensor_core_path_bit_exact(shape):
    """n > 2048: bf16 tcgen05 GEMM bounds + exact re-score.  Includes NaN rows on both sides."""
    m, n, k, topk = shape
    P, Q = factors(m, n, k, 7, nan_users=[3], nan_items=[1, 2, n - 1, n // 2])
    rng = np.random.RandomState(3)
    users = np.concatenate([[0, 3, m - 1], rng.randint(0, m, 140)]).astype(np.int32)
    check(P, Q, 3.5, users, topk)


def test_topk_on_trained_factors_with_ties():
    """Factors from a real (exact-mode) training run, plus duplicated item rows so that scores tie exactly."""
    m, n, nnz, k = 600, 3000, 60000, 32
    R = mfb200.gen_ratings(m, n, 0, nnz)
    P, Q, b, _ = mfb200.train(R, m, n, k, 5, mode=mfb200.MODE_EXACT)
    Q[100:200] = Q[1100:1200]  # exact ties: the lower item id must win
    users = np.arange(0, m, 7, dtype=np.int32)
    check(P, Q, b, users, 20)
