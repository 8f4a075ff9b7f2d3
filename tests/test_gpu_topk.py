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


def test_topk_overflowing_candidate_lists_fall_back_to_the_exact_path():
    """More tied items at the cut than a candidate list holds (2048): an all-zero user row (after L1 / NMF) scores 0 on
    every item, and thousands of duplicate item rows tie for every user.  Round 1 failed the whole call; now those
    users are recomputed exactly (every item scored, full sort) and the rest of the batch is untouched."""
    m, n, k, topk = 300, 9000, 64, 10
    P, Q = factors(m, n, k, 11)
    P[7] = 0.0                      # every item scores exactly 0: ids 0..9 win
    Q[3000:6500] = Q[2999]          # 3501 identical item rows: a user who likes them overflows
    P[9] = Q[2999] * 4.0            # ... like this one
    users = np.array([0, 7, 9, 7, 150, 299], np.int32)
    check(P, Q, 3.5, users, topk)
    idx, _ = mfb200.topk(P, Q, 3.5, users, topk)
    assert list(idx[1]) == list(range(topk)) and list(idx[2]) == list(range(2999, 2999 + topk))


@pytest.mark.parametrize("shape", [(200, 5000, 200, 10), (150, 4000, 64, 300)])
def test_topk_shapes_outside_the_gemm_path_use_the_exact_path(shape):
    """k > 128 or topk > 128 with more than 2048 items: round 1 returned an error; the exact path answers."""
    m, n, k, topk = shape
    P, Q = factors(m, n, k, 5, nan_users=[3], nan_items=[1, n - 1])
    users = np.array([0, 3, m - 1, 17, 400], np.int32)
    check(P, Q, 2.5, users, topk)


def test_resident_model_handle_equals_the_one_shot_calls():
    """mfb200_model_*: factors uploaded once; predictions, RMSE and top-k lists bit-equal to the one-shot calls."""
    m, n, k = 900, 7000, 128
    P, Q = factors(m, n, k, 21, nan_users=[4], nan_items=[6])
    rng = np.random.RandomState(1)
    pairs = np.stack([rng.randint(-1, m + 1, 5000), rng.randint(-1, n + 1, 5000)], 1).astype(np.float32).ravel()
    T = mfb200.gen_ratings(m, n, 0, 20000)
    users = rng.randint(0, m, 300).astype(np.int32)
    M = mfb200.Model(P, Q, 3.5)
    for _ in range(2):  # the handle serves any number of calls
        assert np.array_equal(M.predict_pairs(pairs).view(np.uint32), mfb200.predict_pairs(P, Q, 3.5, pairs).view(np.uint32))
        assert abs(M.rmse(T) / mfb200.rmse(T, P, Q, 3.5) - 1) < 1e-12  # (block sums meet in any order)
        idx, sc = M.topk(users, 25)
        idx1, sc1 = mfb200.topk(P, Q, 3.5, users, 25)
        assert np.array_equal(idx, idx1) and np.array_equal(sc.view(np.uint32), sc1.view(np.uint32))
    assert mfb200.eval_last_ms() > 0
    M.close()


@pytest.mark.parametrize("level", [None, "0", "1", "2"], ids=["auto", "none", "items", "both+bias"])
@pytest.mark.parametrize("kind", ["cold_users", "random", "trained"])
def test_topk_centring_levels_bit_exact(monkeypatch, level, kind):
    """The bf16 GEMM may see centred factors (MFB200_TOPK_CENTRE: 0 none, 1 items centred on their mean row, 2 both sides
    centred + item bias + items sorted by bias + the highest-bias tiles looked at item by item; default: chosen from the
    factors).  Whatever it sees, the lists and scores must be the exact ones.  cold_users: every user is the mean user plus a
    small own part -- the case centring exists for (the best items of a user are then the items with the largest bias);
    NaN rows on both sides, exact ties between duplicated items, more items than one tile of the top-bias region."""
    if level is None:
        monkeypatch.delenv("MFB200_TOPK_CENTRE", raising=False)
    else:
        monkeypatch.setenv("MFB200_TOPK_CENTRE", level)
    rng = np.random.RandomState(11)
    m, n, k, topk = 900, 21000, 128, 100
    if kind == "cold_users":
        common_p = rng.rand(k).astype(np.float32) * 0.4
        common_q = rng.rand(k).astype(np.float32) * 0.4
        own = np.where(rng.rand(m, 1) < 0.7, 0.01, 0.15).astype(np.float32)  # most users hardly differ from the mean user
        P = common_p + own * rng.randn(m, k).astype(np.float32)
        Q = common_q + 0.08 * rng.randn(n, k).astype(np.float32)
        b = 3.2
    elif kind == "random":
        P, Q = factors(m, n, k, 5)
        b = 3.5
    else:
        m, n, k, topk = 2000, 5000, 64, 50
        R = mfb200.gen_ratings(m, n, 0, 400_000)
        P, Q, b, _ = mfb200.train(R, m, n, k, 6, mode=mfb200.MODE_RING)
    P[[4, 77]] = np.nan
    Q[[0, 9, n - 1, n // 3]] = np.nan
    Q[300:340] = Q[5300:5340] if n > 5340 else Q[1300:1340]  # exact ties: the lower item id must win
    users = np.concatenate([[0, 4, 77, m - 1], rng.randint(0, m, 120)]).astype(np.int32)
    check(P, Q, b, users, topk)
