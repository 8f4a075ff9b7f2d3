"""The cell kernel (csrc/sgd_cell.cu, k_sgd_cell_epoch: the CTA owns a T band for a step and its groups share the cell's
ratings dynamically) and the warp kernel (csrc/sgd_warp.cu, k_sgd_warp_epoch: the warp owns a T sub-band and serves its
four groups from one stream).

Checks, all through the C-ABI:
  * every rating is processed exactly once per epoch: with a step size so small that no factor changes (eta * g flushes
    to zero) the per-epoch training RMSE is a sum over all ratings of the initial model's squared error -- it must equal
    the run kernel's sum (same model, any order; double accumulation of fp32 terms) far below the RMSE gates' resolution;
  * held-out RMSE against the oracle's sequential run (the reference's update order) on shapes that exercise runs that
    cross chunk borders, m < n (sides swapped), several passes, k not a multiple of 32, few CTAs, a hot item row."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import mfb200  # noqa: E402
import orc  # noqa: E402

pytestmark = pytest.mark.gpu
KERNEL_CODE = {"run": 2, "cell": 3, "warp": 4, "tlock": 5, "item": 6}


def _train(monkeypatch, kernel, R, m, n, k, it, **kw):
    monkeypatch.setenv("MFB200_KERNEL", kernel)
    s = mfb200.Session(m, n, k, iters=it, mode=mfb200.MODE_RING, **kw)
    s.load(R)
    _, tr = s.epochs(it)
    P, Q, b = s.finish()
    rep = s.report()
    s.close()
    return P, Q, b, [float(x) for x in tr], rep


@pytest.mark.parametrize("shape", [(3000, 2000, 400_000, 128), (10_000, 5_000, 1_000_000, 32), (700, 2600, 90_000, 40),
                                   (60_000, 2_225, 1_562_500, 128)])
@pytest.mark.parametrize("kernel,env", [("cell", {}), ("cell", {"MFB200_CELL_CHUNK": "8"}),
                                        ("cell", {"MFB200_CELL_CHUNK": "1", "MFB200_CELL_S1": "3"}),
                                        ("warp", {}), ("warp", {"MFB200_RING_S1": "2"}), ("warp", {"MFB200_RING_CTAS": "148"}),
                                        ("tlock", {}), ("tlock", {"MFB200_RING_S1": "2"}),
                                        ("item", {}), ("item", {"MFB200_RING_CTAS": "37"})])
def test_every_rating_exactly_once(monkeypatch, shape, kernel, env):
    m, n, nnz, k = shape
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    R = mfb200.gen_ratings(m, n, 0, nnz)
    _, _, _, tr_cell, rep_c = _train(monkeypatch, kernel, R, m, n, k, 3, eta=1e-30)
    _, _, _, tr_run, rep_r = _train(monkeypatch, "run", R, m, n, k, 3, eta=1e-30)
    assert rep_c["kernel"] == KERNEL_CODE[kernel] and rep_r["kernel"] == KERNEL_CODE["run"], (rep_c, rep_r)
    assert np.allclose(tr_cell, tr_run, rtol=1e-9, atol=0), (tr_cell, tr_run, rep_c)


@pytest.mark.parametrize("case", [
    dict(m=3000, n=2000, nnz=400_000, k=128, it=5, env={}),
    dict(m=10_000, n=5_000, nnz=1_000_000, k=32, it=8, env={}),
    dict(m=700, n=2600, nnz=90_000, k=40, it=5, env={}),                      # m < n: the users are the S side
    dict(m=4000, n=3000, nnz=300_000, k=128, it=4, env={"MFB200_RING_CTAS": "3"}),  # 1000 rows per CTA: several passes
    dict(m=400, n=300, nnz=120_000, k=16, it=12, env={"MFB200_CELL_CHUNK": "8"}),     # long runs (dense): tails past a chunk
    dict(m=2000, n=1500, nnz=200_000, k=64, it=4, env={"MFB200_RING_CTAS": "1"}),    # one CTA: no ring at all
])
@pytest.mark.parametrize("kernel", ["cell", "warp", "tlock", "item"])
def test_cell_kernel_rmse_vs_oracle(monkeypatch, case, kernel):
    m, n, nnz, k, it = case["m"], case["n"], case["nnz"], case["k"], case["it"]
    for key, val in case["env"].items():
        monkeypatch.setenv(key, val)
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    P, Q, b, tr, rep = _train(monkeypatch, kernel, R, m, n, k, it)
    assert rep["kernel"] == KERNEL_CODE[kernel], rep
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    got, want = mfb200.rmse(T, P, Q, b), orc.oracle_rmse(T, Po, Qo, bo)
    # (the dense 400 x 300 case trains 12 epochs on 300 ratings per user: the held-out error of such a small model moves
    # by +-1 % with the timing-dependent order of the locks, observed -0.7 % ... +3.7 % over runs of the four opt-in kernels, profiles/r2_gate_margins.txt -- 6 % there)
    # (the item kernel walks all ratings of an item back to back: it converges FASTER than the reference's order during the
    # first epochs -- measured 2.6 % below the oracle after 8 epochs at 10k x 5k -- so its gate is one-sided wider)
    lo = -0.05 if kernel == "item" else -(0.06 if m == 400 else 0.02)
    assert lo < got / want - 1 < (0.06 if m == 400 else 0.02), (got, want, rep)


@pytest.mark.parametrize("kernel", ["cell", "warp", "tlock", "item"])
def test_cell_kernel_hot_item_row(monkeypatch, kernel):
    """Zipf item popularity: one shared-memory row is wanted by every group all the time."""
    m, n, nnz, k, it = 30_000, 8_000, 2_000_000, 32, 12
    R = orc.gen_ratings_zipf(m, n, 0, nnz)
    T = orc.gen_ratings_zipf(m, n, nnz, 200_000)
    P, Q, b, tr, rep = _train(monkeypatch, kernel, R, m, n, k, it)
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    got, want = mfb200.rmse(T, P, Q, b), orc.oracle_rmse(T, Po, Qo, bo)
    assert abs(got / want - 1) < 0.02, (got, want, rep)
