#!/usr/bin/env python
"""Held-out RMSE of the run kernel with the ring hand-off against T-row locks on the shape of
tests/test_gpu_multi_device.py (40k x 9k, 4M ratings, k=64), one device and MFB200_GPUS=2, after 10 and 20 epochs,
relative to the oracle's sequential run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200  # noqa: E402
import orc  # noqa: E402

m, n, nnz, k = 40_000, 9_000, 4_000_000, 64
R = mfb200.gen_ratings(m, n, 0, nnz)
T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
for it in (10, 20):
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    want = orc.oracle_rmse(T, Po, Qo, bo)
    for gpus in ("1", "2"):
        for kern in ("run", "tlock", ""):
            os.environ["MFB200_GPUS"] = gpus
            if kern:
                os.environ["MFB200_KERNEL"] = kern
            else:
                os.environ.pop("MFB200_KERNEL", None)
            if int(gpus) > mfb200.device_count():
                continue
            P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING)
            got = mfb200.rmse(T, P, Q, b)
            print("epochs %2d gpus %s kernel %-7s rmse %.5f oracle %.5f rel %+.4f (kernel code %d, %d CTAs, %d warps)" % (
                it, gpus, kern or "default", got, want, got / want - 1, rep["kernel"], rep["grid_ctas"], rep["cta_warps"]), flush=True)
