#!/usr/bin/env python
"""Held-out RMSE of the throughput modes against the oracle's sequential run on Zipf item popularity, by epoch count:
separates a transient (order-dependent early epochs) from an error.  python tools/zipf_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfb200  # noqa: E402
import orc  # noqa: E402

m, n, nnz, k = 60000, 17800, 5_000_000, 32
R = orc.gen_ratings_zipf(m, n, 0, nnz)
T = orc.gen_ratings_zipf(m, n, nnz, 500_000)
for it in (6, 12, 20):
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it, lam_p=0.05, lam_q=0.05, eta=0.1)
    want = orc.oracle_rmse(T, Po, Qo, bo)
    for mode, name in ((mfb200.MODE_RING, "locks"), (mfb200.MODE_RING_REPRO, "tickets")):
        for kern in ("run", "band", "cell"):
            if kern == "cell" and name == "tickets":
                continue
            os.environ["MFB200_KERNEL"] = kern
            P, Q, b, rep = mfb200.train(R, m, n, k, it, lam_p=0.05, lam_q=0.05, eta=0.1, mode=mode)
            got = mfb200.rmse(T, P, Q, b)
            print("epochs %2d %-7s %-4s rmse %.6f oracle %.6f rel %+.4f ctas %d kernel %d" %
                  (it, name, kern, got, want, got / want - 1, rep["grid_ctas"], rep["kernel"]), flush=True)
