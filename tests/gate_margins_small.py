#!/usr/bin/env python
"""Margins of the 3 % gates of tests/test_gpu_parity.py::test_ring_mode_rmse_parity_small (tiny shapes, throughput mode):
relative deviation of the held-out RMSE from the compiled reference's golden value, several runs per shape."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import mfb200  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name in ("s_1000x500_k20", "s_300x700_k8", "s_64x48_k40"):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, max(nnz // 10, 1))
    devs = []
    for _ in range(reps):
        P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING)
        devs.append(mfb200.rmse(T, P, Q, b) / float(g["heldout_rmse"]) - 1)
    print(name, " ".join("%+.2f%%" % (100 * d) for d in devs), flush=True)
