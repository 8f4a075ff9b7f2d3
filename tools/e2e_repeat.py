#!/usr/bin/env python
"""Repeats the end-to-end training call of bench.py (mfb200_train from host buffers) a few times in one process, after a
session like bench.py's device-resident leg, and prints the phases of every call (mfb200_report): is the end-to-end time
reproducible?   python tools/e2e_repeat.py [workload] [epochs] [calls]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import bench  # noqa: E402
import mfb200  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 4
m, n, nnz, k, _ = bench.WORKLOADS[wl]
R = mfb200.gen_ratings(m, n, 0, nnz)
T = mfb200.gen_ratings(m, n, nnz, min(nnz // 10, 10_000_000))
s = mfb200.Session(m, n, k, iters=5, lam_p=bench.LAMBDA, lam_q=bench.LAMBDA, eta=bench.ETA, mode=mfb200.MODE_RING)
s.load(R)
s.epochs(5)
s.rmse(T)
s.close()
P = np.zeros((m, k), np.float32)
Q = np.zeros((n, k), np.float32)
for c in range(calls):
    t0 = time.perf_counter()
    _, _, b, rep = mfb200.train(R, m, n, k, K, out=(P, Q), lam_p=bench.LAMBDA, lam_q=bench.LAMBDA, eta=bench.ETA, mode=mfb200.MODE_RING)
    dt = time.perf_counter() - t0
    print("call %d: %.3f s  create %.1f prep %.1f epochs %.1f finish %.1f destroy %.1f total %.1f ms" % (
        c, dt, rep["create_ms"], rep["prep_ms"], rep["epochs_ms"], rep["finish_ms"], rep["destroy_ms"], rep["total_ms"]), flush=True)
    if c == 1:
        mfb200.rmse(T, P, Q, b)  # (bench.py evaluates the returned model between calls)
