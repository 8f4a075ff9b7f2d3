#!/usr/bin/env python
"""Profiling driver: load a workload, run a few ring epochs.  Used plain and under ncu
(B200_PROFILING.md): python tools/prof_ring.py [workload | m,n,nnz,k] [epochs] [nnz]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import bench  # noqa: E402
import mfb200  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if "," in wl:  # custom shape "m,n,nnz,k" (e.g. the block one rank trains per sub-step when C3 is split over 8 GPUs)
    m, n, nnz, k = (int(x) for x in wl.split(","))
else:
    m, n, nnz, k, desc = bench.WORKLOADS[wl]
if len(sys.argv) > 3:
    nnz = int(sys.argv[3])
R = mfb200.gen_ratings(m, n, 0, nnz)
s = mfb200.Session(m, n, k, iters=epochs, lam_p=bench.LAMBDA, lam_q=bench.LAMBDA, eta=bench.ETA, mode=mfb200.MODE_RING)
s.load(R)
for e in range(epochs):
    ms, tr = s.epochs(1)
    print("epoch %d: %.3f ms  %.3e upd/s  tr_rmse %.5f" % (e, ms, nnz / ms * 1e3, tr[0]), flush=True)
print(s.report())
s.close()
