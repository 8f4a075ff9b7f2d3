#!/usr/bin/env python
"""Held-out RMSE of the throughput mode after a few epochs (the fast part of the descent), for schedule variants given
through the environment:   [env ...] python tools/transient_check.py [workload] [epochs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import bench  # noqa: E402
import mfb200  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m, n, nnz, k, desc = bench.WORKLOADS[wl]
R = mfb200.gen_ratings(m, n, 0, nnz)
T = mfb200.gen_ratings(m, n, nnz, min(nnz // 10, 10_000_000))
s = mfb200.Session(m, n, k, iters=epochs, lam_p=bench.LAMBDA, lam_q=bench.LAMBDA, eta=bench.ETA, mode=mfb200.MODE_RING)
s.load(R)
ms, tr = s.epochs(epochs)
print("%s epochs=%d env=%s  heldout %.5f  tr %s  ms/epoch %.2f  ref %s" % (
    wl, epochs, {k_: v for k_, v in os.environ.items() if k_.startswith("MFB200_")}, s.rmse(T),
    " ".join("%.4f" % x for x in tr), ms / epochs, bench.golden_rmse(wl, epochs)), flush=True)
s.close()
