#!/usr/bin/env python
"""Margins of the timing-dependent RMSE gates of tests/test_gpu_cell_kernel.py: trains every case with every opt-in kernel
several times and prints the relative deviation of the held-out RMSE from the oracle's sequential run (one line per run,
flushed).  python tests/gate_margins.py [reps of the dense case] [reps of the others]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))  # orc: the oracle is the checker here
import mfb200  # noqa: E402
import orc  # noqa: E402

reps_dense = int(sys.argv[1]) if len(sys.argv) > 1 else 5
reps_other = int(sys.argv[2]) if len(sys.argv) > 2 else 1
CASES = [
    dict(m=400, n=300, nnz=120_000, k=16, it=12, env={"MFB200_CELL_CHUNK": "8"}, reps=reps_dense),
    dict(m=3000, n=2000, nnz=400_000, k=128, it=5, env={}, reps=reps_other),
    dict(m=10_000, n=5_000, nnz=1_000_000, k=32, it=8, env={}, reps=reps_other),
    dict(m=700, n=2600, nnz=90_000, k=40, it=5, env={}, reps=reps_other),
    dict(m=4000, n=3000, nnz=300_000, k=128, it=4, env={"MFB200_RING_CTAS": "3"}, reps=reps_other),
    dict(m=2000, n=1500, nnz=200_000, k=64, it=4, env={"MFB200_RING_CTAS": "1"}, reps=reps_other),
]
for c in CASES:
    m, n, nnz, k, it = c["m"], c["n"], c["nnz"], c["k"], c["it"]
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    want = orc.oracle_rmse(T, Po, Qo, bo)
    for kernel in ("cell", "warp", "tlock", "item", "run"):
        devs = []
        for _ in range(c["reps"]):
            os.environ.update(c["env"])
            os.environ["MFB200_KERNEL"] = kernel
            s = mfb200.Session(m, n, k, iters=it, mode=mfb200.MODE_RING)
            s.load(R)
            s.epochs(it)
            P, Q, b = s.finish()
            s.close()
            for key in c["env"]:
                os.environ.pop(key, None)
            devs.append(mfb200.rmse(T, P, Q, b) / want - 1)
        print("%dx%d nnz=%d k=%d it=%d %-5s " % (m, n, nnz, k, it, kernel) + " ".join("%+.2f%%" % (100 * d) for d in devs), flush=True)
