#!/usr/bin/env python
"""Prints, for every loss case of tests/loss_cases.py, the held-out error measure of the oracle's sequential run and
of the GPU in exact / ring (locks) / ring (tickets) mode on the same data.  python tools/loss_parity.py [m n nnz k it]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import loss_cases  # noqa: E402
import mfb200  # noqa: E402
import orc  # noqa: E402

m, n, nnz, k, it = (int(x) for x in sys.argv[1:6]) if len(sys.argv) > 5 else (3000, 2000, 400000, 32, 8)
for name, fun, kw, kind in loss_cases.CASES:
    R = loss_cases.ratings(m, n, 0, nnz, kind)
    T = loss_cases.ratings(m, n, nnz, nnz // 10, kind)
    w = loss_cases.METRIC_OF[fun]
    Po, Qo, bo, tro, _ = orc.oracle_train_ex(R, m, n, k, it, fun=fun, **kw)
    want = orc.oracle_metric(w, T, Po, Qo, bo)
    row = ["%-16s oracle %.5f (tr %.5f)" % (name, want, tro[-1])]
    for mode, tag in ((mfb200.MODE_EXACT, "exact"), (mfb200.MODE_RING, "locks"), (mfb200.MODE_RING_REPRO, "tickets")):
        s = mfb200.Session(m, n, k, iters=it, mode=mode, fun=fun, lam_p1=kw.get("lam_p1", 0.0),
                           lam_q1=kw.get("lam_q1", 0.0), nmf=kw.get("nmf", False))
        s.load(R)
        _, tr = s.epochs(it)
        P, Q, b = s.finish()
        s.close()
        got = mfb200.metric(w, T, P, Q, b)
        row.append("%s %.5f (%+.2f%%, tr %.5f)" % (tag, got, 100 * (got / want - 1), tr[-1]))
    print("  ".join(row), flush=True)
