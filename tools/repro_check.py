import os, sys
import numpy as np
sys.path.insert(0, "question-recommendation-system_b200")
import mfb200
m, n, nnz, k = 10000, 5000, 1000000, 32
R = mfb200.gen_ratings(m, n, 0, nnz)
for it in (1, 2, 5):
    P1, Q1, b1, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING_REPRO)
    P2, Q2, b2, _ = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING_REPRO)
    dP = (P1.view(np.uint32) != P2.view(np.uint32)).sum(); dQ = (Q1.view(np.uint32) != Q2.view(np.uint32)).sum()
    print(os.environ.get("TAG",""), "iters", it, "ctas", rep["grid_ctas"], "warps", rep["cta_warps"], "diffP", dP, "diffQ", dQ, "rowsP", (P1.view(np.uint32) != P2.view(np.uint32)).any(1).sum(), "rowsQ", (Q1.view(np.uint32) != Q2.view(np.uint32)).any(1).sum(), flush=True)
