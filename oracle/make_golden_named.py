#!/usr/bin/env python
"""Held-out RMSE of the COMPILED REFERENCE (oracle/_ref/libmf_ref.so, mf::mf_train + mf::calc_rmse) at the
named configurations of BASELINE.json -> tests/golden/named_configs.json.

TEST INFRASTRUCTURE.  Run in the build container after `make -C oracle`; takes several minutes and ~10 GB.  The
reference does not exist on the GPU box, so these numbers travel as a small fixture; tests/test_gpu_named_configs.py
and bench.py's `rmse_parity` compare the engine with them (north_star: within 0.5 % after equal epochs).

Data = the generator of SURVEY.md 8d (seed 42): ratings [0, nnz) train, [nnz, nnz + min(nnz/10, 10M)) held out.
lambda_p2 = lambda_q2 = 0.05, eta = 0.1, nr_bins = 20.  nr_threads = 1 is the reference's only reproducible mode
(SURVEY.md F5); the 100M / 250M configurations use all cores here and are therefore one draw of a quantity that
spreads by about 1e-3 relative from run to run.

    python oracle/make_golden_named.py [c2 c3 c4 zipf]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "named_configs.json")
LAM, ETA = 0.05, 0.1
CASES = {
    # name: (m, n, nnz, k, [(epochs, threads), ...])
    "c2": (138000, 27000, 20_000_000, 128, [(20, 1), (5, 1)]),
    "c3": (480000, 17800, 100_000_000, 128, [(20, 8), (5, 8)]),
    "c4": (1_000_000, 625_000, 250_000_000, 128, [(20, 8), (4, 8)]),
}


def ref_rmse(T, P, Q, b):
    L = orc.ref()
    T = np.ascontiguousarray(T, orc.NODE)
    return float(L.ref_rmse(T.ctypes.data, len(T), P.ctypes.data, Q.ctypes.data, P.shape[0], Q.shape[0], P.shape[1], b))


def main():
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle"
    want = sys.argv[1:] or ["c2", "c3"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in want:
        m, n, nnz, k, runs = CASES[name]
        R = orc.gen_ratings(m, n, 0, nnz)
        T = orc.gen_ratings(m, n, nnz, min(nnz // 10, 10_000_000))
        entry = res.setdefault(name, {"m": m, "n": n, "nnz": nnz, "k": k, "lambda": LAM, "eta": ETA, "nr_bins": 20,
                                      "heldout": "ratings [nnz, nnz + %d) of the same generator" % len(T), "runs": {}})
        for epochs, threads in runs:
            t0 = time.time()
            P, Q, b = orc.ref_train(R, m, n, k, epochs, lam_p=LAM, lam_q=LAM, eta=ETA, threads=threads)
            dt = time.time() - t0
            rm = ref_rmse(T, P, Q, b)
            entry["runs"]["%d" % epochs] = {"epochs": epochs, "nr_threads": threads, "heldout_rmse": rm,
                                            "train_seconds_here": round(dt, 1)}
            print(name, epochs, threads, rm, "%.1fs" % dt, flush=True)
            json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
