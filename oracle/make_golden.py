#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libmf_ref.so).

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference to have been compiled
by `make -C oracle`).  The reference does not exist on the GPU box, so its outputs travel as
these small fixtures.  Every array below comes out of the reference's own code
(mf::mf_train at nr_threads=1 -- the reproducible mode, SURVEY.md F5 --, mf::mf_predict,
mf::calc_rmse), never out of the oracle restatement.

    python oracle/make_golden.py
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# the triplets and pairs of the reference's only test program, mfTest/mfTest.cpp:7-27
MFTEST_TRIPLETS = np.array([0, 0, 5, 0, 2, 10, 0, 3, 2, 1, 0, 7, 1, 1, 3, 1, 3, 0, 2, 1, 2, 2, 3, 9], np.float32)
MFTEST_PAIRS = np.array([0, 0, 0, 2, 0, 3, 1, 0, 1, 1, 1, 3, 2, 1, 2, 3, 2, 2], np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_predict_pairs(P, Q, b, pairs):
    L = orc.ref()
    m, k = P.shape
    n = Q.shape[0]
    return np.array([L.ref_predict(P.ctypes.data, Q.ctypes.data, m, n, k, b, int(pairs[2 * i]), int(pairs[2 * i + 1]))
                     for i in range(len(pairs) // 2)], np.float32)


def ref_topk(P, Q, b, users, topk):
    """score = reference mf_predict for every item; order (score desc, id asc) -- SURVEY.md 8c."""
    L = orc.ref()
    m, k = P.shape
    n = Q.shape[0]
    idx = np.empty((len(users), topk), np.int32)
    sc = np.empty((len(users), topk), np.float32)
    for i, u in enumerate(users):
        s = np.array([L.ref_predict(P.ctypes.data, Q.ctypes.data, m, n, k, b, int(u), v) for v in range(n)],
                     np.float32)
        order = np.lexsort((np.arange(n), -s.astype(np.float64)))[:topk]
        idx[i] = order
        sc[i] = s[order]
    return idx, sc


def bpr_golden():
    """3e. the one-class BPR losses (mf::mf_train with fun = 10 / 11, nr_threads=1, after srand(seed)) with the printed
    table, and calc_mpr / calc_auc of the trained model in both orientations.  `python oracle/make_golden.py bpr` alone."""
    import loss_cases
    out = {}
    for name, fun, kw, shape, seed in loss_cases.BPR_CASES:
        m, n, _, k, it = shape
        R = loss_cases.bpr_ratings(shape)
        P, Q, b, rows = orc._stable(lambda: orc.ref_train_ex(R, m, n, k, it, fun=fun, want_table=True, rand_seed=seed, **kw),
                                    orc._same_model)
        out[name + "_P"], out[name + "_Q"], out[name + "_b"] = P, Q, np.float32(b)
        out[name + "_table"] = np.array(rows, np.float64)
        out[name + "_mpr_auc"] = np.array(orc.ref_mpr_auc(R, P, Q, b, False) + orc.ref_mpr_auc(R, P, Q, b, True), np.float64)
        print(name, "table last", rows[-1], "mpr/auc, transposed mpr/auc", out[name + "_mpr_auc"])
    np.savez_compressed(os.path.join(OUT, "bpr.npz"), **out)


def model_text_case():
    """Factors that exercise the "%g" corners: exponents both ways, negative zero, integers, an unseen (NaN) row."""
    rng = np.random.RandomState(2)
    m, n, k = 9, 6, 5
    P = (rng.standard_normal((m, k)) * rng.choice([1e-6, 1e-3, 1, 1e4, 1e12], size=(m, 1))).astype(np.float32)
    Q = rng.rand(n, k).astype(np.float32)
    P[1] = np.nan
    Q[4] = np.nan
    P[3, :4] = [0.0, -0.0, 7.0, 123456.5]
    P[4, :3] = [1e-30, 1e-5, 0.0001]
    return P, Q


def main():
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle"
    os.makedirs(OUT, exist_ok=True)

    # 1. mfTest known-answer test: k=8, 30 iters, eta=.1, lambda=.1, 1 thread.
    tri = MFTEST_TRIPLETS.reshape(-1, 3)
    R = np.empty(len(tri), orc.NODE)
    R["u"], R["v"], R["r"] = tri[:, 0].astype(np.int32), tri[:, 1].astype(np.int32), tri[:, 2]
    P, Q, b = orc.ref_train_stable(R, 3, 4, 8, 30, lam_p=0.1, lam_q=0.1, eta=0.1, threads=1)
    pred = ref_predict_pairs(P, Q, b, MFTEST_PAIRS)
    np.savez(os.path.join(OUT, "mftest_kat.npz"), triplets=MFTEST_TRIPLETS, pairs=MFTEST_PAIRS, P=P, Q=Q,
             b=np.float32(b), pred=pred)
    print("mftest: b=%r pred=%s" % (b, pred))

    # 2. small synthetic cases, full factors (generator: SURVEY.md 8d, seed 42).
    small = [("s_1000x500_k20", 1000, 500, 50000, 20, 5), ("s_300x700_k8", 300, 700, 40000, 8, 3),
             ("s_600x400_k128_nan", 600, 400, 900, 128, 4), ("s_64x48_k40", 64, 48, 3000, 40, 6)]
    for name, m, n, nnz, k, it in small:
        R = orc.gen_ratings(m, n, 0, nnz)
        T = orc.gen_ratings(m, n, nnz, max(nnz // 10, 1))
        P, Q, b = orc.ref_train_stable(R, m, n, k, it, threads=1)
        rm = orc.ref().ref_rmse(T.ctypes.data, len(T), P.ctypes.data, Q.ctypes.data, m, n, k, b)
        users = np.arange(0, m, max(m // 8, 1), dtype=np.int32)[:8]
        tk = min(10, n)
        tidx, tsc = ref_topk(P, Q, b, users, tk)
        pairs = np.stack([T["u"][:64], T["v"][:64]], 1).astype(np.float32).ravel()
        pairs = np.concatenate([pairs, np.array([m, 0, 0, n, -1, 0, m + 5, n + 5], np.float32)])  # out of range -> b
        pp = ref_predict_pairs(P, Q, b, pairs)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), m=m, n=n, nnz=nnz, k=k, iters=it, P=P, Q=Q,
                            b=np.float32(b), heldout_rmse=rm, topk_users=users, topk_idx=tidx, topk_score=tsc,
                            pairs=pairs, pair_pred=pp, R_sha=sha(R))
        print(name, "heldout rmse", rm, "nan rows", int(np.isnan(P[:, 0]).sum()), int(np.isnan(Q[:, 0]).sum()))

    # 3. config #1 (10k x 5k, 1M ratings, k=32, 20 epochs): hash + row subsample + RMSEs.
    m, n, nnz, k, it = 10000, 5000, 1000000, 32, 20
    R = orc.gen_ratings(m, n, 0, nnz)
    T = orc.gen_ratings(m, n, nnz, nnz // 10)
    P, Q, b = orc.ref_train_stable(R, m, n, k, it, threads=1)
    rm = orc.ref().ref_rmse(T.ctypes.data, len(T), P.ctypes.data, Q.ctypes.data, m, n, k, b)
    rtr = orc.ref().ref_rmse(R.ctypes.data, len(R), P.ctypes.data, Q.ctypes.data, m, n, k, b)
    np.savez_compressed(os.path.join(OUT, "c1_10kx5k_k32.npz"), m=m, n=n, nnz=nnz, k=k, iters=it,
                        P_sha=sha(P), Q_sha=sha(Q), P_rows=P[::41], Q_rows=Q[::41], row_step=41,
                        b=np.float32(b), heldout_rmse=rm, train_rmse=rtr, R_sha=sha(R))
    print("c1: heldout rmse %.6f train rmse %.6f" % (rm, rtr))

    # 3b. the other MFSolver losses, L1 regularisation and NMF (mf::mf_train, nr_threads=1): full factors, the
    # printed table (tr_<metric>, obj), and the matching error measure on held-out ratings (calc_mae / calc_gkl /
    # calc_logloss / calc_accuracy).
    import loss_cases
    out = {}
    for name, fun, kw, kind in loss_cases.CASES:
        for shape in loss_cases.SHAPES:
            m, n, nnz, k, it = shape
            R = loss_cases.ratings(m, n, 0, nnz, kind)
            T = loss_cases.ratings(m, n, nnz, nnz // 10, kind)
            P, Q, b, rows = orc._stable(lambda: orc.ref_train_ex(R, m, n, k, it, fun=fun, want_table=True, **kw), orc._same_model)
            key = loss_cases.key(name, shape)
            out[key + "_P"], out[key + "_Q"], out[key + "_b"] = P, Q, np.float32(b)
            out[key + "_table"] = np.array(rows, np.float64)
            out[key + "_metric"] = orc.ref_metric(loss_cases.METRIC_OF[fun], T, P, Q, b)
            print(key, "table last", rows[-1], "heldout metric", out[key + "_metric"])
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)

    # 3c. mf_cross_validation (nr_threads=1): the mean error over the folds
    cv = {}
    for name, m, n, nnz, k, it, folds, bins in loss_cases.CV_CASES:
        _, fun, kw, kind = loss_cases.cv_case(name)
        R = loss_cases.ratings(m, n, 0, nnz, kind)
        cv["%s_%dx%d_f%d" % (name, m, n, folds)] = orc.ref_cross_validation_stable(R, m, n, k, it, folds, fun=fun, bins=bins, **kw)
    print("cv", cv)
    np.savez(os.path.join(OUT, "cv.npz"), **cv)

    bpr_golden()

    # 3d. the text model format: a small model written by the reference's own mf_save_model (mf/mf.cpp:4184-4225)
    import mfb200
    refl = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmf_ref.so"))
    save = getattr(refl, mfb200.SYM_SAVE_MODEL)
    save.restype = C.c_int
    save.argtypes = [C.POINTER(mfb200.MfModel), C.c_char_p]
    Pm, Qm = model_text_case()
    mdl = mfb200.MfModel(0, Pm.shape[0], Qm.shape[0], Pm.shape[1], np.float32(3.14159274),
                         Pm.ctypes.data_as(C.POINTER(C.c_float)), Qm.ctypes.data_as(C.POINTER(C.c_float)))
    assert save(C.byref(mdl), os.path.join(OUT, "model_text_ref.txt").encode()) == 0

    # 4. library-behaviour KATs taken from libc / libstdc++ themselves.
    libc = C.CDLL("libc.so.6")
    libc.srand(0)
    glibc = np.array([libc.rand() for _ in range(2000)], np.int32)
    np.savez(os.path.join(OUT, "lib_kat.npz"), glibc_rand_seed0=glibc,
             shuffle10=np.array([4, 3, 7, 8, 0, 5, 2, 1, 6, 9], np.int32),
             minstd4=np.array([7.82590359e-06, 0.131537795, 0.75560534, 0.458650142], np.float32))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bpr":
        bpr_golden()
    else:
        main()
