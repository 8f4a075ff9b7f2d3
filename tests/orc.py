"""ctypes doorways for the tests: the oracle (oracle/libmf_oracle.so) and, when present,
the compiled reference (oracle/_ref/libref_shim.so -> libmf_ref.so).

TEST INFRASTRUCTURE ONLY: nothing under question-recommendation-system_b200/ may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

NODE = np.dtype([("u", np.int32), ("v", np.int32), ("r", np.float32)])  # mf/mf.h:36-41


class OrcParam(C.Structure):
    _fields_ = [("k", C.c_int), ("nr_bins", C.c_int), ("nr_iters", C.c_int), ("lambda_p2", C.c_float),
                ("lambda_q2", C.c_float), ("eta", C.c_float), ("rsqrt_mode", C.c_int)]


class OrcParamEx(C.Structure):
    _fields_ = [("base", OrcParam), ("fun", C.c_int), ("lambda_p1", C.c_float), ("lambda_q1", C.c_float),
                ("do_nmf", C.c_int), ("rand_seed", C.c_uint)]


# the reference's loss codes, mf/mf.h:25-33
P_L2_MFR, P_L1_MFR, P_KL_MFR, P_LR_MFC, P_L2_MFC, P_L1_MFC = 0, 1, 2, 5, 6, 7
P_ROW_BPR_MFOC, P_COL_BPR_MFOC = 10, 11  # one-class BPR, mf/mf.h:31-32


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


_oracle = None
_ref = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libmf_oracle.so"])
    if os.path.exists("/root/reference/mf/mf.cpp"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libmf_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_train.restype = C.c_int
        L.orc_train.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.POINTER(OrcParam), C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_train_ex.restype = C.c_int
        L.orc_train_ex.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.POINTER(OrcParamEx), C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cross_validation.restype = C.c_double
        L.orc_cross_validation.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.POINTER(OrcParamEx), C.c_int,
                                           C.c_void_p]
        L.orc_metric.restype = C.c_double
        L.orc_metric.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_float]
        L.orc_utility_train.restype = C.c_int
        L.orc_utility_train.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                        C.c_double, C.c_int, C.c_void_p]
        L.orc_read_triplet.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_predict.restype = C.c_float
        L.orc_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.orc_predict_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_void_p, C.c_int, C.c_void_p]
        L.orc_rmse.restype = C.c_double
        L.orc_rmse.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_float]
        L.orc_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                               C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_mpr_auc.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_int, C.c_void_p]
        L.orc_cos_similarity.restype = C.c_int
        L.orc_cos_similarity.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_gen_ratings.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p]
        L.orc_gen_ratings_zipf.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p]
        L.orc_kat_random_map.argtypes = [C.c_int, C.c_void_p]
        L.orc_kat_minstd.argtypes = [C.c_int, C.c_void_p]
        L.orc_kat_glibc_rand.argtypes = [C.c_uint, C.c_int, C.c_void_p]
        L.orc_kat_rsqrt.restype = C.c_float
        L.orc_kat_rsqrt.argtypes = [C.c_float, C.c_int]
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_shim.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_shim.so"))
        L.ref_train.restype = C.c_int
        L.ref_train.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_train_ex.restype = C.c_int
        L.ref_train_ex.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ref_cross_validation.restype = C.c_double
        L.ref_cross_validation.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                           C.c_int, C.c_int]
        L.ref_metric.restype = C.c_double
        L.ref_metric.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_float]
        L.ref_utility_predict.restype = C.c_void_p
        L.ref_utility_predict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_predict.restype = C.c_float
        L.ref_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.ref_rmse.restype = C.c_double
        L.ref_rmse.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_float]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_mpr_auc.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_int, C.c_void_p]
        L.ref_cos_similarity.restype = C.c_void_p
        L.ref_cos_similarity.argtypes = [C.c_int, C.c_void_p, C.c_int]
        _ref = L
    return _ref


# ---- convenience wrappers ---------------------------------------------------------------

def gen_ratings(m, n, first, count, seed=42):
    out = np.empty(count, dtype=NODE)
    oracle().orc_gen_ratings(seed, m, n, first, count, _fp(out))
    return out


def gen_ratings_zipf(m, n, first, count, seed=42):
    """Same planted model, item popularity ~ 1/rank (oracle/mf_oracle.cpp, orc_gen_ratings_zipf)."""
    out = np.empty(count, dtype=NODE)
    oracle().orc_gen_ratings_zipf(seed, m, n, first, count, _fp(out))
    return out


def oracle_train(R, m, n, k, iters, lam_p=0.05, lam_q=0.05, eta=0.1, bins=20, rsqrt_mode=0):
    """Returns (P[m,k], Q[n,k], b, tr_rmse[iters], obj[iters])."""
    R = np.ascontiguousarray(R, dtype=NODE)
    prm = OrcParam(k, bins, iters, lam_p, lam_q, eta, rsqrt_mode)
    P = np.empty((m, k), np.float32)
    Q = np.empty((n, k), np.float32)
    b = C.c_float()
    tr = np.zeros(iters, np.float64)
    ob = np.zeros(iters, np.float64)
    oracle().orc_train(_fp(R), len(R), m, n, C.byref(prm), _fp(P), _fp(Q), C.byref(b), _fp(tr), _fp(ob))
    return P, Q, b.value, tr, ob


def oracle_train_ex(R, m, n, k, iters, fun=0, lam_p1=0.0, lam_q1=0.0, lam_p2=0.05, lam_q2=0.05, eta=0.1, nmf=False,
                    bins=20, rsqrt_mode=0, rand_seed=1):
    """Any MFSolver loss and the two one-class BPR losses (fun 10, 11; rand_seed: the state srand() left the process-wide
    rand() in).  Returns (P, Q, b, tr_metric[iters], obj[iters])."""
    R = np.ascontiguousarray(R, dtype=NODE)
    prm = OrcParamEx(OrcParam(k, bins, iters, lam_p2, lam_q2, eta, rsqrt_mode), fun, lam_p1, lam_q1, int(nmf), rand_seed)
    P = np.empty((m, k), np.float32)
    Q = np.empty((n, k), np.float32)
    b = C.c_float()
    tr = np.zeros(iters, np.float64)
    ob = np.zeros(iters, np.float64)
    oracle().orc_train_ex(_fp(R), len(R), m, n, C.byref(prm), _fp(P), _fp(Q), C.byref(b), _fp(tr), _fp(ob))
    return P, Q, b.value, tr, ob


def ref_train_ex(R, m, n, k, iters, fun=0, lam_p1=0.0, lam_q1=0.0, lam_p2=0.05, lam_q2=0.05, eta=0.1, nmf=False,
                 bins=20, threads=1, want_table=False, rand_seed=None):
    """The compiled reference's mf_train with any loss.  Returns (P, Q, b[, table rows as (tr_metric, obj)]).
    rand_seed: srand() is called with it first (the BPR scheduler seeds its negative generators from rand())."""
    R = np.ascontiguousarray(R, dtype=NODE)
    if rand_seed is not None:
        ref().ref_srand(C.c_uint(rand_seed))
    P = np.empty((m, k), np.float32)
    Q = np.empty((n, k), np.float32)
    b = C.c_float()
    buf = C.create_string_buffer(1 << 16) if want_table else None
    rc = ref().ref_train_ex(_fp(R), len(R), m, n, k, bins, iters, threads, fun, lam_p1, lam_q1, lam_p2, lam_q2, eta,
                            int(nmf), _fp(P), _fp(Q), C.byref(b), buf, (1 << 16) if want_table else 0)
    assert rc == 0
    if want_table:
        rows = []
        for line in buf.value.decode().splitlines()[1:]:
            f = line.split()
            rows.append((float(f[1]), float(f[-1])))
        return P, Q, b.value, rows
    return P, Q, b.value


def oracle_cross_validation(R, m, n, k, iters, folds, fun=0, lam_p1=0.0, lam_q1=0.0, lam_p2=0.05, lam_q2=0.05, eta=0.1,
                            nmf=False, bins=20):
    """Returns (mean error, per-fold errors)."""
    R = np.ascontiguousarray(R, dtype=NODE)
    prm = OrcParamEx(OrcParam(k, bins, iters, lam_p2, lam_q2, eta, 0), fun, lam_p1, lam_q1, int(nmf))
    errs = np.zeros(folds, np.float64)
    avg = oracle().orc_cross_validation(_fp(R), len(R), m, n, C.byref(prm), folds, _fp(errs))
    return avg, errs


def ref_cross_validation(R, m, n, k, iters, folds, fun=0, lam_p1=0.0, lam_q1=0.0, lam_p2=0.05, lam_q2=0.05, eta=0.1,
                         nmf=False, bins=20, threads=1):
    R = np.ascontiguousarray(R, dtype=NODE)
    return ref().ref_cross_validation(_fp(R), len(R), m, n, k, bins, iters, threads, fun, lam_p1, lam_q1, lam_p2,
                                      lam_q2, eta, int(nmf), folds)


def oracle_metric(which, R, P, Q, b):
    R = np.ascontiguousarray(R, dtype=NODE)
    return oracle().orc_metric(which, _fp(R), len(R), _fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b)


def ref_metric(which, R, P, Q, b):
    R = np.ascontiguousarray(R, dtype=NODE)
    return ref().ref_metric(which, _fp(R), len(R), _fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b)


def ref_train(R, m, n, k, iters, lam_p=0.05, lam_q=0.05, eta=0.1, bins=20, threads=1, want_stamps=False):
    """The compiled reference's mf_train. Returns (P, Q, b[, stamps, total_s])."""
    R = np.ascontiguousarray(R, dtype=NODE)
    P = np.empty((m, k), np.float32)
    Q = np.empty((n, k), np.float32)
    b = C.c_float()
    stamps = np.zeros(iters + 8, np.float64)
    ns = C.c_int(0)
    tot = C.c_double(0)
    rc = ref().ref_train(_fp(R), len(R), m, n, k, bins, iters, threads, lam_p, lam_q, eta, _fp(P), _fp(Q),
                         C.byref(b), _fp(stamps) if want_stamps else None, len(stamps), C.byref(ns),
                         C.byref(tot))
    assert rc == 0
    if want_stamps:
        return P, Q, b.value, stamps[:ns.value].copy(), tot.value
    return P, Q, b.value


def oracle_rmse(R, P, Q, b):
    R = np.ascontiguousarray(R, dtype=NODE)
    m, k = P.shape
    return oracle().orc_rmse(_fp(R), len(R), _fp(P), _fp(Q), m, Q.shape[0], k, b)


def oracle_predict_pairs(P, Q, b, pairs):
    pairs = np.ascontiguousarray(pairs, np.float32)
    out = np.empty(len(pairs) // 2, np.float32)
    oracle().orc_predict_pairs(_fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b, _fp(pairs), len(out),
                               _fp(out))
    return out


def oracle_topk(P, Q, b, users, topk):
    users = np.ascontiguousarray(users, np.int32)
    idx = np.empty((len(users), topk), np.int32)
    sc = np.empty((len(users), topk), np.float32)
    oracle().orc_topk(_fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b, _fp(users), len(users), topk,
                      _fp(idx), _fp(sc))
    return idx, sc


# ---- the reference's own race ---------------------------------------------------------------------------------------
# fpsg_core ends with sched.resume(); sched.terminate() (mf/mf.cpp:2910-2915): between the two calls the solver thread
# may start one more block, whose updates then land in the returned model (on tiny inputs the same window
# can dead-lock).  It is rare, but a live comparison must not depend on it: these wrappers repeat a run until two runs
# agree bit for bit (an extra block never repeats identically, the clean result always does).
def _stable(run, same, tries=5):
    prev = run()
    for _ in range(tries):
        cur = run()
        if same(prev, cur):
            return cur
        prev = cur
    raise AssertionError("the compiled reference did not give the same result twice in %d runs" % (tries + 1))


def _same_model(a, b):
    return (np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and
            np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and a[2] == b[2])


def ref_train_stable(*args, **kw):
    return _stable(lambda: ref_train(*args, **kw), _same_model)


def ref_train_ex_stable(*args, **kw):
    return _stable(lambda: ref_train_ex(*args, **kw), _same_model)


def ref_cross_validation_stable(*args, **kw):
    return _stable(lambda: ref_cross_validation(*args, **kw), lambda a, b: a == b)


def q_triplets(Q):
    """Float (item, knowledge point, value) triplets naming EVERY cell of an integer Q matrix."""
    items, k = Q.shape
    ii, kk = np.meshgrid(np.arange(items), np.arange(k), indexing="ij")
    return np.stack([ii.ravel(), kk.ravel(), Q.ravel()], 1).astype(np.float32).ravel()


def oracle_cos_similarity(item_id, tri):
    """mf::cos_similarity restated (oracle/mf_oracle.cpp).  Returns (ids by falling cosine as floats, cosines by item)."""
    tri = np.ascontiguousarray(tri, np.float32)
    items = oracle().orc_cos_similarity(item_id, _fp(tri), len(tri) // 3, None, None)
    out = np.empty(items, np.float32)
    cos = np.empty(items, np.float32)
    oracle().orc_cos_similarity(item_id, _fp(tri), len(tri) // 3, _fp(out), _fp(cos))
    return out, cos


def ref_cos_similarity(item_id, tri, items):
    """The compiled reference's mf::cos_similarity (build container only)."""
    tri = np.ascontiguousarray(tri, np.float32)
    p = ref().ref_cos_similarity(item_id, _fp(tri), len(tri) // 3)
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(items,)).copy()
    ref().ref_free(p)
    return out


def oracle_mpr_auc(R, P, Q, b, transpose=False, prob_m=None, prob_n=None):
    """calc_mpr_auc restated.  Returns (mpr, auc)."""
    R = np.ascontiguousarray(R, dtype=NODE)
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    out = np.zeros(2, np.float64)
    oracle().orc_mpr_auc(_fp(R), len(R), prob_m or P.shape[0], prob_n or Q.shape[0], _fp(P), _fp(Q), P.shape[0], Q.shape[0],
                         P.shape[1], b, int(transpose), _fp(out))
    return float(out[0]), float(out[1])


def ref_mpr_auc(R, P, Q, b, transpose=False, prob_m=None, prob_n=None):
    """The compiled reference's calc_mpr / calc_auc (build container only)."""
    R = np.ascontiguousarray(R, dtype=NODE)
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    out = np.zeros(2, np.float64)
    ref().ref_mpr_auc(_fp(R), len(R), prob_m or P.shape[0], prob_n or Q.shape[0], _fp(P), _fp(Q), P.shape[0], Q.shape[0],
                      P.shape[1], b, int(transpose), _fp(out))
    return float(out[0]), float(out[1])


def unique_pairs(m, n, count, seed=3):
    """`count` distinct (u, v) pairs with rating 1 (one-class data)."""
    rng = np.random.RandomState(seed)
    flat = rng.choice(m * n, size=count, replace=False)
    R = np.zeros(count, NODE)
    R["u"], R["v"], R["r"] = flat // n, flat % n, 1.0
    return R
