"""ctypes binding of lib/libmf.so -- the host-side mirror of the reference interface, for Python
callers (tests/, bench.py).  It binds exactly the C-ABI of include/mfb200.h plus the reference's
own php_* entry points (php_mf/mfWarp.h:6-10) and, for drop-in checks, the mangled mf:: symbols.

There is no fallback: if lib/libmf.so is missing or a call fails, this module raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MFB200_LIB: load another build of the same library (kernel variants under tools/, never a fallback)
LIB_PATH = os.environ.get("MFB200_LIB") or os.path.join(HERE, "lib", "libmf.so")

NODE = np.dtype([("u", np.int32), ("v", np.int32), ("r", np.float32)])  # mf_node, mf/mf.h:36-41

MODE_AUTO, MODE_EXACT, MODE_RING, MODE_RING_REPRO = 0, 1, 2, 3
# loss codes of mf_parameter.fun, mf/mf.h:25-33
P_L2_MFR, P_L1_MFR, P_KL_MFR, P_LR_MFC, P_L2_MFC, P_L1_MFC = 0, 1, 2, 5, 6, 7
P_ROW_BPR_MFOC, P_COL_BPR_MFOC = 10, 11  # one-class BPR (mf/mf.h:31-32): exact mode, one device


class Param(C.Structure):  # mfb200_param
    _fields_ = [("k", C.c_int), ("nr_bins", C.c_int), ("nr_iters", C.c_int), ("lambda_p2", C.c_float),
                ("lambda_q2", C.c_float), ("eta", C.c_float), ("quiet", C.c_int), ("mode", C.c_int),
                ("device", C.c_int), ("fun", C.c_int), ("lambda_p1", C.c_float), ("lambda_q1", C.c_float),
                ("do_nmf", C.c_int)]


class Report(C.Structure):  # mfb200_report
    _fields_ = [("mode_used", C.c_int), ("k_aligned", C.c_int), ("grid_ctas", C.c_int), ("cta_warps", C.c_int),
                ("bands", C.c_int), ("subbands", C.c_int), ("launches", C.c_longlong), ("prep_ms", C.c_double),
                ("epochs_ms", C.c_double), ("finish_ms", C.c_double), ("total_ms", C.c_double),
                ("last_tr_rmse", C.c_double), ("create_ms", C.c_double), ("destroy_ms", C.c_double),
                ("kernel", C.c_int), ("gpus", C.c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MfParameter(C.Structure):  # mf::mf_parameter, mf/mf.h:51-66 (passed BY VALUE to mf_train)
    _fields_ = [("fun", C.c_int), ("k", C.c_int), ("nr_threads", C.c_int), ("nr_bins", C.c_int),
                ("nr_iters", C.c_int), ("lambda_p1", C.c_float), ("lambda_p2", C.c_float),
                ("lambda_q1", C.c_float), ("lambda_q2", C.c_float), ("eta", C.c_float), ("do_nmf", C.c_bool),
                ("quiet", C.c_bool), ("copy_data", C.c_bool)]


class MfProblem(C.Structure):  # mf::mf_problem, mf/mf.h:43-49
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("nnz", C.c_longlong), ("R", C.c_void_p)]


class MfModel(C.Structure):  # mf::mf_model, mf/mf.h:70-79
    _fields_ = [("fun", C.c_int), ("m", C.c_int), ("n", C.c_int), ("k", C.c_int), ("b", C.c_float),
                ("P", C.POINTER(C.c_float)), ("Q", C.POINTER(C.c_float))]


# Itanium-mangled names of the reference API (nm -D on the reference's libmf.so)
SYM_MF_TRAIN = "_ZN2mf8mf_trainEPKNS_10mf_problemENS_12mf_parameterE"
SYM_MF_DEFAULT_PARAM = "_ZN2mf20mf_get_default_paramEv"
SYM_MF_DESTROY = "_ZN2mf16mf_destroy_modelEPPNS_8mf_modelE"
SYM_MF_PREDICT = "_ZN2mf10mf_predictEPKNS_8mf_modelEii"
SYM_CALC_RMSE = "_ZN2mf9calc_rmseEPNS_10mf_problemEPNS_8mf_modelE"
SYM_SAVE_MODEL = "_ZN2mf13mf_save_modelEPKNS_8mf_modelEPKc"
SYM_LOAD_MODEL = "_ZN2mf13mf_load_modelEPKc"
SYM_UTILITY_TRAIN = "_ZN2mf13utility_trainEPfiddiidRi"
SYM_UTILITY_PREDICT = "_ZN2mf15utility_predictEPfiS0_i"

_lib = None
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


class MfError(RuntimeError):
    pass


def build():
    """Compile lib/libmf.so in-tree (nvcc, sm_100a).  Cross-compiles without a GPU."""
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MfError("%s is missing: run `make -C %s` (there is no fallback path)" % (LIB_PATH, HERE))
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, cd, ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong
    L.mfb200_device_count.restype = ci
    L.mfb200_last_error.restype = C.c_char_p
    L.mfb200_version.restype = C.c_char_p
    L.mfb200_default_param.restype = Param
    L.mfb200_train.restype = ci
    L.mfb200_train.argtypes = [vp, ll, ci, ci, C.POINTER(Param), vp, vp, vp, C.POINTER(Report)]
    L.mfb200_predict_pairs.restype = ci
    L.mfb200_predict_pairs.argtypes = [vp, vp, ci, ci, ci, cf, vp, ll, vp]
    L.mfb200_rmse.restype = ci
    L.mfb200_rmse.argtypes = [vp, ll, vp, vp, ci, ci, ci, cf, C.POINTER(cd)]
    L.mfb200_metric.restype = ci
    L.mfb200_metric.argtypes = [ci, vp, ll, vp, vp, ci, ci, ci, cf, C.POINTER(cd)]
    L.mfb200_cross_validation.restype = ci
    L.mfb200_cross_validation.argtypes = [vp, ll, ci, ci, C.POINTER(Param), ci, vp, C.POINTER(cd)]
    L.mfb200_topk.restype = ci
    L.mfb200_topk.argtypes = [vp, vp, ci, ci, ci, cf, vp, ci, ci, vp, vp]
    L.mfb200_topk_last_ms.restype = cd
    L.mfb200_gen_ratings.restype = None
    L.mfb200_gen_ratings.argtypes = [C.c_uint64, ci, ci, ll, ll, vp]
    L.mfb200_session_create.restype = vp
    L.mfb200_session_create.argtypes = [ci, ci, C.POINTER(Param)]
    L.mfb200_session_load.restype = ci
    L.mfb200_session_load.argtypes = [vp, vp, ll]
    L.mfb200_session_reset.restype = ci
    L.mfb200_session_reset.argtypes = [vp]
    L.mfb200_session_epochs.restype = ci
    L.mfb200_session_epochs.argtypes = [vp, ci, C.POINTER(cf), vp]
    L.mfb200_session_finish.restype = ci
    L.mfb200_session_finish.argtypes = [vp, vp, vp, vp]
    L.mfb200_session_rmse.restype = ci
    L.mfb200_session_rmse.argtypes = [vp, vp, ll, C.POINTER(cd)]
    L.mfb200_session_report.restype = ci
    L.mfb200_session_report.argtypes = [vp, C.POINTER(Report)]
    L.mfb200_session_stream.restype = vp
    L.mfb200_session_stream.argtypes = [vp]
    L.mfb200_session_destroy.restype = None
    L.mfb200_session_destroy.argtypes = [vp]
    L.mfb200_dist_unique_id.restype = ci
    L.mfb200_dist_unique_id.argtypes = [vp]
    L.mfb200_dist_session_create.restype = vp
    L.mfb200_dist_session_create.argtypes = [ci, ci, C.POINTER(Param), ci, ci, vp]
    L.mfb200_dist_rotation.restype = None
    L.mfb200_dist_rotation.argtypes = [ci, ci, ll, ci, vp]
    L.mfb200_model_upload.restype = vp
    L.mfb200_model_upload.argtypes = [vp, vp, ci, ci, ci, C.c_float]
    L.mfb200_model_free.restype = None
    L.mfb200_model_free.argtypes = [vp]
    L.mfb200_model_predict_pairs.restype = ci
    L.mfb200_model_predict_pairs.argtypes = [vp, vp, ll, vp]
    L.mfb200_model_metric.restype = ci
    L.mfb200_model_metric.argtypes = [vp, ci, vp, ll, vp]
    L.mfb200_model_topk.restype = ci
    L.mfb200_model_topk.argtypes = [vp, vp, ci, ci, vp, vp]
    L.mfb200_eval_last_ms.restype = C.c_double
    L.mfb200_mpr_auc.restype = ci
    L.mfb200_mpr_auc.argtypes = [vp, ll, ci, ci, vp, vp, ci, ci, ci, cf, ci, C.POINTER(cd), C.POINTER(cd)]
    L.mfb200_cos_similarity.restype = ci
    L.mfb200_cos_similarity.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp, vp, vp]
    L.mfb200_plan_band.restype = ci
    L.mfb200_plan_band.argtypes = [ci, ci, ll, ci, ci, ci, ci, ci, vp]
    L.php_utility_train.restype = C.POINTER(cf)
    L.php_utility_train.argtypes = [vp, ci, cd, cd, ci, ci, cd, C.POINTER(ci)]
    L.php_utility_predict.restype = C.POINTER(cf)
    L.php_utility_predict.argtypes = [vp, ci, vp, ci]
    L.php_mf_my_train.restype = ci
    L.php_mf_my_train.argtypes = [C.c_char_p, C.c_char_p]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise MfError("%s failed: %s" % (what, lib().mfb200_last_error().decode()))


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    return lib().mfb200_device_count()


def make_param(k, iters, lam_p=0.05, lam_q=0.05, eta=0.1, bins=20, quiet=True, mode=MODE_AUTO, device=-1, fun=P_L2_MFR,
               lam_p1=0.0, lam_q1=0.0, nmf=False):
    return Param(k, bins, iters, lam_p, lam_q, eta, 1 if quiet else 0, mode, device, fun, lam_p1, lam_q1,
                 1 if nmf else 0)


def gen_ratings(m, n, first, count, seed=42):
    out = np.empty(count, dtype=NODE)
    lib().mfb200_gen_ratings(seed, m, n, first, count, _fp(out))
    return out


def train(R, m, n, k, iters, out=None, **kw):
    """mfb200_train: host buffers in, host factors out.  Returns (P, Q, b, report dict).
    out=(P, Q): the caller's own (already touched) float32 arrays of shape (m, k) and (n, k)."""
    R = np.ascontiguousarray(R, dtype=NODE)
    prm = make_param(k, iters, **kw)
    if out is not None:
        P, Q = out
        assert P.dtype == np.float32 and Q.dtype == np.float32 and P.shape == (m, k) and Q.shape == (n, k)
        assert P.flags.c_contiguous and Q.flags.c_contiguous
    else:
        P = np.empty((m, k), np.float32)
        Q = np.empty((n, k), np.float32)
    b = C.c_float()
    rep = Report()
    _check(lib().mfb200_train(_fp(R), len(R), m, n, C.byref(prm), _fp(P), _fp(Q), C.byref(b), C.byref(rep)),
           "mfb200_train")
    return P, Q, b.value, rep.as_dict()


def predict_pairs(P, Q, b, pairs):
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    pairs = np.ascontiguousarray(pairs, np.float32)
    out = np.empty(len(pairs) // 2, np.float32)
    _check(lib().mfb200_predict_pairs(_fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b, _fp(pairs), len(out),
                                      _fp(out)), "mfb200_predict_pairs")
    return out


def rmse(R, P, Q, b):
    R = np.ascontiguousarray(R, dtype=NODE)
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    out = C.c_double()
    _check(lib().mfb200_rmse(_fp(R), len(R), _fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b, C.byref(out)),
           "mfb200_rmse")
    return out.value


def metric(which, R, P, Q, b):
    """calc_mae / calc_gkl / calc_logloss / calc_accuracy: `which` = P_L1_MFR / P_KL_MFR / P_LR_MFC / P_L2_MFC."""
    R = np.ascontiguousarray(R, dtype=NODE)
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    out = C.c_double()
    _check(lib().mfb200_metric(which, _fp(R), len(R), _fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b,
                               C.byref(out)), "mfb200_metric")
    return out.value


def mpr_auc(R, P, Q, b, transpose=False, prob_m=None, prob_n=None):
    """calc_mpr / calc_auc (mf/mf.cpp:4406-4536) of the positives in R.  Returns (mpr, auc)."""
    R = np.ascontiguousarray(R, dtype=NODE)
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    mpr, auc = C.c_double(), C.c_double()
    _check(lib().mfb200_mpr_auc(_fp(R), len(R), prob_m or P.shape[0], prob_n or Q.shape[0], _fp(P), _fp(Q), P.shape[0],
                                Q.shape[0], P.shape[1], b, int(transpose), C.byref(mpr), C.byref(auc)), "mfb200_mpr_auc")
    return mpr.value, auc.value


def srand(seed):
    """The C library's srand(): the BPR losses seed their negative generators from the process-wide rand(), exactly as
    the reference's scheduler does (mf/mf.cpp:103-110)."""
    C.CDLL(None).srand(C.c_uint(seed))


def cos_similarity(tri, item_ids=None):
    """mfb200_cos_similarity: cosines of Q-matrix rows for a batch of items (default: all) against all items.
    Returns (order [ids by falling cosine], cos_sorted, cos_by_item, ties)."""
    tri = np.ascontiguousarray(tri, np.float32)
    items, k = C.c_int(), C.c_int()
    L = lib()
    _check(L.mfb200_cos_similarity(_fp(tri), len(tri) // 3, None, 0, None, None, None, None, C.byref(items), C.byref(k)),
           "mfb200_cos_similarity")
    ids = None if item_ids is None else np.ascontiguousarray(item_ids, np.int32)
    n_ids = items.value if ids is None else len(ids)
    order = np.empty((n_ids, items.value), np.int32)
    cs = np.empty((n_ids, items.value), np.float32)
    ci = np.empty((n_ids, items.value), np.float32)
    ties = np.empty(n_ids, np.int32)
    _check(L.mfb200_cos_similarity(_fp(tri), len(tri) // 3, None if ids is None else _fp(ids), 0 if ids is None else n_ids,
                                   _fp(order), _fp(cs), _fp(ci), _fp(ties), None, None), "mfb200_cos_similarity")
    return order, cs, ci, ties


def cross_validation(R, m, n, k, iters, folds, **kw):
    """mfb200_cross_validation.  Returns (mean error, per-fold errors)."""
    R = np.ascontiguousarray(R, dtype=NODE)
    prm = make_param(k, iters, **kw)
    errs = np.zeros(folds, np.float64)
    mean = C.c_double()
    _check(lib().mfb200_cross_validation(_fp(R), len(R), m, n, C.byref(prm), folds, _fp(errs), C.byref(mean)),
           "mfb200_cross_validation")
    return mean.value, errs


def topk(P, Q, b, users, k_top):
    P = np.ascontiguousarray(P, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    users = np.ascontiguousarray(users, np.int32)
    idx = np.empty((len(users), k_top), np.int32)
    sc = np.empty((len(users), k_top), np.float32)
    _check(lib().mfb200_topk(_fp(P), _fp(Q), P.shape[0], Q.shape[0], P.shape[1], b, _fp(users), len(users), k_top,
                             _fp(idx), _fp(sc)), "mfb200_topk")
    return idx, sc


def dist_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, every rank needs the same bytes)."""
    buf = np.zeros(128, np.uint8)
    _check(lib().mfb200_dist_unique_id(_fp(buf)), "mfb200_dist_unique_id")
    return buf


def dist_rotation(world, rank, substep, stripes_per_rank=1):
    out = np.zeros(5, np.int32)
    lib().mfb200_dist_rotation(world, rank, substep, stripes_per_rank, _fp(out))
    return dict(zip(("compute", "send_stripe", "send_to", "recv_stripe", "recv_from"), (int(x) for x in out)))


PLAN_FIELDS = ("nC", "nWarps", "L", "nG", "S1", "nTB", "nPass", "segS", "segT", "segT2", "swap_sides", "nStripes",
               "stripeRows", "tLo", "tRows", "smem_bytes")


def dist_exchange_plan(world, rank, counts):
    """Offsets of the sharded load's all-to-all.  counts[q, d] = ratings rank q holds for rank d.
    Returns (send_off[world], recv_off[world], number of ratings received)."""
    counts = np.ascontiguousarray(counts, np.uint64)
    so, ro = np.zeros(world, np.int64), np.zeros(world, np.int64)
    f = lib().mfb200_dist_exchange_plan
    f.restype = C.c_longlong
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    nr = f(world, rank, _fp(counts), _fp(so), _fp(ro))
    if nr < 0:
        raise MfError("mfb200_dist_exchange_plan: invalid argument")
    return so, ro, int(nr)


def dist_owner_of_row(t_row, t_seg, world):
    return lib().mfb200_dist_owner_of_row(int(t_row), int(t_seg), int(world))


def plan_band(m, n, nnz, k, world=1, rank=0, sm_count=148, max_smem=232448):
    out = np.zeros(16, np.int32)
    _check(lib().mfb200_plan_band(m, n, nnz, k, world, rank, sm_count, max_smem, _fp(out)), "mfb200_plan_band")
    return dict(zip(PLAN_FIELDS, (int(x) for x in out)))


def plan_kernel(m, n, nnz, k, world=1, rank=0, sm_count=148, max_smem=232448):
    """The SGD kernel the planner picks (codes of mfb200_report.kernel: 1 band, 2 run, 5 run with T-row locks, 6 item)."""
    f = lib().mfb200_plan_kernel
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    return f(m, n, nnz, k, world, rank, sm_count, max_smem)


class Model:
    """A model resident on the device (mfb200_model_*): one upload, any number of predict / metric / top-k calls."""

    def __init__(self, P, Q, b):
        P = np.ascontiguousarray(P, np.float32)
        Q = np.ascontiguousarray(Q, np.float32)
        self.m, self.n, self.k = P.shape[0], Q.shape[0], P.shape[1]
        self.h = lib().mfb200_model_upload(_fp(P), _fp(Q), self.m, self.n, self.k, b)
        if not self.h:
            raise MfError("mfb200_model_upload failed: %s" % lib().mfb200_last_error().decode())

    def predict_pairs(self, pairs):
        pairs = np.ascontiguousarray(pairs, np.float32)
        out = np.empty(len(pairs) // 2, np.float32)
        _check(lib().mfb200_model_predict_pairs(self.h, _fp(pairs), len(out), _fp(out)), "mfb200_model_predict_pairs")
        return out

    def metric(self, which, R):
        R = np.ascontiguousarray(R, dtype=NODE)
        out = C.c_double()
        _check(lib().mfb200_model_metric(self.h, which, _fp(R), len(R), C.byref(out)), "mfb200_model_metric")
        return out.value

    def rmse(self, R):
        return self.metric(P_L2_MFR, R)

    def topk(self, users, topk, scores=True):
        users = np.ascontiguousarray(users, np.int32)
        idx = np.empty((len(users), topk), np.int32)
        sc = np.empty((len(users), topk), np.float32) if scores else None
        _check(lib().mfb200_model_topk(self.h, _fp(users), len(users), topk, _fp(idx), _fp(sc) if scores else None),
               "mfb200_model_topk")
        return idx, sc

    def close(self):
        if self.h:
            lib().mfb200_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def eval_last_ms():
    return lib().mfb200_eval_last_ms()


def topk_last_ms():
    return lib().mfb200_topk_last_ms()


class Session:
    """Staged training with the ratings resident in HBM (mfb200_session_*).  rank/world/nccl_id: one
    process per GPU (mfb200_dist_session_create); epochs/finish/rmse are then collective calls."""

    def __init__(self, m, n, k, iters=20, rank=0, world=1, nccl_id=None, **kw):
        self.m, self.n, self.k = m, n, k
        self.prm = make_param(k, iters, **kw)
        if world > 1:
            nccl_id = np.ascontiguousarray(nccl_id, np.uint8)
            self.h = lib().mfb200_dist_session_create(m, n, C.byref(self.prm), rank, world, _fp(nccl_id))
        else:
            self.h = lib().mfb200_session_create(m, n, C.byref(self.prm))
        if not self.h:
            raise MfError("mfb200_session_create failed: %s" % lib().mfb200_last_error().decode())

    def load(self, R):
        R = np.ascontiguousarray(R, dtype=NODE)
        _check(lib().mfb200_session_load(self.h, _fp(R), len(R)), "mfb200_session_load")

    def reset(self):
        _check(lib().mfb200_session_reset(self.h), "mfb200_session_reset")

    def epochs(self, count):
        ms = C.c_float()
        tr = np.zeros(count, np.float64)
        _check(lib().mfb200_session_epochs(self.h, count, C.byref(ms), _fp(tr)), "mfb200_session_epochs")
        return ms.value, tr

    def finish(self, download=True):
        """scale/shrink/un-permute on the device and download.  download=False (P_out = Q_out = NULL in the C call):
        the final model stays on the device -- with several ranks the call is still collective, but only the ranks
        that want the factors on the host pay for the copy.  Returns (P, Q, b) or (None, None, b)."""
        b = C.c_float()
        if not download:
            _check(lib().mfb200_session_finish(self.h, None, None, C.byref(b)), "mfb200_session_finish")
            return None, None, b.value
        P = np.empty((self.m, self.k), np.float32)
        Q = np.empty((self.n, self.k), np.float32)
        _check(lib().mfb200_session_finish(self.h, _fp(P), _fp(Q), C.byref(b)), "mfb200_session_finish")
        return P, Q, b.value

    def rmse(self, R):
        R = np.ascontiguousarray(R, dtype=NODE)
        out = C.c_double()
        _check(lib().mfb200_session_rmse(self.h, _fp(R), len(R), C.byref(out)), "mfb200_session_rmse")
        return out.value

    def report(self):
        rep = Report()
        lib().mfb200_session_report(self.h, C.byref(rep))
        return rep.as_dict()

    def stream(self):
        return lib().mfb200_session_stream(self.h)

    def close(self):
        if self.h:
            lib().mfb200_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the reference's own entry points (parity tests read like mfTest/mfTest.cpp:74-77) -------------

def php_utility_train(triplets, k, iters, p_l2=0.1, q_l2=0.1, eta=0.1):
    tri = np.ascontiguousarray(triplets, np.float32).ravel()
    lens = C.c_int()
    ptr = lib().php_utility_train(_fp(tri), len(tri) // 3, p_l2, q_l2, k, iters, eta, C.byref(lens))
    if not ptr:
        raise MfError("php_utility_train failed: %s" % lib().mfb200_last_error().decode())
    out = np.ctypeslib.as_array(ptr, shape=(lens.value,)).copy()
    _libc.free(C.cast(ptr, C.c_void_p))
    return out


def php_utility_predict(pairs, model_arr):
    pairs = np.ascontiguousarray(pairs, np.float32).ravel()
    model_arr = np.ascontiguousarray(model_arr, np.float32)
    ptr = lib().php_utility_predict(_fp(pairs), len(pairs) // 2, _fp(model_arr), len(model_arr))
    if not ptr:
        raise MfError("php_utility_predict failed: %s" % lib().mfb200_last_error().decode())
    out = np.ctypeslib.as_array(ptr, shape=(len(pairs) // 2,)).copy()
    _libc.free(C.cast(ptr, C.c_void_p))
    return out
