"""GPU parity of the one-class BPR losses and their ranking measures (BPRSolver mf/mf.cpp:2131-2335,
ROW_BPR_MFOC / COL_BPR_MFOC 2608-2707, Scheduler::get_bpr_job / get_negative 152-191, 249-280, calc_mpr_auc 4406-4525),
through the C-ABI of include/mfb200.h and the mangled mf:: entry points.  Run with -m gpu on a B200.

Bars:
  training   factors bit-exact to the compiled reference's golden vectors (tests/golden/bpr.npz: mf::mf_train at one
             thread after srand(seed) -- the reference's scheduler seeds the negatives' generators from the process-wide
             rand(), so the caller's srand() is part of the input; the library draws from the same C library);
             the tr_bprloss column of the table to its printed digits
  mpr / auc  equal to the reference's calc_mpr / calc_auc (1e-12 relative; the device counts exact integers per row, the
             host adds the rows in the reference's order)
Nothing here reads /root/reference.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import mfb200  # noqa: E402
import orc  # noqa: E402
import loss_cases  # noqa: E402

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def bpr_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "bpr.npz"))


def _kw(kw):
    return dict(lam_p=kw.get("lam_p2", 0.05), lam_q=kw.get("lam_q2", 0.05), lam_p1=kw.get("lam_p1", 0.0),
                lam_q1=kw.get("lam_q1", 0.0), nmf=kw.get("nmf", False))


@pytest.mark.parametrize("case", loss_cases.BPR_CASES, ids=[c[0] for c in loss_cases.BPR_CASES])
def test_bpr_training_vs_reference_golden(bpr_golden, case):
    name, fun, kw, shape, seed = case
    m, n, _, k, it = shape
    R = loss_cases.bpr_ratings(shape)
    mfb200.srand(seed)
    s = mfb200.Session(m, n, k, iters=it, mode=mfb200.MODE_AUTO, fun=fun, **_kw(kw))
    s.load(R)
    trs = []
    for _ in range(it):
        _, tr = s.epochs(1)
        trs.append(tr[0])
    P, Q, b = s.finish()
    s.close()
    assert np.array_equal(bits(P), bits(bpr_golden[name + "_P"])), "P differs from the reference"
    assert np.array_equal(bits(Q), bits(bpr_golden[name + "_Q"])), "Q differs from the reference"
    assert np.float32(b) == bpr_golden[name + "_b"]
    assert np.all(np.abs(np.array(trs) - bpr_golden[name + "_table"][:, 0]) <= 0.5e-4 + 1e-6)
    if kw.get("nmf"):
        assert P.min() >= 0 and Q.min() >= 0


@pytest.mark.parametrize("case", loss_cases.BPR_CASES, ids=[c[0] for c in loss_cases.BPR_CASES])
def test_mpr_auc_vs_reference_golden(bpr_golden, case):
    name, fun, kw, shape, seed = case
    R = loss_cases.bpr_ratings(shape)
    P, Q, b = bpr_golden[name + "_P"], bpr_golden[name + "_Q"], float(bpr_golden[name + "_b"])
    got = mfb200.mpr_auc(R, P, Q, b, False) + mfb200.mpr_auc(R, P, Q, b, True)
    assert np.allclose(got, bpr_golden[name + "_mpr_auc"], rtol=1e-12, atol=0)


def test_mpr_auc_vs_oracle_edge_cases():
    """Rows without positives, ratings <= 0 (not positives), a row whose every column is a positive (skipped: no
    negative), NaN factor rows (score b, ties with every other NaN row), a problem larger than the model, several
    batches of rows."""
    m, n, k = 700, 900, 24
    rng = np.random.RandomState(5)
    P = (rng.rand(m, k) - 0.3).astype(np.float32)
    Q = (rng.rand(n, k) - 0.3).astype(np.float32)
    P[3] = np.nan
    Q[7] = np.nan
    Q[8] = np.nan
    R = orc.unique_pairs(m, n, 30000, seed=8)
    R = R[R["u"] != 11]
    R["r"][::7] = 0.0
    R["r"][1::13] = -1.0
    full = np.zeros(n, orc.NODE)
    full["u"], full["v"], full["r"] = 20, np.arange(n), 1.0
    R = np.concatenate([R[R["u"] != 20], full])
    for tr in (False, True):
        want = orc.oracle_mpr_auc(R, P, Q, 0.25, tr, prob_m=m + 5, prob_n=n + 9)
        got = mfb200.mpr_auc(R, P, Q, 0.25, tr, prob_m=m + 5, prob_n=n + 9)
        assert np.allclose(got, want, rtol=1e-12, atol=0), (tr, got, want)
    os.environ["MFB200_RANK_BATCH_ROWS"] = "37"  # many small batches of rows instead of one
    try:
        assert np.allclose(mfb200.mpr_auc(R, P, Q, 0.25), orc.oracle_mpr_auc(R, P, Q, 0.25), rtol=1e-12, atol=0)
    finally:
        os.environ.pop("MFB200_RANK_BATCH_ROWS")


def test_mangled_bpr_train_and_calc_mpr_auc(bpr_golden):
    """mf::mf_train with fun = P_ROW_BPR_MFOC and mf::calc_mpr / calc_auc through their Itanium-mangled symbols; like the
    reference, calc_mpr sorts prob->R in place (mf/mf.cpp:4432)."""
    name, fun, kw, shape, seed = loss_cases.BPR_CASES[0]
    m, n, _, k, it = shape
    R = loss_cases.bpr_ratings(shape).copy()
    L = mfb200.lib()
    dflt = getattr(L, mfb200.SYM_MF_DEFAULT_PARAM)
    dflt.restype = mfb200.MfParameter
    prm = dflt()
    prm.fun, prm.k, prm.nr_iters, prm.lambda_p2, prm.lambda_q2, prm.quiet, prm.nr_threads = fun, k, it, 0.05, 0.05, True, 1
    prob = mfb200.MfProblem(m, n, len(R), R.ctypes.data)
    f = getattr(L, mfb200.SYM_MF_TRAIN)
    f.restype = C.POINTER(mfb200.MfModel)
    f.argtypes = [C.POINTER(mfb200.MfProblem), mfb200.MfParameter]
    mfb200.srand(seed)
    mdl = f(C.byref(prob), prm)
    assert mdl and mdl.contents.fun == fun
    P = np.ctypeslib.as_array(mdl.contents.P, shape=(m, k)).copy()
    Q = np.ctypeslib.as_array(mdl.contents.Q, shape=(n, k)).copy()
    assert np.array_equal(bits(P), bits(bpr_golden[name + "_P"])) and np.array_equal(bits(Q), bits(bpr_golden[name + "_Q"]))
    want = bpr_golden[name + "_mpr_auc"]
    for sym, tr, w in (("_ZN2mf8calc_mprEPNS_10mf_problemEPNS_8mf_modelEb", False, want[0]),
                       ("_ZN2mf8calc_aucEPNS_10mf_problemEPNS_8mf_modelEb", False, want[1]),
                       ("_ZN2mf8calc_mprEPNS_10mf_problemEPNS_8mf_modelEb", True, want[2]),
                       ("_ZN2mf8calc_aucEPNS_10mf_problemEPNS_8mf_modelEb", True, want[3])):
        g = getattr(L, sym)
        g.restype = C.c_double
        g.argtypes = [C.POINTER(mfb200.MfProblem), C.POINTER(mfb200.MfModel), C.c_bool]
        assert abs(g(C.byref(prob), mdl, tr) / w - 1) < 1e-12
        key = (R["v"].astype(np.int64) << 32 | R["u"]) if tr else (R["u"].astype(np.int64) << 32 | R["v"])
        assert np.all(np.diff(key) > 0)  # sorted in place, like the reference
    getattr(L, mfb200.SYM_MF_DESTROY)(C.pointer(mdl))


def test_bpr_refuses_the_throughput_schedule():
    R = orc.unique_pairs(100, 80, 1000)
    with pytest.raises(mfb200.MfError):
        mfb200.train(R, 100, 80, 8, 2, mode=mfb200.MODE_RING, fun=mfb200.P_ROW_BPR_MFOC)
