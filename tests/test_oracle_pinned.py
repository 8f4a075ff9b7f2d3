"""Pins the oracle (oracle/mf_oracle.cpp) before anything is allowed to trust it.

1. against the golden vectors in tests/golden/, which oracle/make_golden.py took from the
   compiled reference (mf::mf_train at nr_threads=1, mf::mf_predict, mf::calc_rmse);
2. against the compiled reference itself, live, when oracle/_ref is present;
3. against libc / libstdc++ for the library-defined sequences the reference depends on.
Bar: bit-exact (the path is fp32 with a fixed order).
"""
import ctypes as C
import hashlib
import os
import platform

import numpy as np
import pytest

import orc

SMALL = ["s_1000x500_k20", "s_300x700_k8", "s_600x400_k128_nan", "s_64x48_k40"]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_library_kats(golden_dir):
    g = np.load(os.path.join(golden_dir, "lib_kat.npz"))
    L = orc.oracle()
    out = np.empty(10, np.int32)
    L.orc_kat_random_map(10, out.ctypes.data)
    assert out.tolist() == g["shuffle10"].tolist()
    d = np.empty(4, np.float32)
    L.orc_kat_minstd(4, d.ctypes.data)
    assert np.array_equal(bits(d), bits(g["minstd4"]))
    r = np.empty(2000, np.int32)
    L.orc_kat_glibc_rand(0, 2000, r.ctypes.data)
    assert np.array_equal(r, g["glibc_rand_seed0"])
    # and against this machine's libc, seeds 0 and 12345
    libc = C.CDLL("libc.so.6")
    for seed in (0, 12345):
        libc.srand(seed)
        live = np.array([libc.rand() for _ in range(500)], np.int32)
        L.orc_kat_glibc_rand(seed, 500, r.ctypes.data)
        assert np.array_equal(r[:500], live)


def test_rsqrt_table_properties():
    L = orc.oracle()
    assert bits(np.float32(L.orc_kat_rsqrt(1.0, 0)))[()] == 0x3F7FF000
    assert bits(np.float32(L.orc_kat_rsqrt(2.0, 0)))[()] == 0x3F34F800
    assert bits(np.float32(L.orc_kat_rsqrt(1.5, 0)))[()] == 0x3F510000
    rng = np.random.RandomState(1)
    x = np.exp(rng.uniform(-20, 40, 4000)).astype(np.float32)
    y = np.array([L.orc_kat_rsqrt(float(v), 0) for v in x], np.float32)
    rel = np.abs(y * np.sqrt(x.astype(np.float64)) - 1.0)
    assert rel.max() <= 1.5 * 2.0 ** -12  # the documented RSQRTPS error bound
    y4 = np.array([L.orc_kat_rsqrt(float(v) * 4.0, 0) for v in x], np.float32)
    assert np.array_equal(bits(y4 * 2.0), bits(y))  # exact exponent scaling
    if "intel" in (platform.processor() or "").lower() or "GenuineIntel" in open("/proc/cpuinfo").read():
        yh = np.array([L.orc_kat_rsqrt(float(v), 1) for v in x], np.float32)
        assert np.array_equal(bits(yh), bits(y))  # table == the instruction on Intel


def test_mftest_kat(golden_dir):
    g = np.load(os.path.join(golden_dir, "mftest_kat.npz"))
    m, n, k = 3, 4, 8
    out = np.empty(5 + m * k + n * k, np.float32)
    lens = orc.oracle().orc_utility_train(g["triplets"].ctypes.data, 8, 0.1, 0.1, k, 30, 0.1, 0, out.ctypes.data)
    assert lens == 5 + m * k + n * k
    assert out[:5].tolist() == [0.0, 3.0, 4.0, 8.0, 4.75]
    assert np.array_equal(bits(out[5:5 + m * k]), bits(g["P"]).ravel())
    assert np.array_equal(bits(out[5 + m * k:]), bits(g["Q"]).ravel())
    P, Q = out[5:5 + m * k].reshape(m, k).copy(), out[5 + m * k:].reshape(n, k).copy()
    pred = orc.oracle_predict_pairs(P, Q, float(out[4]), g["pairs"])
    assert np.array_equal(bits(pred), bits(g["pred"]))
    # the values printed in SURVEY.md Appendix B
    assert abs(pred[0] - 5.28189087) < 1e-6 and abs(pred[8] - 7.9743042) < 1e-6


@pytest.mark.parametrize("name", SMALL)
def test_small_cases_against_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = orc.gen_ratings(m, n, 0, nnz)
    assert sha(R) == str(g["R_sha"])
    P, Q, b, tr, ob = orc.oracle_train(R, m, n, k, it)
    assert np.array_equal(bits(P), bits(g["P"]))  # NaN rows compare equal as bit patterns too
    assert np.array_equal(bits(Q), bits(g["Q"]))
    assert np.float32(b) == g["b"]
    T = orc.gen_ratings(m, n, nnz, max(nnz // 10, 1))
    # calc_rmse sums with an OpenMP reduction (mf/mf.cpp:4321-4323): last-bit order noise only
    assert abs(orc.oracle_rmse(T, P, Q, b) / float(g["heldout_rmse"]) - 1) < 1e-12
    assert np.array_equal(bits(orc.oracle_predict_pairs(P, Q, b, g["pairs"])), bits(g["pair_pred"]))
    idx, sc = orc.oracle_topk(P, Q, b, g["topk_users"], g["topk_idx"].shape[1])
    assert np.array_equal(idx, g["topk_idx"])
    assert np.array_equal(bits(sc), bits(g["topk_score"]))


def test_config1_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "c1_10kx5k_k32.npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = orc.gen_ratings(m, n, 0, nnz)
    assert sha(R) == str(g["R_sha"])
    P, Q, b, tr, ob = orc.oracle_train(R, m, n, k, it)
    assert sha(P) == str(g["P_sha"]) and sha(Q) == str(g["Q_sha"])
    assert np.float32(b) == g["b"]
    T = orc.gen_ratings(m, n, nnz, nnz // 10)
    assert abs(orc.oracle_rmse(T, P, Q, b) / float(g["heldout_rmse"]) - 1) < 1e-12
    assert abs(float(g["heldout_rmse"]) - 0.318745) < 1e-6  # SURVEY.md 8d / BASELINE.md
    assert np.all(np.diff(tr[1:]) < 0)  # training RMSE falls monotonically after the slow-only epoch


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("shape", [(500, 300, 20000, 16, 4), (200, 900, 15000, 24, 3), (50, 40, 60, 8, 5)])
def test_live_against_compiled_reference(shape):
    m, n, nnz, k, it = shape
    R = orc.gen_ratings(m, n, 0, nnz, seed=7)
    P, Q, b, _, _ = orc.oracle_train(R, m, n, k, it, lam_p=0.03, lam_q=0.08, eta=0.07, rsqrt_mode=0)
    Pr, Qr, br = orc.ref_train_stable(R, m, n, k, it, lam_p=0.03, lam_q=0.08, eta=0.07, threads=1)
    assert np.array_equal(bits(P), bits(Pr)) and np.array_equal(bits(Q), bits(Qr)) and b == br


# ---- the other MFSolver losses, L1 regularisation, NMF ------------------------------------------
import loss_cases  # noqa: E402

LOSS_PARAMS = [(c, s) for c in loss_cases.CASES for s in loss_cases.SHAPES]
LOSS_IDS = [loss_cases.key(c[0], s) for c, s in LOSS_PARAMS]


@pytest.fixture(scope="module")
def losses_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


@pytest.mark.parametrize("case,shape", LOSS_PARAMS, ids=LOSS_IDS)
def test_losses_against_golden(losses_golden, case, shape):
    """Factors bit-exact to mf::mf_train (nr_threads=1) for every loss; the printed table to its printed digits;
    the matching error measure (calc_mae / calc_gkl / calc_logloss / calc_accuracy) on held-out ratings."""
    name, fun, kw, kind = case
    m, n, nnz, k, it = shape
    key = loss_cases.key(name, shape)
    R = loss_cases.ratings(m, n, 0, nnz, kind)
    P, Q, b, tr, ob = orc.oracle_train_ex(R, m, n, k, it, fun=fun, **kw)
    gP, gQ = losses_golden[key + "_P"], losses_golden[key + "_Q"]
    if fun == orc.P_LR_MFC:
        # exp() is libm's expf, whose last bit may differ between CPUs (ifunc variants): exact here where the fixture
        # was made (see the live test below), 1e-5 elsewhere
        assert np.allclose(P, gP, rtol=0, atol=1e-5) and np.allclose(Q, gQ, rtol=0, atol=1e-5)
    else:
        assert np.array_equal(bits(P), bits(gP)) and np.array_equal(bits(Q), bits(gQ))
    assert np.float32(b) == losses_golden[key + "_b"]
    table = losses_golden[key + "_table"]  # printed with 4 decimals / 5 significant digits
    assert np.all(np.abs(tr - table[:, 0]) <= 0.5e-4 + 1e-9)
    assert np.all(np.abs(ob / table[:, 1] - 1) <= 1e-4)
    T = loss_cases.ratings(m, n, nnz, nnz // 10, kind)
    got = orc.oracle_metric(loss_cases.METRIC_OF[fun], T, P, Q, b)
    assert abs(got - float(losses_golden[key + "_metric"])) <= 1e-9 * max(1.0, abs(got)) + (1e-6 if fun == orc.P_LR_MFC else 0)
    if kw.get("nmf"):
        assert np.nanmin(P) >= 0 and np.nanmin(Q) >= 0


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("case", loss_cases.CASES, ids=[c[0] for c in loss_cases.CASES])
def test_losses_live_against_compiled_reference(case):
    name, fun, kw, kind = case
    m, n, nnz, k, it = 250, 180, 9000, 24, 3
    R = loss_cases.ratings(m, n, 0, nnz, kind)
    P, Q, b, _, _ = orc.oracle_train_ex(R, m, n, k, it, fun=fun, lam_p2=0.03, lam_q2=0.06, eta=0.08, **kw)
    Pr, Qr, br = orc.ref_train_ex_stable(R, m, n, k, it, fun=fun, lam_p2=0.03, lam_q2=0.06, eta=0.08, **kw)
    assert np.array_equal(bits(P), bits(Pr)) and np.array_equal(bits(Q), bits(Qr)) and b == br
    T = loss_cases.ratings(m, n, nnz, nnz // 10, kind)
    w = loss_cases.METRIC_OF[fun]
    assert abs(orc.oracle_metric(w, T, P, Q, b) / orc.ref_metric(w, T, Pr, Qr, br) - 1) < 1e-12


# ---- the one-class BPR losses and their ranking measures (mf/mf.cpp:2131-2707, 4406-4536) -----------
@pytest.mark.parametrize("case", loss_cases.BPR_CASES, ids=[c[0] for c in loss_cases.BPR_CASES])
def test_bpr_against_golden(golden_dir, case):
    """Factors bit-exact to mf::mf_train (fun = 10 / 11, nr_threads=1, after srand(seed)): the scheduler's second block,
    the negatives drawn from the per-block generators, the three-row update; calc_mpr / calc_auc both ways."""
    name, fun, kw, shape, seed = case
    m, n, _, k, it = shape
    g = np.load(os.path.join(golden_dir, "bpr.npz"))
    R = loss_cases.bpr_ratings(shape)
    P, Q, b, tr, ob = orc.oracle_train_ex(R, m, n, k, it, fun=fun, rand_seed=seed, **kw)
    # the loss goes through libm's expf / logf only by way of the gradient scale z: bits equal here where the fixture
    # was made, 1e-5 on a CPU whose libm picks another variant
    assert np.allclose(P, g[name + "_P"], rtol=0, atol=1e-5) and np.allclose(Q, g[name + "_Q"], rtol=0, atol=1e-5)
    if orc.have_ref():
        assert np.array_equal(bits(P), bits(g[name + "_P"])) and np.array_equal(bits(Q), bits(g[name + "_Q"]))
    assert np.float32(b) == g[name + "_b"]
    table = g[name + "_table"]
    assert np.all(np.abs(tr - table[:, 0]) <= 0.5e-4 + 1e-6)
    assert np.all(np.abs(ob / table[:, 1] - 1) <= 1e-4)
    want = g[name + "_mpr_auc"]
    got = orc.oracle_mpr_auc(R, g[name + "_P"], g[name + "_Q"], float(g[name + "_b"]), False) + \
        orc.oracle_mpr_auc(R, g[name + "_P"], g[name + "_Q"], float(g[name + "_b"]), True)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    if kw.get("nmf"):
        assert P.min() >= 0 and Q.min() >= 0


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("fun", [orc.P_ROW_BPR_MFOC, orc.P_COL_BPR_MFOC], ids=["row", "col"])
def test_bpr_live_against_compiled_reference(fun):
    """Another shape (rows without positives stay zero, bins > rows of a band), another seed, prob sizes larger than the
    model's in calc_mpr / calc_auc."""
    m, n, cnt, k, it = 90, 130, 2500, 24, 3
    R = orc.unique_pairs(m, n, cnt, seed=11)
    R = R[(R["u"] != 5) & (R["v"] != 17)]
    P, Q, b, _, _ = orc.oracle_train_ex(R, m, n, k, it, fun=fun, lam_p2=0.03, lam_q2=0.06, eta=0.08, rand_seed=99)
    Pr, Qr, br = orc.ref_train_ex(R, m, n, k, it, fun=fun, lam_p2=0.03, lam_q2=0.06, eta=0.08, rand_seed=99)
    assert np.array_equal(bits(P), bits(Pr)) and np.array_equal(bits(Q), bits(Qr)) and b == br
    # unseen rows start at zero for these losses, not NaN (mf/mf.cpp:997); one that can be drawn as a negative moves
    assert not (Q[17] if fun == orc.P_COL_BPR_MFOC else P[5]).any()
    assert not np.isnan(P).any() and not np.isnan(Q).any()
    for tr in (False, True):
        assert np.allclose(orc.oracle_mpr_auc(R, P, Q, b, tr, prob_m=m + 3, prob_n=n + 2),
                           orc.ref_mpr_auc(R, Pr, Qr, br, tr, prob_m=m + 3, prob_n=n + 2), rtol=1e-12, atol=0)


# ---- cross-validation (mf_cross_validation, mf/mf.cpp:4117-4129, 3208-3286) ---------------------------------------
@pytest.mark.parametrize("case", loss_cases.CV_CASES, ids=[c[0] for c in loss_cases.CV_CASES])
def test_cross_validation_against_golden(golden_dir, case):
    name, m, n, nnz, k, it, folds, bins = case
    _, fun, kw, kind = loss_cases.cv_case(name)
    g = np.load(os.path.join(golden_dir, "cv.npz"))
    R = loss_cases.ratings(m, n, 0, nnz, kind)
    mean, errs = orc.oracle_cross_validation(R, m, n, k, it, folds, fun=fun, bins=bins, **kw)
    want = float(g["%s_%dx%d_f%d" % (name, m, n, folds)])
    assert abs(mean / want - 1) < (1e-6 if fun == orc.P_LR_MFC else 1e-12)
    assert abs(errs.mean() - mean) < 1e-12 and len(set(np.round(errs, 9))) == folds  # the folds differ
    if orc.have_ref():
        assert abs(mean / orc.ref_cross_validation_stable(R, m, n, k, it, folds, fun=fun, bins=bins, **kw) - 1) < 1e-12
