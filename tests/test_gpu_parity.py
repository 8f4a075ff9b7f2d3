"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C-ABI of
include/mfb200.h (ctypes binding question-recommendation-system_b200/mfb200.py) and is compared with
  * the golden vectors taken from the compiled reference (tests/golden/, oracle/make_golden.py),
  * the oracle restatement (oracle/mf_oracle.cpp) on the same seeded inputs.
Bars:
  EXACT mode  bit-exact factors (the arithmetic is the reference's, operation for operation).
  RING mode   held-out RMSE within 0.5 % of the reference's after equal epochs (north_star); in its
              ticket variant (RING_REPRO) also run-to-run bit-reproducibility.
  predict / rmse / top-k   bit-exact values and indices.
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

pytestmark = pytest.mark.gpu

SMALL = ["s_1000x500_k20", "s_300x700_k8", "s_600x400_k128_nan", "s_64x48_k40"]
RMSE_TOL = 0.005  # north_star: held-out RMSE within 0.5 % of the reference after equal epochs


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_device_present():
    assert mfb200.device_count() >= 1


# ---------------------------------------------------------------------------------------- exact mode
@pytest.mark.parametrize("name", SMALL)
def test_exact_mode_bit_exact_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_EXACT)
    assert rep["mode_used"] == mfb200.MODE_EXACT
    assert np.array_equal(bits(P), bits(g["P"]))
    assert np.array_equal(bits(Q), bits(g["Q"]))
    assert np.float32(b) == g["b"]
    # the predict path and the metric on those factors
    T = mfb200.gen_ratings(m, n, nnz, max(nnz // 10, 1))
    assert abs(mfb200.rmse(T, P, Q, b) / float(g["heldout_rmse"]) - 1) < 1e-12
    assert np.array_equal(bits(mfb200.predict_pairs(P, Q, b, g["pairs"])), bits(g["pair_pred"]))


def test_exact_mode_config1_bit_exact(golden_dir):
    """Config #1 of BASELINE.json (10k x 5k, 1M ratings, k=32, 20 epochs) against the reference."""
    import hashlib
    g = np.load(os.path.join(golden_dir, "c1_10kx5k_k32.npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_EXACT)
    step = int(g["row_step"])
    assert np.array_equal(bits(P[::step]), bits(g["P_rows"])) and np.array_equal(bits(Q[::step]), bits(g["Q_rows"]))
    assert hashlib.sha256(P.tobytes()).hexdigest() == str(g["P_sha"])
    assert hashlib.sha256(Q.tobytes()).hexdigest() == str(g["Q_sha"])
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    assert abs(mfb200.rmse(T, P, Q, b) / float(g["heldout_rmse"]) - 1) < 1e-12


@pytest.mark.parametrize("shape", [(500, 300, 20000, 16, 4), (200, 900, 15000, 24, 3), (50, 40, 60, 8, 5),
                                   (3, 4, 8, 8, 30), (40, 30, 1, 12, 2)])
def test_exact_mode_vs_oracle_seeded(shape):
    m, n, nnz, k, it = shape
    R = orc.gen_ratings(m, n, 0, nnz, seed=7)
    Po, Qo, bo, tro, _ = orc.oracle_train(R, m, n, k, it, lam_p=0.03, lam_q=0.08, eta=0.07)
    s = mfb200.Session(m, n, k, it, lam_p=0.03, lam_q=0.08, eta=0.07, mode=mfb200.MODE_EXACT)
    s.load(R)
    _, tr = s.epochs(it)
    P, Q, b = s.finish()
    s.close()
    assert np.array_equal(bits(P), bits(Po)) and np.array_equal(bits(Q), bits(Qo)) and b == bo
    assert np.allclose(tr, tro, rtol=1e-9)  # the tr_rmse column (double sums in a different order)


def test_php_entry_points_mftest_kat(golden_dir):
    """mfTest/mfTest.cpp:74-77 through php_utility_train / php_utility_predict (mfWarp.h:7-8)."""
    g = np.load(os.path.join(golden_dir, "mftest_kat.npz"))
    model = mfb200.php_utility_train(g["triplets"], k=8, iters=30, p_l2=0.1, q_l2=0.1, eta=0.1)
    assert len(model) == 5 + 3 * 8 + 4 * 8 and model[:5].tolist() == [0.0, 3.0, 4.0, 8.0, 4.75]
    assert np.array_equal(bits(model[5:29]), bits(g["P"]).ravel())
    assert np.array_equal(bits(model[29:]), bits(g["Q"]).ravel())
    pred = mfb200.php_utility_predict(g["pairs"], model)
    assert np.array_equal(bits(pred), bits(g["pred"]))
    with pytest.raises(mfb200.MfError):  # wrong length: the reference crashes, we fail loudly
        mfb200.php_utility_predict(g["pairs"], model[:-1])


def test_mangled_mf_train_entry(golden_dir):
    """mf::mf_train(mf_problem const*, mf_parameter) called through its Itanium-mangled symbol."""
    g = np.load(os.path.join(golden_dir, "s_64x48_k40.npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    L = mfb200.lib()
    dflt = getattr(L, mfb200.SYM_MF_DEFAULT_PARAM)
    dflt.restype = mfb200.MfParameter
    prm = dflt()
    prm.k, prm.nr_iters, prm.lambda_p2, prm.lambda_q2, prm.quiet, prm.nr_threads = k, it, 0.05, 0.05, True, 1
    prob = mfb200.MfProblem(m, n, nnz, R.ctypes.data)
    f = getattr(L, mfb200.SYM_MF_TRAIN)
    f.restype = C.POINTER(mfb200.MfModel)
    f.argtypes = [C.POINTER(mfb200.MfProblem), mfb200.MfParameter]
    os.environ["MFB200_MODE"] = "exact"
    mdl = f(C.byref(prob), prm)
    os.environ.pop("MFB200_MODE")
    assert mdl and (mdl.contents.m, mdl.contents.n, mdl.contents.k) == (m, n, k)
    P = np.ctypeslib.as_array(mdl.contents.P, shape=(m, k)).copy()
    Q = np.ctypeslib.as_array(mdl.contents.Q, shape=(n, k)).copy()
    assert np.array_equal(bits(P), bits(g["P"])) and np.array_equal(bits(Q), bits(g["Q"]))
    rm = getattr(L, mfb200.SYM_CALC_RMSE)
    rm.restype = C.c_double
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    tprob = mfb200.MfProblem(m, n, len(T), T.ctypes.data)
    assert abs(rm(C.byref(tprob), mdl) / float(g["heldout_rmse"]) - 1) < 1e-12
    pp = C.pointer(mdl)
    getattr(L, mfb200.SYM_MF_DESTROY)(pp)
    bad = dflt()
    bad.k = 0
    assert not f(C.byref(prob), bad)  # check_parameter -> nullptr, mf/mf.cpp:3312-3313


# ----------------------------------------------------------------------------------------- ring mode
@pytest.mark.parametrize("name", ["s_1000x500_k20", "s_300x700_k8", "s_64x48_k40"])
def test_ring_mode_rmse_parity_small(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, max(nnz // 10, 1))
    P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING)
    assert rep["mode_used"] == mfb200.MODE_RING
    assert np.float32(b) == g["b"]
    got = mfb200.rmse(T, P, Q, b)
    # few epochs on tiny data (down to 64 x 48): the update ORDER differs from the reference's and a few
    # hundred held-out ratings are a noisy estimate, so allow 3 % here (6 % at 64 x 48, where eight runs spread from
    # -0.1 % to +2.3 % with the lock order: profiles/r2_gate_margins.txt); the 0.5 % gate is config #1 below
    assert abs(got / float(g["heldout_rmse"]) - 1) < (0.06 if m == 64 else 0.03), (got, float(g["heldout_rmse"]), rep)


def test_ring_mode_config1_rmse_parity(golden_dir):
    """Config #1: held-out RMSE within 0.5 % of the reference's 0.318745 after 20 epochs."""
    g = np.load(os.path.join(golden_dir, "c1_10kx5k_k32.npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    s = mfb200.Session(m, n, k, it, mode=mfb200.MODE_RING)
    s.load(R)
    ms, tr = s.epochs(it)
    got = s.rmse(T)
    P, Q, b = s.finish()
    rep = s.report()
    s.close()
    assert abs(got / float(g["heldout_rmse"]) - 1) < RMSE_TOL, (got, float(g["heldout_rmse"]), rep)
    assert abs(mfb200.rmse(T, P, Q, b) - got) < 1e-9
    assert np.all(np.diff(tr[1:]) < 0)  # training RMSE falls after the slow-only epoch
    assert not np.isnan(P).any() and not np.isnan(Q).any()


def test_ring_mode_is_reproducible_and_handles_unseen_rows():
    m, n, nnz, k, it = 600, 400, 900, 128, 4  # most rows never rated -> NaN rows (mf/mf.cpp:996-999)
    R = mfb200.gen_ratings(m, n, 0, nnz)
    P1, Q1, b1, _ = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING_REPRO)  # rows by tickets, not locks
    P2, Q2, b2, _ = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING_REPRO)
    assert np.array_equal(bits(P1), bits(P2)) and np.array_equal(bits(Q1), bits(Q2)) and b1 == b2
    seen_u = np.zeros(m, bool)
    seen_u[R["u"]] = True
    seen_v = np.zeros(n, bool)
    seen_v[R["v"]] = True
    assert np.array_equal(np.isnan(P1[:, 0]), ~seen_u) and np.array_equal(np.isnan(Q1[:, 0]), ~seen_v)
    assert not np.isnan(P1[seen_u]).any() and not np.isnan(Q1[seen_v]).any()


def test_ring_mode_equals_exact_arithmetic_on_one_worker(monkeypatch):
    """With a 1x1 ring (one warp) the ring kernel walks the ratings sequentially: its result must
    be a valid sequential SGD pass (same arithmetic as the oracle, different visiting order)."""
    monkeypatch.setenv("MFB200_RING_CTAS", "1")
    monkeypatch.setenv("MFB200_RING_WARPS", "1")
    m, n, nnz, k, it = 300, 200, 20000, 32, 5
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, 2000)
    P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING)
    assert (rep["grid_ctas"], rep["cta_warps"]) == (1, 1)
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    # same arithmetic, different ORDER (one block sorted by row vs the reference's 20x20 grid): order noise only
    assert abs(mfb200.rmse(T, P, Q, b) / orc.oracle_rmse(T, Po, Qo, bo) - 1) < 0.03


@pytest.mark.parametrize("shape", [(2000, 1500, 200000, 64, 6), (1500, 2500, 150000, 128, 5), (5000, 300, 100000, 8, 5),
                                   (900, 700, 50000, 200, 4)])
def test_ring_mode_shapes_against_oracle_rmse(shape):
    """Other shapes (m<n, k=8, k>128 with two vectors per lane): RMSE vs the oracle after equal epochs."""
    m, n, nnz, k, it = shape
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING)
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    got, want = mfb200.rmse(T, P, Q, b), orc.oracle_rmse(T, Po, Qo, bo)
    assert abs(got / want - 1) < 0.02, (got, want, rep)


# -------------------------------------------------------------------------------- predict and metrics
def test_predict_pairs_and_rmse_bit_exact_vs_oracle():
    rng = np.random.RandomState(11)
    m, n, k = 300, 500, 37  # k not a multiple of 4: rows are unaligned
    P = rng.randn(m, k).astype(np.float32)
    Q = rng.randn(n, k).astype(np.float32)
    P[17] = np.nan
    Q[3] = np.nan
    pairs = np.stack([rng.randint(-2, m + 2, 5000), rng.randint(-2, n + 2, 5000)], 1).astype(np.float32).ravel()
    assert np.array_equal(bits(mfb200.predict_pairs(P, Q, 3.5, pairs)), bits(orc.oracle_predict_pairs(P, Q, 3.5, pairs)))
    R = np.empty(20000, orc.NODE)
    R["u"], R["v"], R["r"] = rng.randint(0, m, 20000), rng.randint(0, n, 20000), rng.rand(20000) * 4 + 1
    assert abs(mfb200.rmse(R, P, Q, 3.5) / orc.oracle_rmse(R, P, Q, 3.5) - 1) < 1e-12
    assert mfb200.rmse(R[:0], P, Q, 3.5) == 0.0  # mf/mf.cpp:4318-4319


# ------------------------------------------------------------------------ ticket variant, several GPUs
def test_ring_repro_mode_config1_rmse_parity_and_reproducible(golden_dir):
    """The ticket variant at config #1: same 0.5 % RMSE gate, and two runs agree bit for bit."""
    g = np.load(os.path.join(golden_dir, "c1_10kx5k_k32.npz"))
    m, n, nnz, k, it = (int(g[x]) for x in ("m", "n", "nnz", "k", "iters"))
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    P1, Q1, b1, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING_REPRO)
    P2, Q2, b2, _ = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING_REPRO)
    assert rep["mode_used"] == mfb200.MODE_RING
    assert np.array_equal(bits(P1), bits(P2)) and np.array_equal(bits(Q1), bits(Q2)) and b1 == b2
    got = mfb200.rmse(T, P1, Q1, b1)
    assert abs(got / float(g["heldout_rmse"]) - 1) < RMSE_TOL, (got, float(g["heldout_rmse"]), rep)


def test_ring_mode_many_passes_small_shared_memory(monkeypatch):
    """k=512 rows (2 KB each) with few CTAs: a CTA's S band does not fit shared memory, so the epoch runs in
    several passes; 32 lanes per rating.  RMSE against the oracle."""
    monkeypatch.setenv("MFB200_RING_CTAS", "2")
    m, n, nnz, k, it = 1500, 900, 60000, 512, 3
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, nnz // 10)
    P, Q, b, rep = mfb200.train(R, m, n, k, it, mode=mfb200.MODE_RING)
    plan = mfb200.plan_band(m, n, nnz, k)
    assert plan["nPass"] > 1 and plan["L"] == 32, plan
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it)
    got, want = mfb200.rmse(T, P, Q, b), orc.oracle_rmse(T, Po, Qo, bo)
    assert abs(got / want - 1) < 0.02, (got, want, rep, plan)


def test_two_gpus_stripe_rotation_matches_one_gpu():
    """Two ranks (one process per GPU, NCCL): every rank ends with the same model, and its held-out RMSE
    equals the one-GPU run's within the order noise.  Skipped on a one-GPU box."""
    import json
    import subprocess
    if mfb200.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(ROOT, "tools", "dist_check.py"), "c1", "20"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["all_ranks_same_model"] and abs(d["heldout_rmse"] - d["heldout_rmse_host_model"]) < 1e-9
    assert abs(d["heldout_rmse"] / 0.318745 - 1) < RMSE_TOL, d  # the reference's value at config #1


def test_train_with_validation_prints_the_reference_table(capfd):
    """mf::mf_train_with_validation(tr, va, param) through its mangled symbol, quiet = false: the iteration table
    of fpsg_core (mf/mf.cpp:2818-2832, 2880-2907) with the va_rmse column; tr_rmse and obj equal the oracle's."""
    m, n, nnz, k, it = 300, 200, 20000, 16, 4
    R = mfb200.gen_ratings(m, n, 0, nnz)
    V = mfb200.gen_ratings(m, n, nnz, 3000)
    L = mfb200.lib()
    dflt = getattr(L, mfb200.SYM_MF_DEFAULT_PARAM)
    dflt.restype = mfb200.MfParameter
    prm = dflt()
    prm.k, prm.nr_iters, prm.lambda_p2, prm.lambda_q2, prm.quiet, prm.nr_threads = k, it, 0.05, 0.05, False, 1
    f = getattr(L, "_ZN2mf24mf_train_with_validationEPKNS_10mf_problemES2_NS_12mf_parameterE")
    f.restype = C.POINTER(mfb200.MfModel)
    f.argtypes = [C.POINTER(mfb200.MfProblem), C.POINTER(mfb200.MfProblem), mfb200.MfParameter]
    tr = mfb200.MfProblem(m, n, nnz, R.ctypes.data)
    va = mfb200.MfProblem(m, n, len(V), V.ctypes.data)
    os.environ["MFB200_MODE"] = "exact"
    capfd.readouterr()
    mdl = f(C.byref(tr), C.byref(va), prm)
    os.environ.pop("MFB200_MODE")
    out = capfd.readouterr().out
    assert mdl
    P = np.ctypeslib.as_array(mdl.contents.P, shape=(m, k)).copy()
    Q = np.ctypeslib.as_array(mdl.contents.Q, shape=(n, k)).copy()
    b = mdl.contents.b
    lines = [l for l in out.splitlines() if l.strip()]
    assert lines[0].split() == ["iter", "tr_rmse", "va_rmse", "obj"], lines[:2]
    rows = [l.split() for l in lines[1:1 + it]]
    assert [int(r[0]) for r in rows] == list(range(it))
    _, _, _, tr_o, obj_o = orc.oracle_train(R, m, n, k, it)
    for r, t_o, o_o in zip(rows, tr_o, obj_o):
        assert r[1] == "%.4f" % t_o and r[3] == "%.4e" % o_o, (r, t_o, o_o)
    assert abs(float(rows[-1][2]) - mfb200.rmse(V, P, Q, b)) < 2e-4  # the last row is the final model
    assert all(float(a[2]) > float(c[2]) for a, c in zip(rows[1:], rows[2:]))  # and it falls after the first epochs


# ------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("mode", ["ring", "repro"])
@pytest.mark.parametrize("case", ["one_rating", "one_user", "one_item", "heavy_duplicates", "hot_item", "k8_wide"])
def test_ring_mode_edge_cases(case, mode):
    """Degenerate inputs through the throughput schedule: no time-out, NaN exactly on the unseen rows, finite
    elsewhere, and the training RMSE falls."""
    rng = np.random.RandomState(5)
    k, it = 16, 6
    if case == "one_rating":
        m, n = 40, 30
        R = np.array([(7, 3, 4.0)], dtype=orc.NODE)
    elif case == "one_user":
        m, n = 1, 500
        R = np.zeros(400, orc.NODE); R["v"] = rng.randint(0, n, 400); R["r"] = 1 + 4 * rng.rand(400)
    elif case == "one_item":
        m, n = 700, 1
        R = np.zeros(600, orc.NODE); R["u"] = rng.randint(0, m, 600); R["r"] = 1 + 4 * rng.rand(600)
    elif case == "heavy_duplicates":
        m, n = 50, 40
        R = np.zeros(20000, orc.NODE); R["u"] = rng.randint(0, 5, 20000); R["v"] = rng.randint(0, 4, 20000)
        R["r"] = 3 + 0.1 * rng.randn(20000)
    elif case == "hot_item":
        m, n = 3000, 200
        R = np.zeros(60000, orc.NODE); R["u"] = rng.randint(0, m, 60000)
        R["v"] = np.where(rng.rand(60000) < 0.5, 0, rng.randint(0, n, 60000)); R["r"] = 1 + 4 * rng.rand(60000)
    else:
        m, n, k = 200, 50000, 8
        R = np.zeros(100000, orc.NODE); R["u"] = rng.randint(0, m, 100000); R["v"] = rng.randint(0, n, 100000)
        R["r"] = 1 + 4 * rng.rand(100000)
    s = mfb200.Session(m, n, k, it, mode=mfb200.MODE_RING if mode == "ring" else mfb200.MODE_RING_REPRO)
    s.load(R)
    _, tr = s.epochs(it)
    P, Q, b = s.finish()
    s.close()
    seen_u = np.zeros(m, bool); seen_u[R["u"]] = True
    seen_v = np.zeros(n, bool); seen_v[R["v"]] = True
    assert np.array_equal(np.isnan(P).all(1), ~seen_u) and np.array_equal(np.isnan(Q).all(1), ~seen_v)
    assert np.isfinite(P[seen_u]).all() and np.isfinite(Q[seen_v]).all() and np.isfinite(tr).all()
    if len(R) > 1 and float(R["r"].std()) > 0:
        assert tr[-1] < tr[1] or tr[-1] < tr[0]
    assert abs(b - float(np.float32(R["r"].astype(np.float64).mean()))) < 1e-5


def test_ring_mode_empty_problem():
    s = mfb200.Session(10, 10, 8, 3, mode=mfb200.MODE_RING)
    s.load(np.zeros(0, orc.NODE))
    s.epochs(2)
    P, Q, b = s.finish()
    s.close()
    assert np.isnan(P).all() and np.isnan(Q).all()
