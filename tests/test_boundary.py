"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the headers
declare and every symbol the reference's libmf.so exports for the path, keeps the reference's POD
layouts, fails loudly without a GPU, and the reference's own callers (php_mf/mfWarp.cpp,
mfTest/mfTest.cpp) compile and link against it UNMODIFIED.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "question-recommendation-system_b200")
sys.path.insert(0, PKG)
import mfb200  # noqa: E402
import orc  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(mfb200.LIB_PATH):
        mfb200.build()
    return mfb200.lib()


def exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", mfb200.LIB_PATH]).decode()
    return {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] in "TW"}


def test_every_declared_c_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mfb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b((?:mfb200|php)_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    missing = names - exported()
    assert not missing, missing


def test_reference_mangled_api_is_exported(lib):
    # the 24 MF_API functions of mf/mf.h:68-151 plus the helpers libphp_mf/mfTest import
    need = """_ZN2mf10mf_predictEPKNS_8mf_modelEii _ZN2mf11mf_my_trainEPKcS1_
    _ZN2mf12calc_loglossEPNS_10mf_problemEPNS_8mf_modelE _ZN2mf12read_problemEPKc _ZN2mf12read_tripletEPfi
    _ZN2mf13calc_accuracyEPNS_10mf_problemEPNS_8mf_modelE _ZN2mf13mf_load_modelEPKc
    _ZN2mf13mf_save_modelEPKNS_8mf_modelEPKc _ZN2mf13utility_trainEPfiddiidRi _ZN2mf14array_to_modelEPfi
    _ZN2mf14cos_similarityEiPfi _ZN2mf14model_to_arrayEPNS_8mf_modelERi _ZN2mf15utility_predictEPfiS0_i
    _ZN2mf16mf_destroy_modelEPPNS_8mf_modelE _ZN2mf16mf_train_on_diskEPKcNS_12mf_parameterE
    _ZN2mf19mf_cross_validationEPKNS_10mf_problemEiNS_12mf_parameterE _ZN2mf20mf_get_default_paramEv
    _ZN2mf24mf_train_with_validationEPKNS_10mf_problemES2_NS_12mf_parameterE
    _ZN2mf27mf_cross_validation_on_diskEPKciNS_12mf_parameterE
    _ZN2mf32mf_train_with_validation_on_diskEPKcS1_NS_12mf_parameterE _ZN2mf4DINAEPfiS0_ii
    _ZN2mf8calc_aucEPNS_10mf_problemEPNS_8mf_modelEb _ZN2mf8calc_gklEPNS_10mf_problemEPNS_8mf_modelE
    _ZN2mf8calc_maeEPNS_10mf_problemEPNS_8mf_modelE _ZN2mf8calc_mprEPNS_10mf_problemEPNS_8mf_modelEb
    _ZN2mf8mf_trainEPKNS_10mf_problemENS_12mf_parameterE _ZN2mf9calc_rmseEPNS_10mf_problemEPNS_8mf_modelE""".split()
    missing = set(need) - exported()
    assert not missing, missing


def test_pod_layouts_match_the_reference():
    assert C.sizeof(mfb200.MfParameter) == 44 and C.sizeof(mfb200.MfProblem) == 24 and C.sizeof(mfb200.MfModel) == 40
    assert mfb200.NODE.itemsize == 12


def test_default_parameters(lib):
    f = getattr(lib, mfb200.SYM_MF_DEFAULT_PARAM)
    f.restype = mfb200.MfParameter
    p = f()  # mf/mf.cpp:4538-4557
    assert (p.fun, p.k, p.nr_threads, p.nr_bins, p.nr_iters) == (0, 8, 12, 20, 20)
    assert (p.lambda_p1, p.lambda_q1, p.do_nmf, p.quiet, p.copy_data) == (0.0, 0.0, False, False, True)
    assert abs(p.lambda_p2 - 0.1) < 1e-7 and abs(p.lambda_q2 - 0.1) < 1e-7 and abs(p.eta - 0.1) < 1e-7
    d = lib.mfb200_default_param()
    assert (d.k, d.nr_bins, d.nr_iters, d.quiet, d.mode) == (8, 20, 20, 0, 0)


def test_generator_equals_the_oracles(lib):
    for (m, n, first, cnt) in [(10000, 5000, 0, 20000), (300, 700, 12345, 5000), (480189, 17770, 99_000_000, 3000)]:
        a = mfb200.gen_ratings(m, n, first, cnt)
        b = orc.gen_ratings(m, n, first, cnt)
        assert a.tobytes() == b.tobytes()


def test_text_model_round_trip(lib, tmp_path):
    save = getattr(lib, mfb200.SYM_SAVE_MODEL)
    load = getattr(lib, mfb200.SYM_LOAD_MODEL)
    load.restype = C.POINTER(mfb200.MfModel)
    destroy = getattr(lib, mfb200.SYM_MF_DESTROY)
    rng = np.random.RandomState(0)
    m, n, k = 5, 4, 3
    P = rng.rand(m, k).astype(np.float32)
    Q = rng.rand(n, k).astype(np.float32)
    P[2] = np.nan  # an unseen row is written as "F 0 0 ..." (mf/mf.cpp:4202-4207)
    mdl = mfb200.MfModel(0, m, n, k, 3.25, P.ctypes.data_as(C.POINTER(C.c_float)), Q.ctypes.data_as(C.POINTER(C.c_float)))
    path = str(tmp_path / "model.txt").encode()
    assert save(C.byref(mdl), path) == 0
    lines = open(path).read().splitlines()
    assert lines[:5] == ["f 0", "m 5", "n 4", "k 3", "b 3.25"] and lines[7].startswith("p2 F 0 0 0")
    back = load(path)
    assert (back.contents.m, back.contents.n, back.contents.k, back.contents.b) == (m, n, k, 3.25)
    P2 = np.ctypeslib.as_array(back.contents.P, shape=(m, k))
    assert np.isnan(P2[2]).all() and np.allclose(np.delete(P2, 2, 0), np.delete(P, 2, 0), rtol=1e-5)
    pp = C.pointer(back)
    destroy(pp)
    assert not pp.contents  # *model = nullptr, mf/mf.cpp:4292


def test_host_scalar_predict_matches_oracle(lib):
    rng = np.random.RandomState(3)
    m, n, k = 7, 9, 13
    P = rng.randn(m, k).astype(np.float32)
    Q = rng.randn(n, k).astype(np.float32)
    Q[4] = np.nan
    mdl = mfb200.MfModel(0, m, n, k, 2.5, P.ctypes.data_as(C.POINTER(C.c_float)), Q.ctypes.data_as(C.POINTER(C.c_float)))
    f = getattr(lib, mfb200.SYM_MF_PREDICT)
    f.restype = C.c_float
    f.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for (u, v) in [(0, 0), (6, 8), (3, 4), (-1, 2), (7, 0), (2, 9)]:
        got = f(C.byref(mdl), u, v)
        want = orc.oracle().orc_predict(P.ctypes.data, Q.ctypes.data, m, n, k, 2.5, u, v)
        assert np.float32(got).tobytes() == np.float32(want).tobytes()


@pytest.mark.skipif(mfb200.lib().mfb200_device_count() > 0 if os.path.exists(mfb200.LIB_PATH) else False,
                    reason="a GPU is present")
def test_fails_loudly_without_a_gpu(lib):
    R = mfb200.gen_ratings(50, 40, 0, 500)
    with pytest.raises(mfb200.MfError) as ei:
        mfb200.train(R, 50, 40, 8, 2)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(mfb200.MfError):
        mfb200.predict_pairs(np.zeros((2, 4), np.float32), np.zeros((2, 4), np.float32), 0.0, [0, 0])
    # the entry points added for the other losses, the error measures and cross-validation have no CPU path either
    with pytest.raises(mfb200.MfError):
        mfb200.train(R, 50, 40, 8, 2, fun=mfb200.P_LR_MFC)
    with pytest.raises(mfb200.MfError):
        mfb200.metric(mfb200.P_L1_MFR, R, np.zeros((50, 4), np.float32), np.zeros((40, 4), np.float32), 0.0)
    with pytest.raises(mfb200.MfError):
        mfb200.cross_validation(R, 50, 40, 8, 2, 5)
    with pytest.raises(mfb200.MfError):
        mfb200.topk(np.zeros((50, 4), np.float32), np.zeros((40, 4), np.float32), 0.0, np.arange(4, dtype=np.int32), 3)


def test_cross_validation_rejects_bad_fold_counts(lib):
    """Host-side argument checks of mfb200_cross_validation run before any device work."""
    R = mfb200.gen_ratings(50, 40, 0, 500)
    for folds in (0, 401, 1):  # nr_bins = 20: 400 blocks; one fold would hide every block
        with pytest.raises(mfb200.MfError):
            mfb200.cross_validation(R, 50, 40, 8, 2, folds)


def test_default_param_of_the_c_abi_is_the_l2_path(lib):
    p = lib.mfb200_default_param()
    assert (p.fun, p.lambda_p1, p.lambda_q1, p.do_nmf) == (0, 0.0, 0.0, 0)
    assert (p.k, p.nr_bins, p.nr_iters) == (8, 20, 20) and abs(p.lambda_p2 - 0.1) < 1e-7 and abs(p.eta - 0.1) < 1e-7


@pytest.mark.skipif(not os.path.exists("/root/reference/php_mf/mfWarp.cpp"), reason="reference tree not present")
def test_reference_callers_link_unmodified(lib, tmp_path):
    """php_mf/mfWarp.cpp and mfTest/mfTest.cpp, as they lie in the reference tree, against OUR libmf.so."""
    libdir = os.path.dirname(mfb200.LIB_PATH)
    so = str(tmp_path / "libmfwarp.so")
    subprocess.check_call(["/usr/bin/g++", "-w", "-O2", "-fPIC", "-shared", "-o", so, "/root/reference/php_mf/mfWarp.cpp",
                           "-L" + libdir, "-lmf", "-Wl,-rpath," + libdir])
    W = C.CDLL(so)
    for name in ("php_mf_my_train", "php_utility_train", "php_utility_predict", "php_cos_similarity", "php_DINA"):
        assert hasattr(W, name)
    exe = str(tmp_path / "mfTest")
    subprocess.check_call(["/usr/bin/g++", "-w", "-O2", "-o", exe, "/root/reference/mfTest/mfTest.cpp", "-L" + libdir, "-lmf",
                           "-Wl,-rpath," + libdir])
    und = subprocess.check_output(["nm", "-D", "--undefined-only", exe]).decode()
    wanted = {l.split()[-1] for l in und.splitlines() if "_ZN2mf" in l}
    assert wanted and wanted <= exported()


# ---- text model format on the fast path (csrc/model_text.cpp; mf_save_model / mf_load_model, mf/mf.cpp:4184-4278) -----
def _model_io(L):
    save = getattr(L, mfb200.SYM_SAVE_MODEL)
    save.restype = C.c_int
    save.argtypes = [C.POINTER(mfb200.MfModel), C.c_char_p]
    load = getattr(L, mfb200.SYM_LOAD_MODEL)
    load.restype = C.POINTER(mfb200.MfModel)
    load.argtypes = [C.c_char_p]
    return save, load


def _as_model(P, Q, b):
    return mfb200.MfModel(0, P.shape[0], Q.shape[0], P.shape[1], b, P.ctypes.data_as(C.POINTER(C.c_float)),
                          Q.ctypes.data_as(C.POINTER(C.c_float)))


def _factors(mdl):
    c = mdl.contents
    return (np.ctypeslib.as_array(c.P, shape=(c.m, c.k)).copy(), np.ctypeslib.as_array(c.Q, shape=(c.n, c.k)).copy(), c.b)


def test_text_model_bytes_equal_the_references(lib, golden_dir, tmp_path):
    """The file we write is byte for byte the file the reference's mf_save_model wrote (golden, oracle/make_golden.py);
    loading it gives back the floats the text denotes."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden
    P, Q = make_golden.model_text_case()
    save, load = _model_io(lib)
    path = str(tmp_path / "ours.txt")
    assert save(C.byref(_as_model(P, Q, np.float32(3.14159274))), path.encode()) == 0
    want = open(os.path.join(golden_dir, "model_text_ref.txt"), "rb").read()
    assert open(path, "rb").read() == want
    P2, Q2, b2 = _factors(load(os.path.join(golden_dir, "model_text_ref.txt").encode()))
    assert b2 == np.float32(3.14159)
    assert np.isnan(P2[1]).all() and np.isnan(Q2[4]).all()
    seen = ~np.isnan(P[:, 0])
    assert np.array_equal(P2[seen], np.array([[np.float32("%g" % v) for v in row] for row in P[seen]], np.float32))
    assert np.signbit(P2[3, 1]) and P2[3, 1] == 0  # "-0" survives


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("shape", [(3, 2, 1), (700, 300, 40), (20000, 9000, 128)])
def test_text_model_interchange_with_the_compiled_reference(lib, tmp_path, shape):
    """Both directions, live: same bytes out, same floats in -- also above the size where rows are handled in parallel."""
    import hashlib
    m, n, k = shape
    rng = np.random.RandomState(m)
    P = (rng.standard_normal((m, k)) * rng.choice([1e-6, 1e-3, 1, 1e4, 1e12], size=(m, 1))).astype(np.float32)
    Q = rng.rand(n, k).astype(np.float32)
    P[m // 2] = np.nan
    ref_lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmf_ref.so"))
    save, load = _model_io(lib)
    rsave, rload = _model_io(ref_lib)
    ours, theirs = str(tmp_path / "ours.txt"), str(tmp_path / "ref.txt")
    mdl = _as_model(P, Q, np.float32(2.5))
    assert save(C.byref(mdl), ours.encode()) == 0 and rsave(C.byref(mdl), theirs.encode()) == 0
    digest = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()
    assert digest(ours) == digest(theirs)
    a, b = _factors(load(theirs.encode())), _factors(rload(ours.encode()))
    for x, y in zip(a[:2], b[:2]):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    assert a[2] == b[2]


def test_text_model_unwritable_path_and_missing_file(lib, tmp_path):
    save, load = _model_io(lib)
    P = np.zeros((2, 2), np.float32)
    assert save(C.byref(_as_model(P, P, 0.0)), str(tmp_path / "no_such_dir" / "m.txt").encode()) == 1  # mf/mf.cpp:4187-4188
    assert not load(str(tmp_path / "missing.txt").encode())  # nullptr, mf/mf.cpp:4230-4231


# ---- read_problem (mf/mf.cpp:4143-4182) on the fast path ---------------------------------------------------------------
def _read_problem(L):
    f = getattr(L, "_ZN2mf12read_problemEPKc")
    f.restype = mfb200.MfProblem
    f.argtypes = [C.c_char_p]
    return f


def _nodes(p):
    if p.nnz == 0:
        return np.zeros(0, orc.NODE)
    return np.ctypeslib.as_array(C.cast(p.R, C.POINTER(C.c_byte)), shape=(p.nnz * 12,)).view(orc.NODE).copy()


@pytest.mark.parametrize("nnz", [0, 1, 777, 600000])
def test_read_problem_text(lib, tmp_path, nnz):
    """One "u v r" triple per line; m, n = largest id + 1.  600 000 lines is above the size where the file is parsed in
    chunks by several threads (8 MB per chunk)."""
    R = orc.gen_ratings(9000, 4000, 0, max(nnz, 1))[:nnz]
    path = str(tmp_path / "tr.txt")
    with open(path, "w") as f:
        f.write("".join("%d %d %g\n" % (u, v, r) for u, v, r in R))
    p = _read_problem(lib)(path.encode())
    assert p.nnz == nnz
    got = _nodes(p)
    assert np.array_equal(got["u"], R["u"]) and np.array_equal(got["v"], R["v"])
    assert np.array_equal(got["r"], np.array([np.float32("%g" % r) for r in R["r"]], np.float32))
    assert (p.m, p.n) == ((int(R["u"].max()) + 1, int(R["v"].max()) + 1) if nnz else (0, 0))
    if orc.have_ref():  # the reference itself on the same file
        q = _read_problem(C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmf_ref.so")))(path.encode())
        assert (q.m, q.n, q.nnz) == (p.m, p.n, p.nnz) and np.array_equal(_nodes(q), got)


def test_read_problem_stops_at_the_first_bad_token(lib, tmp_path):
    path = str(tmp_path / "odd.txt")
    open(path, "w").write("1 2 3.5 4 5 +6\n7 8 9e0\n10 11 x 12 13 14\n")  # the stream loop ends at "x" (mf/mf.cpp:4168)
    p = _read_problem(lib)(path.encode())
    got = _nodes(p)
    assert p.nnz == 3 and got["u"].tolist() == [1, 4, 7] and got["r"].tolist() == [3.5, 6.0, 9.0] and (p.m, p.n) == (8, 9)
    assert _read_problem(lib)(str(tmp_path / "missing.txt").encode()).nnz == 0


def test_the_two_copies_of_the_rsqrt_table_are_one_table():
    """oracle/rsqrt12_table.h (the checker's) and csrc/rsqrt12_table.h (the device's) are two copies of the table that
    oracle/gen_rsqrt_table.c generates and checks against the rsqrtps instruction: they must not drift apart."""
    import re
    a = open(os.path.join(ROOT, "oracle", "rsqrt12_table.h")).read()
    b = open(os.path.join(ROOT, "question-recommendation-system_b200", "csrc", "rsqrt12_table.h")).read()
    num = re.compile(r"0x[0-9a-fA-F]+|\b\d+\b")
    body = lambda t: num.findall(t[t.index("{"):])  # noqa: E731
    assert body(a) == body(b) and len(body(a)) >= 2048
