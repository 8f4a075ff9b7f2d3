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
