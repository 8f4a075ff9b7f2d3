"""File-to-file entry point of the PHP surface on the GPU: php_mf_my_train (php_mf/mfWarp.h:6 -> mf::mf_my_train,
mf/mf.cpp:3397-3413): text ratings in, 40 epochs with the default parameters, text model out.  Run with -m gpu.
(The file sorts last on purpose: it was added after the last GPU session of round 1.)"""
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


def test_php_mf_my_train_file_to_file(tmp_path):
    m0, n0, nnz = 300, 200, 6000
    R = mfb200.gen_ratings(m0, n0, 0, nnz)
    m, n = int(R["u"].max()) + 1, int(R["v"].max()) + 1
    tr, out = str(tmp_path / "tr.txt"), str(tmp_path / "model.txt")
    with open(tr, "w") as f:
        f.write("".join("%d %d %g\n" % (u, v, r) for u, v, r in R))
    Rt = R.copy()
    Rt["r"] = np.array([np.float32("%g" % r) for r in R["r"]], np.float32)  # the ratings the file denotes
    L = mfb200.lib()
    L.php_mf_my_train.restype = C.c_int
    L.php_mf_my_train.argtypes = [C.c_char_p, C.c_char_p]
    assert L.php_mf_my_train(tr.encode(), out.encode()) == 0  # mf_save_model's status, mf/mf.cpp:3407-3411
    lines = open(out).read().splitlines()
    assert lines[:4] == ["f 0", "m %d" % m, "n %d" % n, "k 8"]
    assert len(lines) == 5 + m + n
    # below 262 144 ratings the library trains in the reference's single-thread order: the model is the oracle's,
    # written with six significant digits
    Po, Qo, bo, _, _ = orc.oracle_train(Rt, m, n, 8, 40, lam_p=0.1, lam_q=0.1, eta=0.1)
    assert np.float32(lines[4].split()[1]) == np.float32("%g" % bo)
    for rows, want, tag in ((lines[5:5 + m], Po, "p"), (lines[5 + m:], Qo, "q")):
        for i, (line, w) in enumerate(zip(rows, want)):
            f = line.split()
            assert f[0] == "%s%d" % (tag, i)
            if np.isnan(w[0]):
                assert f[1] == "F"
                continue
            assert f[1] == "T"
            assert [np.float32(x) for x in f[2:]] == [np.float32("%g" % v) for v in w], (tag, i)
