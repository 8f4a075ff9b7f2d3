"""mf::cos_similarity (mf/mf.cpp:3591-3683): the oracle against golden lists from the compiled
reference (CPU), the device path against both (GPU), through the reference's own entry points."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import mfb200  # noqa: E402
import orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "cos_similarity.npz")
CASES = ["binary_60x12", "binary_200x9_zero_rows", "ints_150x20", "distinct_40x64", "wide_7x300"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_compiled_reference(name):
    g = np.load(GOLD)
    Q, ids, want = g[name + "_Q"], g[name + "_ids"], g[name + "_order"]
    tri = orc.q_triplets(Q)
    for row, item in enumerate(ids):
        got, _ = orc.oracle_cos_similarity(int(item), tri)
        assert np.array_equal(got, want[row]), (name, int(item))


def test_oracle_live_against_the_reference_when_present():
    if not orc.have_ref():
        pytest.skip("compiled reference not present")
    rng = np.random.RandomState(3)
    Q = (rng.rand(90, 15) < 0.4).astype(np.int32)
    tri = orc.q_triplets(Q)
    for item in (0, 44, 89):
        assert np.array_equal(orc.oracle_cos_similarity(item, tri)[0], orc.ref_cos_similarity(item, tri, 90))


def _php_cos(item, tri, items):
    L = mfb200.lib()
    L.php_cos_similarity.restype = C.POINTER(C.c_float)
    L.php_cos_similarity.argtypes = [C.c_int, C.c_void_p, C.c_int]
    tri = np.ascontiguousarray(tri, np.float32)
    p = L.php_cos_similarity(int(item), tri.ctypes.data_as(C.c_void_p), len(tri) // 3)
    assert p
    out = np.ctypeslib.as_array(p, shape=(items,)).copy()
    C.CDLL(None).free(p)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_php_cos_similarity_equals_the_reference(name):
    """The drop-in entry point (php_mf/mfWarp.h:9): the list the compiled reference returned, entry for entry --
    including the order among equal cosines and the places of NaNs."""
    g = np.load(GOLD)
    Q, ids, want = g[name + "_Q"], g[name + "_ids"], g[name + "_order"]
    tri = orc.q_triplets(Q)
    for row, item in enumerate(ids):
        assert np.array_equal(_php_cos(item, tri, Q.shape[0]), want[row]), (name, int(item))


@pytest.mark.gpu
def test_batched_cosines_all_items_at_once():
    """mfb200_cos_similarity: every item against every item in one call; cosines bit-equal to the oracle's, every list
    ordered by (cosine falling, id rising) with zero rows last; sparse triplets (missing cells are 0)."""
    rng = np.random.RandomState(11)
    items, k = 700, 24
    Q = (rng.rand(items, k) < 0.25).astype(np.int32) * rng.randint(1, 3, size=(items, k))
    Q[[5, 600]] = 0
    Q[items - 1, k - 1] = 1  # the matrix size comes from the largest index named
    ii, kk = np.nonzero(Q)
    tri = np.stack([ii, kk, Q[ii, kk]], 1).astype(np.float32).ravel()  # only the non-zero cells
    order, cos_sorted, cos_item, ties = mfb200.cos_similarity(tri)
    assert order.shape == (items, items)
    for a in (0, 5, 123, 699):
        _, want_cos = orc.oracle_cos_similarity(a, orc.q_triplets(Q))
        nan = np.isnan(want_cos)
        assert np.array_equal(np.isnan(cos_item[a]), nan)
        assert np.array_equal(cos_item[a][~nan].view(np.uint32), want_cos[~nan].view(np.uint32))
        key = np.where(nan, -np.inf, want_cos.astype(np.float64))
        want_order = np.lexsort((np.arange(items), -key))
        if not nan.all():
            assert np.array_equal(order[a], want_order), a
    assert ties.shape == (items,) and ties[5] == 1


@pytest.mark.gpu
def test_bad_input_never_returns_null():
    tri = np.array([0, 0, 1, 1, 1, 1], np.float32)
    out = _php_cos(5, tri, 2)  # item id out of range: message on stderr, zeroed list
    assert np.array_equal(out, np.zeros(2, np.float32))
    L = mfb200.lib()
    L.php_DINA.restype = C.POINTER(C.c_int)
    L.php_DINA.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    p = L.php_DINA(tri.ctypes.data_as(C.c_void_p), 2, tri.ctypes.data_as(C.c_void_p), 2, 3)
    assert p and all(p[i] == 0 for i in range(20))  # php_mf.c:1281 reads 20 entries unchecked
    C.CDLL(None).free(p)
