#!/usr/bin/env python
"""Golden vectors for mf::cos_similarity (mf/mf.cpp:3591-3683) from the COMPILED REFERENCE (oracle/_ref), written to
tests/golden/cos_similarity.npz.  Build container only (needs /root/reference through oracle/Makefile).

Every case passes a triplet for every cell of the Q matrix: the reference reads the cells no triplet names
uninitialised.  Cases: 0/1 Q matrices as the PHP caller has them (many equal cosines: the order among them is whatever
the reference's exchange sort leaves), small-integer matrices, zero rows (0/0 -> NaN, which the sort never moves),
values above 1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402

rng = np.random.RandomState(7)
cases = {}


def add(name, Q, ids):
    tri = orc.q_triplets(Q)
    cases[name + "_Q"] = Q.astype(np.int32)
    cases[name + "_ids"] = np.asarray(ids, np.int32)
    cases[name + "_order"] = np.stack([orc.ref_cos_similarity(int(i), tri, Q.shape[0]) for i in ids])


add("binary_60x12", (rng.rand(60, 12) < 0.35).astype(np.int32), [0, 7, 31, 59])
Qz = (rng.rand(200, 9) < 0.3).astype(np.int32)
Qz[[3, 50, 51, 120]] = 0  # zero rows
Qz[0, 0] = 1
add("binary_200x9_zero_rows", Qz, [0, 3, 77, 199])
add("ints_150x20", rng.randint(0, 4, size=(150, 20)).astype(np.int32), [1, 2, 149])
add("distinct_40x64", rng.randint(0, 50, size=(40, 64)).astype(np.int32), [0, 13, 39])
add("wide_7x300", rng.randint(0, 3, size=(7, 300)).astype(np.int32), [0, 6])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cos_similarity.npz"), **cases)
print("wrote", sorted(k for k in cases if k.endswith("_order")))
