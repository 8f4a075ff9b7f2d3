"""The loss / regulariser combinations of MFSolver (mf/mf.cpp:1749-2126, 1494-1541) that the oracle, the golden
fixture tests/golden/losses.npz (oracle/make_golden.py) and the GPU parity tests share.  TEST INFRASTRUCTURE."""
import numpy as np

import orc

# name, fun, kwargs of oracle_train_ex / ref_train_ex, kind of ratings
CASES = [
    ("l2mfr_l1reg", orc.P_L2_MFR, dict(lam_p1=0.01, lam_q1=0.02), "reg"),
    ("l2mfr_nmf", orc.P_L2_MFR, dict(nmf=True), "reg"),
    ("l1mfr", orc.P_L1_MFR, dict(), "reg"),
    ("l1mfr_l1reg", orc.P_L1_MFR, dict(lam_p1=0.01, lam_q1=0.01), "reg"),
    ("klmfr_nmf", orc.P_KL_MFR, dict(nmf=True), "reg"),
    ("lrmfc", orc.P_LR_MFC, dict(), "cls"),
    ("l2mfc", orc.P_L2_MFC, dict(), "cls"),
    ("l1mfc", orc.P_L1_MFC, dict(), "cls"),
    ("l1mfc_l1reg_nmf", orc.P_L1_MFC, dict(lam_p1=0.005, lam_q1=0.005, nmf=True), "cls"),
]
# shapes: m, n, nnz, k, iters  (m < n, m > n with k > 8)
SHAPES = [(300, 700, 20000, 8, 5), (600, 400, 30000, 40, 4)]

# which error measure belongs to which loss (Utility::calc_error / get_error_legend, mf/mf.cpp:635-674, 745-773)
METRIC_OF = {orc.P_L2_MFR: 0, orc.P_L1_MFR: 1, orc.P_KL_MFR: 2, orc.P_LR_MFC: 5, orc.P_L2_MFC: 6, orc.P_L1_MFC: 6}


def ratings(m, n, first, count, kind):
    """Synthetic ratings of SURVEY.md 8d; the classification losses get labels +1 (rating > 3) / -1."""
    R = orc.gen_ratings(m, n, first, count).copy()
    if kind == "cls":
        R["r"] = np.where(R["r"] > 3.0, 1.0, -1.0).astype(np.float32)
    return R


def key(name, shape):
    return "%s_%dx%d_k%d" % (name, shape[0], shape[1], shape[3])


# cross-validation cases (mf_cross_validation, mf/mf.cpp:4117-4129): case name, m, n, nnz, k, iters, folds, bins
CV_CASES = [("l2mfr", 300, 700, 20000, 8, 4, 5, 20), ("l1mfr", 300, 700, 20000, 8, 4, 5, 20),
            ("lrmfc", 600, 400, 30000, 16, 3, 3, 10), ("l2mfc", 600, 400, 30000, 16, 3, 3, 10),
            ("l2mfr_l1reg", 600, 400, 30000, 16, 3, 4, 10)]


def cv_case(name):
    if name == "l2mfr":
        return name, orc.P_L2_MFR, dict(), "reg"
    return [c for c in CASES if c[0] == name][0]


# the two one-class BPR losses (BPRSolver, mf/mf.cpp:2131-2707) with calc_mpr / calc_auc (4406-4536):
# name, fun, kwargs, (m, n, distinct positive pairs, k, iters), rand_seed = the srand() the process called before training
# (the reference's scheduler seeds the negatives' generators from the process-wide rand(), mf/mf.cpp:103-110)
BPR_CASES = [
    ("rowbpr", orc.P_ROW_BPR_MFOC, dict(), (300, 200, 5000, 16, 3), 1),
    ("rowbpr_k40", orc.P_ROW_BPR_MFOC, dict(), (150, 400, 6000, 40, 4), 7),
    ("rowbpr_l1reg_nmf", orc.P_ROW_BPR_MFOC, dict(lam_p1=0.002, lam_q1=0.003, nmf=True), (150, 400, 6000, 40, 3), 7),
    ("colbpr", orc.P_COL_BPR_MFOC, dict(), (300, 200, 5000, 16, 3), 1),
    ("colbpr_k40", orc.P_COL_BPR_MFOC, dict(lam_p2=0.02, lam_q2=0.07), (150, 400, 6000, 40, 4), 123),
    ("colbpr_l1reg", orc.P_COL_BPR_MFOC, dict(lam_p1=0.002, lam_q1=0.003), (64, 50, 900, 8, 5), 123),
]


def bpr_ratings(shape):
    return orc.unique_pairs(shape[0], shape[1], shape[2])
