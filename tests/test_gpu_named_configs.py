"""RMSE parity at the NAMED configurations of BASELINE.json (north_star: held-out RMSE within 0.5 % of the reference
after equal epochs) and on skewed item popularity.

The reference values come from the compiled reference run in the build container (oracle/make_golden_named.py ->
tests/golden/named_configs.json): nothing here reads /root/reference.  Everything goes through the C-ABI."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "question-recommendation-system_b200"))
import mfb200  # noqa: E402
import orc  # noqa: E402

pytestmark = pytest.mark.gpu
RMSE_TOL = 0.005
LAM, ETA = 0.05, 0.1


def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "named_configs.json")))


def _train_and_score(name, epochs):
    g = golden()[name]
    m, n, nnz, k = g["m"], g["n"], g["nnz"], g["k"]
    R = mfb200.gen_ratings(m, n, 0, nnz)
    T = mfb200.gen_ratings(m, n, nnz, min(nnz // 10, 10_000_000))
    P, Q, b, rep = mfb200.train(R, m, n, k, epochs, lam_p=LAM, lam_q=LAM, eta=ETA, mode=mfb200.MODE_RING)
    assert rep["mode_used"] == mfb200.MODE_RING
    return mfb200.rmse(T, P, Q, b), float(g["runs"][str(epochs)]["heldout_rmse"]), rep


def test_config2_movielens_shape_20_epochs_rmse_parity():
    """BASELINE.json configs[1]: 138k x 27k, 20M ratings, k=128, one B200, 20 epochs, vs the reference at nr_threads=1."""
    got, want, rep = _train_and_score("c2", 20)
    assert abs(got / want - 1) < RMSE_TOL, (got, want, rep)


# After FEW epochs the comparison is looser, and deliberately so.  During the fast part of the descent (epochs 3-7 at
# these shapes) the held-out RMSE depends on how the ratings of a row are grouped in time, not only on how many have
# been applied: the reference's 20 x 20 grid visits a user 20 times per epoch with ~10 ratings each, the band schedule
# 148 times with ~1.4.  Measured on a B200 at config #3 after 5 epochs (profiles/r2_transient_vs_granularity.txt):
# reference 0.3401 (1 thread) / 0.3413 (8 threads); this engine 0.3273 with 20 CTAs, 0.3342 with 40, 0.3533 with 74,
# 0.3565 with 148 (the default) -- the same kernel, only the grouping changes -- and all of them meet the reference's
# 0.3078 within 0.05 % after 20 epochs.  So these tests bound the transient gap (a schedule that skipped or repeated
# ratings would be far outside it); the 0.5 % bar of north_star is asserted at the configured 20 epochs.
def test_config2_after_5_epochs_same_epoch_reference():
    got, want, rep = _train_and_score("c2", 5)
    assert abs(got / want - 1) < 0.02, (got, want, rep)


def test_config3_netflix_shape_5_epochs_same_epoch_reference():
    """BASELINE.json configs[2] (the configuration the kernel is tuned on: 148 CTAs x 64 groups), 5 epochs against the
    reference's value after 5 epochs (the reference ran with 8 threads: one draw, about 1e-3 of run-to-run spread)."""
    got, want, rep = _train_and_score("c3", 5)
    assert abs(got / want - 1) < 0.06, (got, want, rep)
    assert rep["grid_ctas"] == 148 and rep["kernel"] == 2


def test_config3_netflix_shape_20_epochs_rmse_parity():
    got, want, rep = _train_and_score("c3", 20)
    assert abs(got / want - 1) < RMSE_TOL, (got, want, rep)


@pytest.mark.parametrize("mode", [mfb200.MODE_RING, mfb200.MODE_RING_REPRO])
def test_zipf_item_popularity_no_timeout_and_rmse(mode):
    """Item popularity ~ 1/rank: the most popular of 17.8k items receives 1/15 of all ratings, so one shared-memory row
    is wanted by every group of its CTA all the time.  The conflict-free schedule serialises those updates; it must
    neither give up (wait limit) nor lose accuracy.  Oracle = the reference's sequential order on the same ratings.
    Equal epochs, and enough of them to be past the order-dependent transient of the descent (see the note above;
    tools/zipf_check.py measured on a B200, relative to the oracle: after 6 epochs +0.06 % with locks but +2.1 % with
    tickets, after 12 epochs +0.4 % / +0.9 %, after 20 epochs +0.10 % / +0.26 %)."""
    m, n, nnz, k, it = 60000, 17800, 5_000_000, 32, 20
    R = orc.gen_ratings_zipf(m, n, 0, nnz)
    T = orc.gen_ratings_zipf(m, n, nnz, 500_000)
    counts = np.bincount(R["v"], minlength=n)
    assert counts.max() > nnz // 20  # the skew is real
    Po, Qo, bo, _, _ = orc.oracle_train(R, m, n, k, it, lam_p=LAM, lam_q=LAM, eta=ETA)
    want = orc.oracle_rmse(T, Po, Qo, bo)
    P, Q, b, rep = mfb200.train(R, m, n, k, it, lam_p=LAM, lam_q=LAM, eta=ETA, mode=mode)
    got = mfb200.rmse(T, P, Q, b)
    assert abs(got / want - 1) < RMSE_TOL, (got, want, rep)
