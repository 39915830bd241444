"""Hyper-parameter stage: oracle vs the reference's golden vectors on CPU; GPU kernel vs both."""
import os

import numpy as np
import pytest

from oracle import hparam_oracle as H
from oracle import lemon_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hparam.npz"))
COLS = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")
GRID = {"beta": np.arange(0, 20.01, 5), "gamma": np.arange(0, 20.01, 5), "tau_1": [0, 1, 5], "tau_2": [0, 5]}


def test_oracle_f1_objective_is_bit_exact_vs_reference():
    for i in G["f1_cases"]:
        f1, thr = H.optimize_f1_efficient(G[f"f1_y_{i}"], G[f"f1_s_{i}"], True)
        assert f1 == G[f"f1_res_{i}"][0] and thr == G[f"f1_res_{i}"][1], i


def test_oracle_grid_matches_reference():
    rec = {c: G["grid_" + c] for c in COLS + ("d_1",)}
    pts = H.grid_points(GRID)
    assert np.array_equal(np.array(pts), G["grid_points"])          # same visiting order as utils.py:167-181
    _, best, vals = H.grid_search(rec, G["grid_y"], GRID)
    n = len(G["grid_y"])
    assert np.abs(vals - G["grid_f1"]).max() <= 2.0 / n             # float64 oracle vs the reference's fp32 exp: borderline samples
    assert abs(best - G["grid_f1"].max()) <= 2.0 / n


def test_brent_on_smooth_function():
    x, fx, nfev = H.brent_bounded(lambda t: (t - 0.3) ** 2, -1.0, 2.0, xatol=1e-10)
    assert abs(x - 0.3) < 1e-7 and nfev < 60


@pytest.mark.gpu
def test_gpu_f1_objective_is_bit_exact():
    from lemon_b200 import hparam_compat as hc
    for i in G["f1_cases"]:
        f1, thr = hc.optimize_f1_efficient(G[f"f1_y_{i}"], G[f"f1_s_{i}"], return_thres=True)
        assert f1 == G[f"f1_res_{i}"][0] and thr == G[f"f1_res_{i}"][1], (i, f1, thr, G[f"f1_res_{i}"])


@pytest.mark.gpu
def test_gpu_grid_matches_reference_and_oracle():
    from lemon_b200 import hparam_compat as hc
    rec = {c: G["grid_" + c] for c in COLS + ("d_1",)}
    bx, best, f1, thr = hc.grid_search(rec, G["grid_y"], GRID)
    n = len(G["grid_y"])
    assert f1.shape == G["grid_f1"].shape
    assert np.abs(f1 - G["grid_f1"]).max() <= 2.0 / n
    _, obest, ovals = H.grid_search(rec, G["grid_y"], GRID)
    assert np.abs(f1 - ovals).max() <= 2.0 / n and abs(best - obest) <= 2.0 / n
    # thresholds reproduce their F1 on the oracle's scores
    pts, eff = hc.grid_points(GRID)
    for g in (0, 37, len(pts) - 1):
        s, _, _ = O.calc_scores_vectorized(rec, dict(zip(O.HP_KEYS, eff[g])))
        assert abs(H.f1_at_threshold(G["grid_y"], s, thr[g]) - f1[g]) <= 2.0 / n
    # force_zero / force_one follow unpack_vector (utils.py:84-103)
    bx2, best2, f1b, _ = hc.grid_search(rec, G["grid_y"], GRID, force_zero=["gamma"], force_one=["beta"])
    _, eff2 = hc.grid_points(GRID, ["gamma"], ["beta"])
    assert all(e[0] == 1.0 and e[1] == 0.0 for e in eff2)
    s, _, _ = O.calc_scores_vectorized(rec, dict(zip(O.HP_KEYS, eff2[5])))
    assert abs(H.optimize_f1_efficient(G["grid_y"], s) - f1b[5]) <= 2.0 / n
