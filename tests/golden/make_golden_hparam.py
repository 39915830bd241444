"""Golden vectors for the F1-threshold objective and the hyper-parameter grid, from the reference's OWN functions
run live (lib.metrics.utils.optimize_f1_efficient :286-296, optim_func :117-121, and the grid loop's arithmetic via
calc_scores_given_hparams_vectorized :47-82).  Builder container only:  python tests/golden/make_golden_hparam.py"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_live  # noqa: E402
from oracle import lemon_oracle as O  # noqa: E402
from oracle import hparam_oracle as H  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
COLS = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")


def main():
    mu = ref_live.import_reference_metrics()
    rng = np.random.RandomState(777)
    out = {}
    # ---- optimize_f1_efficient on assorted score distributions
    cases = []
    for i, (n, frac, sep) in enumerate([(200, 0.4, 1.5), (1000, 0.2, 0.7), (5000, 0.4, 1.0), (64, 0.5, 0.1), (333, 0.05, 2.0),
                                        (500, 0.0, 1.0), (500, 1.0, 1.0)]):
        y = (rng.rand(n) < frac).astype(np.int64)
        s = rng.standard_normal(n) + sep * y
        if i == 3:
            s = np.round(s, 1)                  # many tied scores
        f1, thr = mu.optimize_f1_efficient(y, s, return_thres=True)
        out[f"f1_y_{i}"], out[f"f1_s_{i}"], out[f"f1_res_{i}"] = y, s, np.array([f1, thr])
        cases.append(i)
    out["f1_cases"] = np.array(cases)
    # ---- grid stage on LEMoN-like records
    n, k = 400, 8
    from tests.helpers import clustered_pairs
    x, yv, _, mis = clustered_pairs(n, 48, n_clusters=10, seed=12, noise_frac=0.35)
    o = O.lemon_oracle(x, yv, x, yv, k=k, query_in_db=np.arange(n))
    rec = {c: o[c].astype(np.float32) for c in COLS}
    rec["d_1"] = o["d_1"].astype(np.float32).astype(np.float64)
    df = pd.DataFrame([{**{c: rec[c][i] for c in COLS}, "d_1": float(rec["d_1"][i]), "is_mislabel": int(mis[i])} for i in range(n)])
    grid = {"beta": np.arange(0, 20.01, 5), "gamma": np.arange(0, 20.01, 5), "tau_1": [0, 1, 5], "tau_2": [0, 5]}
    pts = H.grid_points(grid)
    vals = []
    for g in pts:
        vals.append(-mu.optim_func(g, df, mu.optimize_f1_efficient, {}))
    for c in COLS + ("d_1",):
        out["grid_" + c] = rec[c]
    out["grid_y"] = mis.astype(np.int64)
    out["grid_points"] = np.array(pts)
    out["grid_f1"] = np.array(vals)
    np.savez_compressed(os.path.join(HERE, "hparam.npz"), **out)
    print("written", len(pts), "grid points; best", max(vals))


if __name__ == "__main__":
    main()
