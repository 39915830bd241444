"""Times the hyper-parameter grid stage (run_lemon.py:332-337 grid: 21 x 21 x 4 x 4 = 7056 points) on a 5000-row
validation split, k = 30: GPU (one lemon_f1_grid launch) vs the CPU port of the reference loop on a sample of
grid points (extrapolated linearly).  usage: python tools/hparam_bench.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lemon_b200 import hparam_compat as hc
from oracle import hparam_oracle as H, lemon_oracle as O

n, k = 5000, 30
rng = np.random.RandomState(0)
y = (rng.rand(n) < 0.4).astype(np.int64)
rec = {"D_n": -rng.uniform(0.5, 1, (n, k)), "D_m": -rng.uniform(0.5, 1, (n, k)), "dists_tr_n": rng.uniform(0, 1, (n, k)),
       "dists_tr_m": rng.uniform(0, 1, (n, k)), "dists_n": rng.uniform(0, 1, (n, k)) + 0.3 * y[:, None],
       "dists_m": rng.uniform(0, 1, (n, k)) + 0.2 * y[:, None]}
rec = {c: v.astype(np.float32) for c, v in rec.items()}
rec["d_1"] = (rng.uniform(0, 1, n) + 0.3 * y).astype(np.float64)
grid = {"beta": np.arange(0, 100.01, 5), "gamma": np.arange(0, 100.01, 5), "tau_1": [0, 1, 5, 10], "tau_2": [0, 1, 5, 10]}
hc.grid_search(rec, y, grid)                       # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
bx, best, f1, thr = hc.grid_search(rec, y, grid)
torch.cuda.synchronize()
gpu_s = time.perf_counter() - t0
pts = H.grid_points(grid)
sample = pts[:: len(pts) // 24][:24]
t0 = time.perf_counter()
for g in sample:
    s, _, _ = O.calc_scores_vectorized(rec, dict(zip(O.HP_KEYS, g)))
    H.optimize_f1_efficient(y, s)
cpu_s = (time.perf_counter() - t0) / len(sample) * len(pts)
idx = [pts.index(g) for g in sample]
ref = []
for g in sample:
    s, _, _ = O.calc_scores_vectorized(rec, dict(zip(O.HP_KEYS, g)))
    ref.append(H.optimize_f1_efficient(y, s))
print(json.dumps({"grid_points": len(pts), "n_val": n, "k": k, "gpu_seconds": gpu_s, "cpu_port_seconds_extrapolated": cpu_s,
                  "cpu_sample_points": len(sample), "speedup": cpu_s / gpu_s, "best_f1": best, "best_x": bx,
                  "max_abs_f1_diff_on_sample": float(np.abs(f1[idx] - np.array(ref)).max())}))
