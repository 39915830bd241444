"""Small run of every NON-tensor-core kernel of the library for `compute-sanitizer --tool memcheck` (the persistent tcgen05
kernel is left out on purpose: it spins on mbarriers and would crawl under instrumentation):
    timeout 600 compute-sanitizer --tool memcheck python tools/memcheck_case.py
(compute-sanitizer was closed on the round-2 GPU pool, so this case was not run under it; without the tool it is a
plain smoke run of those kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lemon_b200
from lemon_b200 import baselines
from tests.helpers import clustered_pairs

HP = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}
sc = lemon_b200.get_scorer(0, "exact")
for n, d, k in ((777, 96, 7), (2100, 516, 30)):
    x, y, lab, _ = clustered_pairs(n, d, n_clusters=9, seed=n, dup_text_classes=5)
    out = lemon_b200.score_pairs(x, y, k=k, query_in_db=np.arange(n), hparams=HP, text_label_ids_q=lab, text_label_ids_db=lab,
                                 knn_mode="exact", class_text_emb=y[:5], noisy_label=lab)
    out = lemon_b200.score_pairs(x[:300], y[:300], x[300:], y[300:], k=k, dist_type="euclidean", hparams=HP, knn_mode="exact")
    p = sc.prepare(y, True)
    dd = sc.dedup_finish(p, sc.dedup_start(p), min_saving=0.0)
    assert dd is not None and dd.n_unique == 5
    s = sc.split_operands(p, 0); s2 = sc.split_operands(p, 1)
    idx, val = sc.keep_lowest(out["score"], 123)
    sc.combine_scores({c: out[c] for c in ("D_n", "dists_tr_n", "dists_n", "D_m", "dists_tr_m", "dists_m", "d_1")}, HP)
xs, ys, _, _ = clustered_pairs(70_001, 32, n_clusters=50, seed=3, noise_frac=0.5)
p = sc.prepare(ys, True)
dd = sc.dedup_finish(p, sc.dedup_start(p), min_saving=0.0)          # multi-block radix sort, scans, grouping
idx, val = sc.keep_lowest(torch.randn(300_001, dtype=torch.float64, device="cuda"), 100_000)
x, y, _, _ = clustered_pairs(1500, 64, n_clusters=10, seed=1)
for method in baselines.METHODS:
    baselines.discrepancy_scores(x[:200], y[:200], x, y, k=5, method=method, train=True, device=0)
torch.cuda.synchronize()
print("MEMCHECK_CASE_DONE")
