"""Seam 1 at the reference's call pattern (run_lemon.py:45,235-236: index.search with 128 queries per call): time per
call of faiss_compat.IndexFlatIP.search as a function of the number of DB segments the planner would choose.
    python tools/seam1_bench.py [M] [D]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemon_b200
from lemon_b200 import faiss_compat, scoring
from bench import synth_pairs

m = int(sys.argv[1]) if len(sys.argv) > 1 else 370_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
sc = lemon_b200.get_scorer(0)
x, _, _ = synth_pairs(m, d, 0.0, 1, dev)
xn = sc.prepare(x, True, need_f16=False).f32
index = faiss_compat.IndexFlatIP(d)
index.add(xn)
orig = scoring.plan_segments
for nq in (128, 1024):
    q = xn[:nq * 32]
    for forced in (None, 8, 16, 24, 32, 48, 64):
        scoring.plan_segments = orig if forced is None else (lambda *a, _f=forced, **k: _f)
        for as_numpy in (False, True):
            qq = q.cpu().numpy() if as_numpy else q
            index.search(qq[:nq], 31)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for b in range(0, qq.shape[0], nq):
                D, I = index.search(qq[b:b + nq], 31)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / (qq.shape[0] // nq)
            print(f"m={m} d={d} nq={nq} nseg={'planner' if forced is None else forced} {'numpy' if as_numpy else 'device'} in/out: "
                  f"{dt * 1e3:.3f} ms per call", flush=True)
scoring.plan_segments = orig
