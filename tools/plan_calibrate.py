"""Calibration of scoring.plan_segments: K1 launch time for few query tiles as a function of the number of DB segments.
    python tools/plan_calibrate.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemon_b200
from lemon_b200.scoring import _slice_prepared, plan_segments
from bench import synth_pairs

dev = torch.device("cuda", 0)
sc = lemon_b200.get_scorer(0)
for m, d in ((370_000, 512), (50_000, 512), (1_000_000, 768)):
    x, _, _ = synth_pairs(m, d, 0.0, 1, dev)
    dbp = sc.prepare(x, True)
    del x
    for nq in (128, 256, 1024, 4096, 8448, 12_000):
        qp = _slice_prepared(dbp, 0, nq)
        line = []
        for nseg in (1, 2, 4, 8, 16, 24, 32, 48, 64):
            if nseg > 1 and m // nseg < 4096:
                continue
            for _ in range(2):
                sc.knn_candidates(qp, dbp, nseg=nseg, keep=40)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10):
                cand = sc.knn_candidates(qp, dbp, nseg=nseg, keep=40)
            e1.record(); torch.cuda.synchronize()
            line.append((nseg, e0.elapsed_time(e1) / 10))
        best = min(line, key=lambda t: t[1])
        print(f"m={m} d={d} nq={nq}: planner picks nseg={plan_segments(nq, m, 148, 2, dbp.d16)}; measured "
              + "  ".join(f"{s}:{t:.3f}" for s, t in line) + f"  -> best nseg={best[0]} ({best[1]:.3f} ms)", flush=True)
    del dbp
