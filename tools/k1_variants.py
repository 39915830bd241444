"""K1 experiments: times one launch shape of the tensor-core candidate kernel under the tuning knobs of a
-DLEMON_TC_EXPERIMENT build (read from the environment at ctx creation).
    LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so python tools/k1_variants.py NQ M D [cfg ...]
cfg = stagger:variant[:debug]   (default list below)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemon_b200
from lemon_b200 import _lib
from bench import synth_pairs

nq, m, d = (int(a) for a in sys.argv[1:4])
cfgs = sys.argv[4:] or ["0:0", "4:0", "8:0", "0:1", "0:2", "0:3", "8:1", "8:3", "0:0:2"]
dev = torch.device("cuda", 0)
sc = lemon_b200.get_scorer(0)
x, _, _ = synth_pairs(m, d, 0.0, 1, dev)
dbp = sc.prepare(x, True)
qp = lemon_b200.scoring._slice_prepared(dbp, 0, nq)
del x
keep = int(os.environ.get("K1_KEEP", "40"))
for cfg in cfgs:
    parts = cfg.split(":")
    os.environ["LEMON_TC_STAGGER"], os.environ["LEMON_TC_VARIANT"] = parts[0], parts[1]
    os.environ["LEMON_TC_DEBUG"] = parts[2] if len(parts) > 2 else "0"
    ctx = _lib.Context(0)
    sc.ctx, sc.lib = ctx, ctx.lib
    for _ in range(2):
        ck, cc, ct, _ = sc.knn_candidates(qp, dbp, nseg=1, keep=keep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 4
    e0.record()
    for _ in range(reps):
        ck, cc, ct, _ = sc.knn_candidates(qp, dbp, nseg=1, keep=keep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"nq={nq} m={m} d={d} stagger={parts[0]} variant={parts[1]} debug={os.environ['LEMON_TC_DEBUG']}: {ms:.3f} ms -> "
          f"{2.0 * nq * m * dbp.d16 / ms / 1e9:.1f} TFLOP/s, mean list length {float(cc[:nq].float().mean()):.1f}", flush=True)
    del ck, cc, ct
