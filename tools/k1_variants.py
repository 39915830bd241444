"""K1 experiments: times one launch shape of the tensor-core candidate kernel under the tuning knobs of a
-DLEMON_TC_EXPERIMENT build (read from the environment at ctx creation).
    LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so python tools/k1_variants.py NQ M D [cfg ...]
cfg = comma separated KEY=VALUE pairs of LEMON_TC_* knobs, e.g.  PACE=0  PACE=24  PACE=24,DEBUG=2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemon_b200
from lemon_b200 import _lib
from bench import synth_pairs

nq, m, d = (int(a) for a in sys.argv[1:4])
cfgs = sys.argv[4:] or ["PACE=0", "PACE=24", "PACE=48", "PACE=12", "PACE=0,DEBUG=2", "PACE=24,DEBUG=2"]
dev = torch.device("cuda", 0)
sc = lemon_b200.get_scorer(0)
x, _, _ = synth_pairs(m, d, 0.0, 1, dev)
dbp = sc.prepare(x, True)
qp = lemon_b200.scoring._slice_prepared(dbp, 0, nq)
del x
keep = int(os.environ.get("K1_KEEP", "40"))
reps = int(os.environ.get("K1_REPS", "3"))
for cfg in cfgs:
    for k in [k for k in os.environ if k.startswith("LEMON_TC_")]:
        del os.environ[k]
    for kv in cfg.split(","):
        k, v = kv.split("=")
        os.environ["LEMON_TC_" + k] = v
    ctx = _lib.Context(0)
    sc.ctx, sc.lib = ctx, ctx.lib
    ck, cc, ct, _ = sc.knn_candidates(qp, dbp, nseg=1, keep=keep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        ck, cc, ct, _ = sc.knn_candidates(qp, dbp, nseg=1, keep=keep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"nq={nq} m={m} d={d} {cfg}: {ms:.3f} ms -> {2.0 * nq * m * dbp.d16 / ms / 1e9:.1f} TFLOP/s, "
          f"mean list length {float(cc[:nq].float().mean()):.1f}", flush=True)
    del ck, cc, ct
