"""One warm-up + N measured launches of the tensor-core candidate kernel K1 on a given shape (product library), for
ncu captures:   python tools/k1_launch.py NQ M D [KEEP] [REPS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemon_b200
from bench import synth_pairs

nq, m, d = (int(a) for a in sys.argv[1:4])
keep = int(sys.argv[4]) if len(sys.argv) > 4 else 40
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda", 0)
sc = lemon_b200.get_scorer(0)
x, _, _ = synth_pairs(m, d, 0.0, 1, dev)
dbp = sc.prepare(x, True)
qp = lemon_b200.scoring._slice_prepared(dbp, 0, nq)
del x
sc.knn_candidates(qp, dbp, nseg=1, keep=keep)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps):
    ck, cc, ct, _ = sc.knn_candidates(qp, dbp, nseg=1, keep=keep)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"K1 nq={nq} m={m} d={d} keep={keep}: {ms:.3f} ms -> {2.0 * nq * m * dbp.d16 / ms / 1e9:.1f} TFLOP/s, "
      f"mean list length {float(cc[:nq].float().mean()):.1f}")
