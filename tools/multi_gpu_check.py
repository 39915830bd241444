"""Run under torchrun on N GPUs: checks that the row-sharded N-GPU result is BITWISE identical, row by
row, to the single-GPU result (SURVEY.md §8e determinism).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import lemon_b200
from lemon_b200 import dist as ldist
from bench import synth_pairs, HP

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for n, d in ((20011, 512), (9000, 768)):
    x, y, _ = synth_pairs(n, d, 0.4, 7, dev)
    r0, r1, per = ldist.shard_bounds(n, world, rank)
    pad = lambda t: torch.cat([t[r0:r1], torch.zeros(per - (r1 - r0), d, device=dev)])
    sc = lemon_b200.get_scorer(lr)
    out = ldist.score_pairs_sharded(pad(x), pad(y), n, k=30, hparams=HP, scorer=sc)
    full = lemon_b200.score_pairs(x, y, k=30, query_in_db=np.arange(n), hparams=HP, device=lr)
    for c in ("score", "s_n", "s_m", "d_1", "I_n", "I_m", "D_n", "D_m", "dists_n", "dists_m", "dists_tr_n", "dists_tr_m"):
        same = bool((out[c] == full[c][r0:r1]).all())
        ok &= same
        if not same:
            print(f"rank {rank}: MISMATCH in {c} (n={n}, d={d})")
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_BITWISE_OK" if int(t.item()) else "MULTI_GPU_BITWISE_FAIL", "world", world)
dist.destroy_process_group()
