"""BASELINE config 5 / SURVEY.md §8f-2: PyTorch CLIP ViT-B/32 embedding extraction + fused kNN scoring, with the
embeddings handed over ON THE DEVICE (the reference copies every batch to the CPU, run_lemon.py:158-161,230-233).

No weights or datasets are available offline: the model is `transformers.CLIPModel(CLIPConfig())` (default config
== ViT-B/32: projection 512, vision width 768 / patch 32, text width 512 / context 77) with random initialisation,
pixels and token ids are synthetic.  Extraction is stock PyTorch (library code); scoring is lemon_b200.

  python tools/e2e_clip.py [--pairs 118000] [--batch 512]          (1 GPU)
  python -m torch.distributed.run --nproc-per-node 8 ... tools/e2e_clip.py --pairs 118000
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import lemon_b200
from lemon_b200 import dist as ldist
from bench import HP

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=118000)
ap.add_argument("--batch", type=int, default=512)
args = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from transformers import CLIPConfig, CLIPModel
torch.manual_seed(0)
model = CLIPModel(CLIPConfig()).to(dev).eval()
n = args.pairs
r0, r1, per = ldist.shard_bounds(n, world, rank)
img = torch.zeros((per, 512), dtype=torch.float32, device=dev)
txt = torch.zeros((per, 512), dtype=torch.float32, device=dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)


def extract():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for b0 in range(0, r1 - r0, args.batch):
            b = min(args.batch, r1 - r0 - b0)
            pix = torch.randn((b, 3, 224, 224), generator=g, device=dev)
            ids = torch.randint(0, 49408, (b, 77), generator=g, device=dev)
            ids[:, -1] = 49407                                          # eos position for the pooled output
            mask = torch.ones_like(ids)
            fi = model.get_image_features(pixel_values=pix)
            ft = model.get_text_features(input_ids=ids, attention_mask=mask)
            fi = fi if torch.is_tensor(fi) else fi.pooler_output
            ft = ft if torch.is_tensor(ft) else ft.pooler_output
            img[b0:b0 + b] = fi.float()                                 # written straight into the rank's shard
            txt[b0:b0 + b] = ft.float()


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


extract(); out = ldist.score_pairs_sharded(img, txt, n, k=30, hparams=HP)   # warm-up
sync()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); extract(); e[1].record()
out = ldist.score_pairs_sharded(img, txt, n, k=30, hparams=HP)
e[2].record(); sync()
t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ext_ms, sc_ms = t.tolist()
if rank == 0:
    flops_pair = 14.8e9
    print(json.dumps({"config": "C5 end-to-end MSCOCO-shaped: CLIP ViT-B/32 (random init, bf16 autocast) extraction + fused scoring",
                      "pairs": n, "n_gpus": world, "extract_ms": ext_ms, "score_ms": sc_ms,
                      "e2e_pairs_per_s": n / ((ext_ms + sc_ms) * 1e-3), "scoring_share": sc_ms / (ext_ms + sc_ms),
                      "extract_tflops_per_gpu": (r1 - r0) * flops_pair / (ext_ms * 1e-3) / 1e12,
                      "handoff": "embeddings stay on the device (no .cpu() per batch); score in [%.3f, %.3f]" %
                                 (float(out["score"].min()), float(out["score"].max()))}))
if world > 1:
    dist.destroy_process_group()
