"""BASELINE config 5 / SURVEY.md §8f-2: embedding extraction + fused kNN scoring with the embeddings handed over ON THE
DEVICE through lemon_b200.handoff.extract_and_score (the reference copies every batch to the CPU,
run_lemon.py:158-161,230-233): encoder outputs go straight into the rank's device shard and the replicated database is
all-gathered / normalised chunk by chunk while the encoder is still running.

No weights or datasets are available offline, so two encoders are offered:
  --encoder clip        `transformers.CLIPModel(CLIPConfig())` == ViT-B/32 with RANDOM weights, bf16 autocast, synthetic
                        pixels / token ids: the realistic extraction COST, but a random-init CLIP maps every input to
                        almost the same embedding (a degenerate distribution: no row can be certified, everything
                        takes the exact fp32 fallback) -- it shows the hand-off and the worst case of the scorer;
  --encoder projection  structured inputs (clustered latents) through fixed random 2-layer projections: CLIP-like,
                        non-degenerate embeddings, so the scoring number is the representative one.

  python tools/e2e_clip.py [--pairs 118000] [--batch 512] [--encoder projection]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_clip.py ...
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import lemon_b200
from lemon_b200 import dist as ldist, handoff
from lemon_b200.scoring import count_uncertified
from bench import HP

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=118000)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--encoder", default="projection", choices=["clip", "projection"])
args = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = args.pairs
r0, r1, per = ldist.shard_bounds(n, world, rank)
g = torch.Generator(device=dev).manual_seed(100 + rank)
torch.manual_seed(0)

if args.encoder == "clip":
    from transformers import CLIPConfig, CLIPModel
    model = CLIPModel(CLIPConfig()).to(dev).eval()
    flops_pair = 14.8e9

    def feats(o):
        return o if torch.is_tensor(o) else o.pooler_output

    def enc_img(pix):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return feats(model.get_image_features(pixel_values=pix))

    def enc_txt(ids):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return feats(model.get_text_features(input_ids=ids, attention_mask=torch.ones_like(ids)))

    def batches():
        for b0 in range(0, per, args.batch):
            b = min(args.batch, per - b0)
            ids = torch.randint(0, 49408, (b, 77), generator=g, device=dev)
            ids[:, -1] = 49407                                  # eos position for the pooled output
            yield torch.randn((b, 3, 224, 224), generator=g, device=dev), ids
else:
    latent, hidden, d = 64, 2048, 512
    gw = torch.Generator(device=dev).manual_seed(7)             # same weights / centroids on every rank
    cen = torch.randn(1000, latent, generator=gw, device=dev)
    Wi1, Wi2 = torch.randn(latent, hidden, generator=gw, device=dev) / 8, torch.randn(hidden, d, generator=gw, device=dev) / 45
    Wt1, Wt2 = torch.randn(latent, hidden, generator=gw, device=dev) / 8, torch.randn(hidden, d, generator=gw, device=dev) / 45
    flops_pair = 2 * 2 * (latent * hidden + hidden * d)
    enc_img = lambda p: torch.nn.functional.gelu(p @ Wi1) @ Wi2
    enc_txt = lambda t: torch.nn.functional.gelu(t @ Wt1) @ Wt2 + 0.5 * (torch.nn.functional.gelu(t @ Wi1) @ Wi2)

    def batches():
        for b0 in range(0, per, args.batch):
            b = min(args.batch, per - b0)
            z = torch.randint(0, 1000, (b,), generator=g, device=dev)
            yield cen[z] + 0.6 * torch.randn(b, latent, generator=g, device=dev), cen[z] + 0.6 * torch.randn(b, latent, generator=g, device=dev)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


sc = lemon_b200.get_scorer(lr)
run = lambda: handoff.extract_and_score(batches(), enc_img, enc_txt, n, k=30, hparams=HP, scorer=sc)
out = run()                                                     # warm-up
sync()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
with torch.no_grad():                                           # extraction alone, for the split of the total
    for a, b in batches():
        enc_img(a), enc_txt(b)
e[1].record()
out = run()
e[2].record(); sync()
t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ext_ms, tot_ms = t.tolist()
info = sc.last_info
unc = {s: (info[s].get("n_uncertified_first_pass"), count_uncertified(info[s])) for s in ("img", "txt")}
if rank == 0:
    print(json.dumps({"config": "C5 end-to-end MSCOCO-shaped: embedding extraction (%s) + fused scoring, device hand-off" % args.encoder,
                      "pairs": n, "n_gpus": world, "extract_only_ms": ext_ms, "extract_and_score_ms": tot_ms,
                      "scoring_ms_on_top_of_extraction": tot_ms - ext_ms, "e2e_pairs_per_s": n / (tot_ms * 1e-3),
                      "extract_tflops_per_gpu": (r1 - r0) * flops_pair / (ext_ms * 1e-3) / 1e12,
                      "uncertified_rows_rank0 (first pass, after second pass)": unc,
                      "score_range": [float(out["score"].min()), float(out["score"].max())]}))
if world > 1:
    dist.destroy_process_group()
