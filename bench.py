#!/usr/bin/env python
"""Benchmark of the LEMoN pair-scoring hot path (BASELINE.json metric: pairs scored/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

One "step" = one pass of the hot path over the workload's pairs: (all-gather of the embedding
shards when N>1) -> K0 normalise/cast -> dists_tr -> K1 tensor-core kNN candidates (image and
text) -> K2a fp32 re-rank + certificate -> GPU exact fallback for uncertified rows -> K2b
records + score.  Inputs are resident in HBM when the timed region starts (`value`); `e2e` is the
same step through the public API with pinned HOST buffers, H2D and D2H copies inside the region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HP = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}  # train_clip_from_scratch.py:102-109
WORKLOADS = {  # BASELINE.json configs
    "c1": dict(n=50_000, d=512, k=30, noise=0.0, text_classes=10,
               name="C1 CIFAR-10-shaped: 50k pairs, 512-d, k=30, text side = 10 distinct prompt vectors, discrete text metric"),
    "c2": dict(n=118_000, d=512, k=30, noise=0.4, name="C2 MSCOCO-shaped: 118k pairs, 512-d, cat noise 0.4, k=30"),
    "c3": dict(n=370_000, d=512, k=30, noise=0.0, name="C3 MIMIC-CXR-shaped: 370k pairs, 512-d, k=30"),
    "c4": dict(n=3_300_000, d=768, k=30, noise=0.0, name="C4 CC3M-shaped: 3.3M pairs, 768-d, k=30"),
    "tiny": dict(n=8_192, d=512, k=30, noise=0.4, name="tiny debug workload: 8192 pairs, 512-d"),
}
METRIC = "LEMoN pairs scored/s"


def synth_pairs(n, d, noise, seed, device, text_classes=0):
    """clustered unit-norm CLIP-like embeddings (1000 centroids) + 'cat' caption noise
    (a caption replaced by another caption of the same cluster: exact duplicate text rows)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    C = 1000
    cen = torch.randn(C, d, generator=g, device=device)
    cen2 = torch.randn(C, d, generator=g, device=device)
    z = torch.randint(0, C, (n,), generator=g, device=device)
    x = cen[z] + 0.6 * torch.randn(n, d, generator=g, device=device)
    y = 0.5 * x + 0.5 * cen2[z] + 0.6 * torch.randn(n, d, generator=g, device=device)
    mis = torch.zeros(n, dtype=torch.bool, device=device)
    if text_classes:     # classification datasets: every caption is one of `text_classes` prompt embeddings (exact duplicates)
        protos = torch.randn(text_classes, d, generator=g, device=device)
        lab = (z % text_classes).to(torch.int32)
        return x.contiguous(), protos[lab.long()].contiguous(), lab
    if noise > 0:
        order = torch.argsort(z, stable=True)
        counts = torch.bincount(z, minlength=C)
        starts = torch.cumsum(counts, 0) - counts
        pos = torch.empty_like(order)
        pos[order] = torch.arange(n, device=device)
        chosen = torch.randperm(n, generator=g, device=device)[: int(noise * n)]
        cz = z[chosen]
        ok = counts[cz] > 1
        chosen, cz = chosen[ok], cz[ok]
        r = (torch.rand(len(chosen), generator=g, device=device) * (counts[cz] - 1).float()).long()
        r = torch.minimum(r, counts[cz] - 2)
        js = starts[cz] + r
        js = js + (js >= pos[chosen]).long()
        partner = order[js]
        y0 = y.clone()
        y[chosen] = y0[partner]
        mis[chosen] = True
    # raw (un-normalised) embeddings: K0 normalises, as run_lemon.py:163-164 does
    return x.contiguous(), y.contiguous(), mis


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.08)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            if t0 - 0.03 <= ts <= t1 + 0.03:
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
                except ValueError:
                    continue
                for nme, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"burst": float(j["bf16_tflops"]), "sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "hbm_gbs": float(j["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def k1_traffic_per_launch():
    """DRAM bytes per K1 launch from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def cpu_reference_sample(wl, n_queries, seed=1234):
    """Times the reference CPU scorer port (oracle.reference_cpu_scorer: fp32 normalise -> per-128-batch
    matmul+topk x2 -> per-sample python loop -> DataFrame -> score) on `n_queries` train queries against
    the FULL workload DB.  Returns (pairs_per_s, seconds, threads)."""
    import torch
    from oracle import lemon_oracle as O
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    x, y, _ = synth_pairs(wl["n"], wl["d"], wl["noise"], seed, dev, wl.get("text_classes", 0))
    x, y = x.cpu().numpy(), y.cpu().numpy()
    idx = np.arange(wl["n"])
    t0 = time.perf_counter()
    df = O.reference_cpu_scorer(x[:n_queries], y[:n_queries], x, y, k=wl["k"], dist_type="cosine",
                                train_indices_in_compr=idx, hparams=HP)
    dt = time.perf_counter() - t0
    assert len(df) == n_queries
    return n_queries / dt, dt, torch.get_num_threads()


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port: faiss is not
    installable here) on the host cores; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    nq = args.ref_queries
    times = []
    for i in range(args.warmup + args.steps):
        pps, dt, thr = cpu_reference_sample(wl, nq)
        if i >= args.warmup:
            times.append(dt)
    dt = float(np.mean(times))
    val = nq / dt
    sample = f"{nq} train queries x full {wl['n']}-row DB per step (extrapolates linearly in N)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "dist_type": "cosine", "self_exclusion": True, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("LEMON_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="lemon_b200", choices=["lemon_b200", "reference"])
    ap.add_argument("--ref-queries", type=int, default=1024)
    ap.add_argument("--cpu-queries", type=int, default=16384)   # ~15 s of host work on the bench box
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    import lemon_b200
    from lemon_b200 import dist as ldist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    n, d, k = wl["n"], wl["d"], wl["k"]
    x, y, lab = synth_pairs(n, d, wl["noise"], 1234, dev, wl.get("text_classes", 0))
    lab = lab if wl.get("text_classes") else None
    r0, r1, per = ldist.shard_bounds(n, world, rank)

    def padded(t):
        out = torch.zeros((per, t.shape[1]), dtype=t.dtype, device=dev)
        out[: r1 - r0] = t[r0:r1]
        return out
    img_local, txt_local = padded(x), padded(y)
    lab_local = padded(lab.view(-1, 1)).view(-1) if lab is not None else None
    del x, y
    scorer = lemon_b200.get_scorer(local_rank)

    def step(img, txt):
        return ldist.score_pairs_sharded(img, txt, n, k=k, dist_type="cosine", hparams=HP, scorer=scorer,
                                         text_label_ids_local=lab_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (`value`) ----------------
    for _ in range(max(args.warmup, 3)):
        out = step(img_local, txt_local)
    barrier()
    info = scorer.last_info
    n_unc = {s: int(info[s]["n_uncertified"].item()) if "n_uncertified" in info[s] else None for s in ("img", "txt")}
    nseg = {s: info[s].get("nseg") for s in ("img", "txt")}
    scorer.k1_events = []
    l0 = scorer.ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.12)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step(img_local, txt_local)
    e1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    launches = scorer.ctx.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    k1 = scorer.k1_events
    scorer.k1_events = None
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in k1]))
    k1_total_ms = float(np.sum([a.elapsed_time(b) for a, b, _ in k1])) / args.steps
    k1_flop = float(np.mean([f for _, _, f in k1]))
    k1t = torch.tensor([k1_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1t, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    k1_ms = float(k1t.item())
    value = n / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        h_img = img_local.cpu().pin_memory()
        h_txt = txt_local.cpu().pin_memory()
        host_out = None
        h2d = h_img.numel() * 4 + h_txt.numel() * 4

        def e2e_step():
            nonlocal host_out
            o = step(h_img, h_txt)        # pinned host shards go straight into the public API
            o.pop("rows")
            if host_out is None:
                host_out = {name: torch.empty(t.shape, dtype=t.dtype).pin_memory() for name, t in o.items()}
            for name, t in o.items():
                host_out[name].copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()     # the caller holds the results on the host
        for _ in range(2):
            e2e_step()
        barrier()
        d2h = sum(t.numel() * t.element_size() for t in host_out.values())
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        esteps = max(3, args.steps // 2)
        tw0 = time.perf_counter()
        f0.record()
        for _ in range(esteps):
            e2e_step()
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - tw0) * 1e3 / esteps
        ems = torch.tensor([max(f0.elapsed_time(f1) / esteps, wall_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": n / (float(ems.item()) * 1e-3), "unit": "pairs/s", "ms_per_step": float(ems.item()),
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "lemon_b200.dist.score_pairs_sharded(pinned host shards) -> all df columns + scores on host"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    ach = k1_flop / (k1_ms * 1e-3) / 1e12
    traffic = k1_traffic_per_launch()
    roofline = {"bound": "tensor", "kernel": "knn_tc_kernel (K1: fused tcgen05 similarity + streaming top-64)",
                "achieved": ach, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": ach / peaks["sustained"],
                "peak_kind": "sustained bf16 cuBLAS, " + peaks["source"], "frac_of_burst_peak": ach / peaks["burst"],
                "algorithmic_flops_per_launch": k1_flop, "launch_ms": k1_ms,
                "k1_launches_per_step": len(k1) / args.steps, "k1_share_of_step": k1_total_ms / ms_per_step,
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "traffic_note": traffic.get("note") if traffic else "no ncu capture committed yet"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        pps, dt, thr = cpu_reference_sample(wl, args.cpu_queries)
        cpu = {"value": pps, "unit": "pairs/s", "cores": thr, "kind": "port", "seconds": dt,
               "sample": f"{args.cpu_queries} train queries x full {n}-row DB (oracle.reference_cpu_scorer; extrapolates linearly in N)"}
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate + f32 exact re-rank", "data": "synthetic",
            "config": {"workload": wl["name"], "pairs": n, "dim": d, "k": k, "dist_type": "cosine",
                       "database": "all pairs (N==M), train-split self-exclusion", "hparams": HP,
                       "parallelism": f"query rows sharded over {world} GPU(s), DB replicated by all-gather",
                       "l2_policy": "inputs larger than L2 (fp32+fp16 DB copies = %.0f MB)" % (2 * n * d * 6 / 1e6),
                       "nseg": nseg, "uncertified_rows_per_step": n_unc},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
