#!/usr/bin/env python
"""Benchmark of the LEMoN pair-scoring hot path (BASELINE.json metric: pairs scored/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = one pass of the hot path over the workload's pairs: (all-gather of the embedding
shards when N>1) -> K0 normalise/cast -> duplicate grouping -> dists_tr -> K1 tensor-core kNN
candidates (image and text) -> K2a fp32 re-rank + certificate -> GPU exact fallback for uncertified
rows -> K2b records + score.  Inputs are resident in HBM when the timed region starts (`value`);
`e2e` is the same step through the public API with pinned HOST buffers, H2D and D2H copies inside
the region.

After the warm-up a PARITY GATE compares sampled rows of the step's result with the float64 oracle
(SURVEY.md §8c acceptance rule) and the JSON line carries the counts; a run with a wrong row prints
no `value`.  The default workload is C3 (370k x 370k x 512) for every --gpus N: it is the largest
BASELINE.json configuration whose 25 steps fit the driver's per-N time limit on one GPU, and it is
the same fixed problem at N = 1, 2, 4, 8 ("scaling": "strong").  --workload c4 is the 3.3M x 768-d
configuration of the north-star target (8 GPUs).
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
    "c2iid": dict(n=118_000, d=512, k=30, noise=0.0, iid=True,
                  name="stress row: 118k iid-Gaussian unit vectors, 512-d, k=30 (SURVEY.md 8d worst case)"),
    "tiny": dict(n=8_192, d=512, k=30, noise=0.4, name="tiny debug workload: 8192 pairs, 512-d"),
}
METRIC = "LEMoN pairs scored/s"


def synth_pairs(n, d, noise, seed, device, text_classes=0, iid=False):
    """clustered unit-norm CLIP-like embeddings (1000 centroids) + 'cat' caption noise
    (a caption replaced by another caption of the same cluster: exact duplicate text rows)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    if iid:
        x = torch.randn(n, d, generator=g, device=device)
        y = torch.randn(n, d, generator=g, device=device)
        return x.contiguous(), y.contiguous(), torch.zeros(n, dtype=torch.bool, device=device)
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


def workload_pairs(wl, device, seed=1234):
    return synth_pairs(wl["n"], wl["d"], wl["noise"], seed, device, wl.get("text_classes", 0), wl.get("iid", False))


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


def k1_traffic(workload, world):
    """DRAM bytes per K1 launch of THIS run's dominant launch shape from the committed ncu captures
    (profiles/k1_traffic.json: {"<workload>@<n_gpus>": {dram_bytes_per_launch, note}}), or None when no capture of
    that shape exists."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        j = json.load(open(p))
    except Exception:
        return None
    return j.get(f"{workload}@{world}")


def bench_config(wl, world):
    """The `config` object; identical for both arms (--impl lemon_b200 / reference)."""
    n, d = wl["n"], wl["d"]
    return {"workload": wl["name"], "pairs": n, "dim": d, "k": wl["k"], "dist_type": "cosine",
            "database": "all pairs (N==M), train-split self-exclusion", "hparams": HP,
            "parallelism": f"query rows sharded over {world} GPU(s), DB replicated by all-gather",
            "l2_policy": "inputs larger than L2 (fp32+fp16 DB copies = %.0f MB)" % (2 * n * d * 6 / 1e6)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(wl, x, y, n_queries, threads=None):
    """Times the reference CPU scorer port (oracle.reference_cpu_scorer: fp32 normalise -> index build -> per-128-
    batch matmul+topk x2 -> per-sample python loop -> DataFrame -> score) on `n_queries` train queries against the
    FULL workload DB.  The database preparation is paid once per run whatever the number of queries, so the
    full-workload rate is extrapolated as N / (t_prep + N * t_query_per_pair) instead of amortising the preparation
    over the sample.  Returns dict(value, seconds, prep_s, query_s, threads)."""
    import torch
    from oracle import lemon_oracle as O
    if threads:
        torch.set_num_threads(int(threads))
    idx = np.arange(wl["n"])
    tm = {}
    t0 = time.perf_counter()
    df = O.reference_cpu_scorer(x[:n_queries], y[:n_queries], x, y, k=wl["k"], dist_type="cosine",
                                train_indices_in_compr=idx, hparams=HP, timings=tm)
    dt = time.perf_counter() - t0
    assert len(df) == n_queries
    per_pair = tm["query_s"] / n_queries
    return {"value": wl["n"] / (tm["prep_s"] + wl["n"] * per_pair), "seconds": dt, "prep_s": tm["prep_s"],
            "query_s": tm["query_s"], "threads": torch.get_num_threads()}


def host_pairs(wl):
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    x, y, _ = workload_pairs(wl, dev)
    return x.cpu().numpy(), y.cpu().numpy()


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port: faiss is not installable
    here) on ALL host cores; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1        # torchrun exports OMP_NUM_THREADS=1: set the thread count explicitly
    torch.set_num_threads(cores)
    nq = args.ref_queries
    x, y = host_pairs(wl)
    res = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(wl, x, y, nq, cores)
        if i >= args.warmup:
            res.append(r)
    dt = float(np.mean([r["seconds"] for r in res]))
    val = float(np.mean([r["value"] for r in res]))
    sample = (f"{nq} train queries x full {wl['n']}-row DB per step, {cores} threads; value = N / (t_prep + N * t_per_query): "
              f"DB preparation {np.mean([r['prep_s'] for r in res]):.2f} s once + {1e3 * np.mean([r['query_s'] for r in res]) / nq:.3f} ms per query")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(wl, args.gpus),
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": int(res[0]["threads"]), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ parity gate
def parity_gate(out, rows_global, x_host, y_host, wl, n_rows, lab_host=None, seed=7):
    """Sampled rows of a step's result against the float64 oracle (tests.helpers.check_against_oracle, the
    acceptance rule of SURVEY.md §8c).  `out`: this rank's outputs (device tensors, rows r0..r1)."""
    from tests.helpers import check_against_oracle
    r0, r1 = rows_global
    rng = np.random.RandomState(seed)
    take = np.sort(rng.choice(r1 - r0, min(n_rows, r1 - r0), replace=False))
    sub = {c: t[take].cpu().numpy() for c, t in out.items() if c != "rows"}
    sub["I_n"], sub["I_m"] = sub["I_n"].astype(np.int64), sub["I_m"].astype(np.int64)
    g = take + r0
    t0 = time.perf_counter()
    st = check_against_oracle(sub, x_host[g], y_host[g], x_host, y_host, k=wl["k"], dist_type="cosine", query_in_db=g,
                              hparams=HP, lab_q=None if lab_host is None else lab_host[g], lab_db=lab_host, strict=False)
    return {"rows_checked": int(len(take)), "exact": int(min(st["exact_n"], st["exact_m"])),
            "tie_excused": int(st["tie_rows"]), "wrong": int(st["wrong"]), "wrong_rows": st["wrong_rows"],
            "oracle_seconds": round(time.perf_counter() - t0, 2),
            "rule": "neighbour sets == float64 oracle modulo eps-ties (2e-6) at the k-th boundary; records and scores within 1e-5 relative"}


def timed(fn, steps, dev, warm=1):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def secondary_rows(wl, x, y, scorer, dev, x_host, y_host, parity_rows):
    """Single-GPU secondary measurements SURVEY.md §8d / VERDICT ask for: the reference-default DB cap
    (run_lemon.py:48,122-127: M = 50 000 random rows), seam 1 at the reference's call pattern (index.search with
    128 queries per call, run_lemon.py:45,235-236) and the iid-Gaussian stress distribution."""
    import torch
    import lemon_b200
    from lemon_b200 import faiss_compat
    from lemon_b200.scoring import count_uncertified
    n, d, k = wl["n"], wl["d"], wl["k"]
    res = {}
    # ---- (i) DB cap: shuffled 50k-of-N train_indices_in_compr, every pair is a query
    cap = 50_000
    if n > cap:
        idx = lemon_b200.subsample_db(n, cap, np.random.RandomState(99))
        qid = lemon_b200.query_in_db_from_indices(n, idx)
        idx_t = torch.from_numpy(idx).to(dev)
        xdb, ydb = x[idx_t].contiguous(), y[idx_t].contiguous()
        qid_t = torch.from_numpy(qid).to(dev)
        fn = lambda: lemon_b200.score_pairs(x, y, xdb, ydb, k=k, query_in_db=qid_t, hparams=HP, device=dev.index)
        ms = timed(fn, 3, dev)
        out = fn()
        info = scorer.last_info
        from tests.helpers import check_against_oracle
        take = np.sort(np.random.RandomState(5).choice(n, min(parity_rows, 512), replace=False))
        sub = {c: t[take].cpu().numpy() for c, t in out.items()}
        st = check_against_oracle(sub, x_host[take], y_host[take], x_host[idx], y_host[idx], k=k, query_in_db=qid[take],
                                  hparams=HP, strict=False)
        res["db_cap_50k"] = {"queries": n, "db_rows": cap, "ms_per_step": ms, "pairs_per_s": n / (ms * 1e-3),
                             "parity": {"rows_checked": int(len(take)), "tie_excused": st["tie_rows"], "wrong": st["wrong"]},
                             "uncertified_rows": {s: count_uncertified(info[s]) for s in ("img", "txt")},
                             "queries_in_db": int((qid >= 0).sum())}
        del xdb, ydb, out
    # ---- (ii) seam 1: faiss API, 128 queries per call against the full DB
    xn = scorer.prepare(x, True, need_f16=False).f32
    index = faiss_compat.IndexFlatIP(d)
    index.add(xn)
    q = xn[:128 * 64]
    index.search(q[:128], k + 1)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for b in range(0, q.shape[0], 128):
        D, I = index.search(q[b:b + 128], k + 1)
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / (q.shape[0] // 128)
    res["seam1_search_nq128"] = {"db_rows": n, "ms_per_call": dt * 1e3, "queries_per_s": 128 / dt,
                                 "api": "faiss_compat.IndexFlatIP.search(device tensor [128,d], k+1)"}
    del index, xn, q
    # ---- (iii) iid-Gaussian stress: how many rows lose the certificate, what the step costs then
    ni = min(n, 118_000)
    xi, yi, _ = synth_pairs(ni, d, 0.0, 4321, dev, iid=True)
    qi = torch.arange(ni, device=dev)
    fn = lambda: lemon_b200.score_pairs(xi, yi, k=k, query_in_db=qi, hparams=HP, device=dev.index)
    ms = timed(fn, 3, dev)
    info = scorer.last_info
    res["iid_stress"] = {"pairs": ni, "ms_per_step": ms, "pairs_per_s": ni / (ms * 1e-3),
                         "uncertified_rows": {s: count_uncertified(info[s]) for s in ("img", "txt")}}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("LEMON_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="lemon_b200", choices=["lemon_b200", "reference"])
    ap.add_argument("--ref-queries", type=int, default=2048)    # reference arm: queries per step (DB prep is charged pro rata)
    ap.add_argument("--cpu-queries", type=int, default=8192)    # cpu_baseline inside the default run: ~20 s of host work
    ap.add_argument("--parity-rows", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
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
    from lemon_b200.scoring import count_uncertified

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    n, d, k = wl["n"], wl["d"], wl["k"]
    x, y, lab = workload_pairs(wl, dev)
    lab = lab if wl.get("text_classes") else None
    r0, r1, per = ldist.shard_bounds(n, world, rank)

    def padded(t):
        out = torch.zeros((per, t.shape[1]), dtype=t.dtype, device=dev)
        out[: r1 - r0] = t[r0:r1]
        return out
    img_local, txt_local = padded(x), padded(y)
    lab_local = padded(lab.view(-1, 1)).view(-1) if lab is not None else None
    need_host = rank == 0 and not (args.no_parity and args.no_cpu_baseline)
    x_host = x.cpu().numpy() if need_host else None
    y_host = y.cpu().numpy() if need_host else None
    lab_host = lab.cpu().numpy() if (need_host and lab is not None) else None
    keep_dev = world == 1 and not args.no_secondary and wl["n"] <= 400_000
    if not keep_dev:
        del x, y
    scorer = lemon_b200.get_scorer(local_rank)

    def step(img, txt, **kw):
        return ldist.score_pairs_sharded(img, txt, n, k=k, dist_type="cosine", hparams=HP, scorer=scorer,
                                         text_label_ids_local=lab_local, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up + parity gate ----------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
        out = step(img_local, txt_local)
    barrier()
    info = scorer.last_info
    n_unc = {s: count_uncertified(info[s]) for s in ("img", "txt")}
    nseg = {s: info[s].get("nseg") for s in ("img", "txt")}
    parity = None
    if rank == 0 and not args.no_parity:
        n_rows = args.parity_rows if n <= 1_000_000 else min(args.parity_rows, 256)      # the oracle is O(rows x N x d) on the host
        parity = parity_gate(out, (r0, r1), x_host, y_host, wl, n_rows, lab_host)
        parity["uncertified_rows_per_step"] = n_unc
    barrier()

    # ---------------- device-resident timing (`value`) ----------------
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
    k1_main = [(a.elapsed_time(b), f) for a, b, f in k1]
    k1_total_ms = float(np.sum([t for t, _ in k1_main])) / args.steps
    k1_flop_total = float(np.sum([f for _, f in k1_main])) / args.steps
    big = max(f for _, f in k1_main)                      # the dominant launch shape (main launches of the step)
    k1_ms = float(np.mean([t for t, f in k1_main if f == big]))
    k1t = torch.tensor([k1_ms, k1_total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1t, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    k1_ms, k1_total_ms = float(k1t[0].item()), float(k1t[1].item())
    value = n / (ms_per_step * 1e-3)
    del out

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        h_img = img_local.cpu().pin_memory()
        h_txt = txt_local.cpu().pin_memory()
        host_out: dict = {}
        h2d = h_img.numel() * 4 + h_txt.numel() * 4

        def e2e_step():
            # pinned host shards in, every df column + scores in pinned host buffers out (the call returns after the
            # last device->host copy has landed); int32 neighbour ids halve their bytes
            return step(h_img, h_txt, host_out=host_out, index_dtype=torch.int32)
        for _ in range(2):
            o = e2e_step()
        barrier()
        d2h = sum(t.numel() * t.element_size() for name, t in o.items() if name != "rows")
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
               "api": "lemon_b200.dist.score_pairs_sharded(pinned host shards, host_out=pinned buffers, index_dtype=int32) "
                      "-> all df columns + scores on the host"}
        del h_img, h_txt, host_out, o

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    ach = big / (k1_ms * 1e-3) / 1e12
    traffic = k1_traffic(args.workload, world)
    roofline = {"bound": "tensor", "kernel": "knn_tc_kernel (K1: fused tcgen05 similarity + streaming top-k)",
                "achieved": ach, "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": ach / peaks["sustained"],
                "peak_kind": "sustained bf16 cuBLAS, " + peaks["source"], "frac_of_burst_peak": ach / peaks["burst"],
                "algorithmic_flops_per_launch": big, "launch_ms": k1_ms,
                "k1_launches_per_step": len(k1) / args.steps, "k1_share_of_step": k1_total_ms / ms_per_step,
                "k1_all_launches_tflops": k1_flop_total / (k1_total_ms * 1e-3) / 1e12,
                "step_frac_of_peak": 4.0 * n * n * d / world / (ms_per_step * 1e-3) / 1e12 / peaks["sustained"],
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "traffic_note": traffic.get("note") if traffic else "no ncu capture of this workload's launch shape committed"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_sample(wl, x_host, y_host, args.cpu_queries, os.cpu_count())
        cpu = {"value": r["value"], "unit": "pairs/s", "cores": r["threads"], "kind": "port", "seconds": r["seconds"],
               "sample": f"{args.cpu_queries} train queries x full {n}-row DB (oracle.reference_cpu_scorer); value = N / (t_prep + N * "
                         f"t_per_query) with DB preparation {r['prep_s']:.2f} s and {1e3 * r['query_s'] / args.cpu_queries:.3f} ms per query"}
    secondary = None
    if keep_dev:
        secondary = secondary_rows(wl, x, y, scorer, dev, x_host, y_host, args.parity_rows)
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate + f32 exact re-rank", "data": "synthetic",
            "config": bench_config(wl, world), "run_info": {"nseg": nseg, "uncertified_rows_per_step": n_unc},
            "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "secondary": secondary,
            "gpu_launches": int(launches), "clocks": clocks}
    if parity is not None and parity["wrong"] > 0:
        line["value"] = None
        line["error"] = "parity gate failed: %d of %d sampled rows differ from the oracle" % (parity["wrong"], parity["rows_checked"])
        if e2e:
            e2e["value"] = None
        print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        sys.exit(1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
