"""Host-side mirror of the LEMoN scoring path (run_lemon.py:163-176, 235-307 and
lib/metrics/utils.py:47-82) on top of liblemon_b200.so.

PyTorch is used for device memory, streams and (in dist.py) NCCL plumbing only; all
arithmetic on the path runs in the hand-written sm_100a kernels behind the C ABI.
There is no CPU fallback: without a B200 and the built library every call raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

HP_KEYS = ("beta", "gamma", "tau_1_n", "tau_2_n", "tau_1_m", "tau_2_m")
KPRIME = 64
LIST_CAP = 1024
MAX_KP = 64
MAX_D_TC = 768


def acc_eps_coef(d16: int, d: int) -> float:
    """RELATIVE bound (a factor of ||q|| * max||b||) on the accumulation error the re-rank certificate must allow
    for: the tensor-core inner product adds d16 exact fp16 x fp16 products into an fp32 accumulator (gamma_n bound
    with truncating adds: d16 * 2^-23 * sum|q_i b_i| <= d16 * 2^-23 * ||q|| ||b||), and the fp32 re-evaluation
    every kernel reports (warp_pair_value: d/32 sequential fma per lane + a 5-level butterfly) adds
    (d/32 + 6) * 2^-24.  Both scale with the norms, so un-normalised inputs (faiss_compat, normalize=False) stay
    rigorously bounded; tests/test_gpu_parity.py::test_tc_error_bound_is_rigorous checks it on the GPU."""
    return float(d16) * 2.0 ** -23 + (d / 32.0 + 6.0) * 2.0 ** -24

METRIC = {"ip": 0, "cosine": 0, "l2": 1, "euclidean": 1}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_dev(x, device, dtype):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


@dataclass
class Prepared:
    """One embedding matrix staged for the kernels: fp32 master (normalised when asked),
    fp16 tensor-core operand padded to a multiple of 64 columns, rounding statistics."""
    f32: torch.Tensor          # [n, d]   (d padded to a multiple of 4)
    f16: torch.Tensor | None   # [n, d16]
    row_stats: torch.Tensor    # [n, 4]  {||x||, ||x16||, ||x-x16||, ||x||^2}
    stats_max: torch.Tensor    # [4]     maxima over rows (last: | ||x||^2 - 1 |)
    n: int
    d: int
    d16: int
    dedup: "Dedup | None" = None   # set for databases with many bit-identical rows


@dataclass
class Dedup:
    """Exact-duplicate structure of a database: the search runs on `uniq` (one representative per group of
    bit-identical rows, in ascending order of the representative's DB index) and is expanded back."""
    uniq: "Prepared"
    offsets: torch.Tensor    # int64 [n_unique + 1]
    members: torch.Tensor    # int32 [n]: DB rows grouped by unique row, ascending inside a group
    n_unique: int


def decode_candidates(cand_keys: torch.Tensor, cand_cnt: torch.Tensor, n_rows: int):
    """Host-side view of K1's output for tests/debugging: per row the valid (value, db_row) pairs of all lists,
    sorted by value descending then row ascending.  Returns (vals float32 [n, L], idx int64 [n, L]) padded with
    (-inf, -1)."""
    k = cand_keys[:n_rows].cpu().numpy().view(np.uint64)
    c = cand_cnt[:n_rows].cpu().numpy()
    n, nlist, cap = k.shape
    valid = np.arange(cap)[None, None, :] < c[:, :, None]
    hi = (k >> np.uint64(32)).astype(np.uint32)          # raw fp32 bit pattern of the approximate value
    lo = (k & np.uint64(0xffffffff)).astype(np.uint32)
    vals = hi.view(np.float32).astype(np.float32)
    idx = (~lo).astype(np.int64)
    vals = np.where(valid, vals, -np.inf).reshape(n, -1)
    idx = np.where(valid, idx, -1).reshape(n, -1)
    order = np.lexsort((idx, -vals), axis=1)
    return np.take_along_axis(vals, order, 1), np.take_along_axis(idx, order, 1)


def subsample_db(n_train: int, limit: int = 50_000, rng=None) -> np.ndarray:
    """run_lemon.py:122-127 (`--compr_dataset_size_limit`, default 50 000 at :48): the kNN database is the whole
    train split, or `limit` rows drawn without replacement (unsorted, global numpy RNG unless `rng` is given)
    when the split is larger.  Returns ``train_indices_in_compr``: DB row j holds train sample out[j]."""
    if n_train > limit:
        rng = np.random if rng is None else rng
        return rng.choice(np.arange(n_train), limit, replace=False)
    return np.arange(n_train)


def query_in_db_from_indices(n_queries: int, train_indices_in_compr) -> np.ndarray:
    """The ``query_in_db`` argument of score_pairs for train-split queries 0 .. n_queries-1: the DB row that holds
    each sample, or -1 when the subsample left it out.  The reference only tests membership
    (`sample_idx in train_indices_in_compr`, run_lemon.py:258,278, an O(M) scan per sample); one scatter does it
    for all samples."""
    idx = np.asarray(train_indices_in_compr.cpu() if torch.is_tensor(train_indices_in_compr) else train_indices_in_compr,
                     dtype=np.int64).reshape(-1)
    out = np.full(int(n_queries), -1, dtype=np.int64)
    ok = (idx >= 0) & (idx < n_queries)
    out[idx[ok]] = np.nonzero(ok)[0]
    return out


def acc_eps_coef_split(d16: int, d: int) -> float:
    """Accumulation bound of the split-precision second pass (operands [hi | lo | hi] x [lo | hi | hi], d16 columns
    per block).  The two cross terms are accumulated first, while the partial sums are <= 2^-10 ||q|| ||b||, so only
    the last d16 additions round at full magnitude: d16 * 2^-23 * (1 + 2^-8) covers all 3 * d16 of them.  Plus the
    fp32 re-evaluation term of acc_eps_coef and the neglected q_lo.b_lo product (<= 2^-22 ||q|| ||b||, counted 2^-21)."""
    return float(d16) * 2.0 ** -23 * (1.0 + 2.0 ** -8) + (d / 32.0 + 6.0) * 2.0 ** -24 + 2.0 ** -21


def count_uncertified(info: dict) -> "int | None":
    """Rows of a kNN call (``LemonScorer.last_info`` of one side) that the re-rank certificate sent to the exact
    fp32 kernel.  Reads the device counters: synchronises."""
    c = info.get("n_uncertified")
    if c is None:
        return None
    return int(sum(int(t.item()) if torch.is_tensor(t) else int(t) for t in (c if isinstance(c, (list, tuple)) else [c])))


def tc_keep(kp: int) -> int:
    """Columns every K1 list certifies above its threshold: the top kp plus a margin for the fp16 rounding
    error (the re-rank certificate decides; rows it cannot certify take the exact kernel)."""
    return int(min(MAX_KP, max(32, kp + 9)))


def _slice_prepared(p: "Prepared", r0: int, r1: int) -> "Prepared":
    return Prepared(p.f32[r0:r1], None if p.f16 is None else p.f16[r0:r1], p.row_stats[r0:r1], p.stats_max,
                    r1 - r0, p.d, p.d16, None)


def plan_segments(nq: int, m: int, num_sms: int, cta_group: int, d16: int = 512) -> int:
    """Number of DB segments the tensor-core kernel scans independently.  Work items are (row tile, segment); more
    segments fill the CTA pairs when there are fewer row tiles than pairs, at the price of 2 more candidate lists per
    row and segment for the re-rank and of a fixed cost per wave of items.  Calibrated on B200 with the round-2 kernel
    (profiles/r02_plan_calibrate.log): a wave costs its columns plus ~0.08 ms (d = 512) / ~0.14 ms (d = 768), i.e. the
    MMA time of ~6000 columns either way; one row tile against 370 000 columns: 5.1 ms unsplit, 0.42 ms in 16
    segments, 0.25 ms in 32, 0.16 ms in 64.  Launches of two or more full rounds are never split (their lists would double for a gain
    the tail launch of plan_tail gets more cheaply, and only unsplit launches pace their DB walk)."""
    units = max(1, num_sms // cta_group)
    tiles = max(1, -(-nq // (128 * cta_group)))
    if tiles >= 2 * units:
        return 1
    overhead_cols = 6000.0
    # at most 32 segments: beyond that the re-rank's walk over 2 * nseg lists per row and the candidate-list allocation
    # cost what the shorter K1 items save (profiles/r02_seam1_search_nq128.log: 0.70 ms per 128-query call at 32)
    cands = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32}
    if tiles < units:
        cands.add(min(32, units // tiles))          # exactly one wave
    best, best_cost = 1, None
    for nseg in sorted(cands):
        if nseg > 1 and m // nseg < 4096:
            break
        items = tiles * nseg
        waves = -(-items // units)
        cost = waves * (m / nseg + overhead_cols)
        if best_cost is None or cost < best_cost * 0.97:
            best, best_cost = nseg, cost
    return best


def plan_tail(nq: int, m: int, num_sms: int, cta_group: int, d16: int = 512) -> tuple[int, int]:
    """Wave-quantisation fix for the tensor-core kernel.  Its work items are 128*cta_group-row query tiles
    dealt to num_sms/cta_group CTA (pairs); when the last round is not full, the rows of that round are searched by a
    second launch with the DB split into segments, so that its items fill whole (shorter) waves.  The number of
    segments minimises waves * (columns per segment + per-wave overhead), the cost model of plan_segments, and the
    split must save at least 10 % of a round.  Databases beyond ~1 GB of fp16 keep the round-1 rule (split only a
    mostly empty last round): a segmented launch walks its segments unpaced, and at that size every pair would stream
    its segment from DRAM.  Returns (rows of the main launch, segments of the tail launch); (nq, 1) means one launch."""
    units = max(1, num_sms // cta_group)
    rows_per_tile = 128 * cta_group
    tiles = -(-nq // rows_per_tile)
    full = (tiles // units) * units
    tail = tiles - full
    if full == 0 or tail == 0:
        return nq, 1
    if 2.0 * m * d16 > 1.0e9 and tail > 0.6 * units:
        return nq, 1
    overhead_cols = 6000.0
    best, best_cost = 1, m + overhead_cols
    for nseg in range(2, 17):
        if m // nseg < 4096:
            break
        waves = -(-(tail * nseg) // units)
        cost = waves * (m / nseg + overhead_cols)
        if cost < best_cost:
            best, best_cost = nseg, cost
    if best == 1 or best_cost > 0.9 * (m + overhead_cols):
        return nq, 1
    return full * rows_per_tile, best


def plan_parts(nq: int, m: int, num_sms: int, cta_group: int, max_rows: int = 1 << 19,
               d16: int = 512) -> list[tuple[int, int, int | None]]:
    """K1 launches of one kNN call as (row0, row1, nseg or None = planner's choice).  The main part is cut into
    whole rounds of query tiles (one round = one 128*cta_group-row tile per CTA pair, so every cut launch is as
    efficient as the uncut one) of at most `max_rows` rows: a launch's candidate lists take 2 * 8 KB per row, which
    bounds them to ~8.6 GB however many query rows there are.  The tail part comes from plan_tail."""
    n_main, nseg_tail = plan_tail(nq, m, num_sms, cta_group, d16)
    round_rows = max(1, num_sms // cta_group) * 128 * cta_group
    step = max(1, max_rows // round_rows) * round_rows
    parts: list[tuple[int, int, int | None]] = []
    if n_main >= nq:
        cuts = list(range(0, nq, step)) + [nq]
        if len(cuts) > 2 and cuts[-1] - cuts[-2] < round_rows:      # do not leave a short last launch
            del cuts[-2]
        for a, b in zip(cuts[:-1], cuts[1:]):
            parts.append((a, b, None if len(cuts) == 2 else 1))
        if len(parts) > 1:                                          # the last piece may again want a tail split
            a, b, _ = parts.pop()
            nm2, ns2 = plan_tail(b - a, m, num_sms, cta_group, d16)
            parts += [(a, b, None)] if nm2 >= b - a else [(a, a + nm2, 1), (a + nm2, b, ns2)]
        return parts
    for a in range(0, n_main, step):
        parts.append((a, min(a + step, n_main), 1))
    parts.append((n_main, nq, nseg_tail))
    return parts


class LemonScorer:
    """Drop-in engine for the scoring path.  Mirrors the order of run_lemon.py:
    ``set_database`` == lines 163-176 (normalise, dists_tr, index.add),
    ``score``        == lines 235-307 for all queries of a split + utils.py:47-82."""

    def __init__(self, device=None, knn_mode: str = "auto", cta_group: int = 0, dedup: bool = True):
        if not torch.cuda.is_available():
            raise _lib.LemonError("lemon_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        self.ctx = _lib.get_context(self.device.index)
        self.lib = self.ctx.lib
        assert knn_mode in ("auto", "tc", "exact")
        self.knn_mode = knn_mode
        self.cta_group = cta_group
        self.dedup = dedup
        self.num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.db = None
        self._pin, self._pin_next = None, 0
        self.second_pass_enabled = True      # split-precision tensor-core pass for uncertified rows (tests switch it off)
        self.last_info: dict = {}
        self.k1_events: list | None = None

    def _pinned_slot(self) -> torch.Tensor:
        """One int32 of pinned host memory for an asynchronous counter read-back (a small ring, allocated once)."""
        if self._pin is None:
            self._pin = torch.empty(256, dtype=torch.int32).pin_memory()
        self._pin_next = (self._pin_next + 1) % 256
        return self._pin[self._pin_next:self._pin_next + 1]

    # ------------------------------------------------------------------ K0
    def prepare(self, x, normalize: bool = True, need_f16: bool = True) -> Prepared:
        """K0.  `x` may be a row-strided 2-D device view (unit column stride), e.g. the embedding columns of a
        wider all-gathered matrix: the kernel reads it in place."""
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
                and x.stride(0) >= x.shape[1]):
            x = _to_dev(x, self.device, torch.float32)
        assert x.dim() == 2
        n, d = x.shape
        if d % 4:
            x = torch.nn.functional.pad(x, (0, 4 - d % 4))
            d = x.shape[1]
        d16 = -(-d // 64) * 64
        out32 = torch.empty((n, d), dtype=torch.float32, device=self.device)
        out16 = torch.empty((n, d16), dtype=torch.float16, device=self.device) if need_f16 else None
        row_stats = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        stats_max = torch.empty(4, dtype=torch.float32, device=self.device)     # zeroed by the library call
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_normalize_cast(self.ctx.handle, _ptr(x), _ptr(out32), _ptr(out16),
                                                         _ptr(row_stats), _ptr(stats_max), n, d, d16, x.stride(0) if n > 1 else d,
                                                         int(bool(normalize)), _stream()), "lemon_normalize_cast")
        return Prepared(out32, out16, row_stats, stats_max, n, d, d16)

    def _aligned_ws(self, nbytes: int) -> torch.Tensor:
        ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
        return ws[(-ws.data_ptr()) % 256:]

    def dedup_start(self, p: Prepared) -> "dict | None":
        """Queues the duplicate COUNT of a database (lemon_dedup_count: row hashes into an open-addressing table, two
        kernels) and an asynchronous read-back of the counter.  No host synchronisation happens here;
        ``dedup_finish`` waits for the counter and only then, for databases that are worth it, runs the grouping."""
        n = p.n
        if n < 64:
            return None
        pend = {"ws": self._aligned_ws(self.lib.lemon_dedup_count_workspace_bytes(n)),
                "counters": torch.empty(2, dtype=torch.int32, device=self.device),
                "host": self._pinned_slot(), "event": torch.cuda.Event()}
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_dedup_count(self.ctx.handle, _ptr(p.f32), n, p.d, _ptr(pend["ws"]),
                                                      _ptr(pend["counters"]), _stream()), "lemon_dedup_count")
        pend["host"].copy_(pend["counters"][:1], non_blocking=True)
        pend["event"].record()
        return pend

    def dedup_finish(self, p: Prepared, pend: "dict | None", min_saving: float = 0.1) -> "Dedup | None":
        """Reads the duplicate count (one host round trip per database; the number of unique rows sizes the search
        operands).  Fewer than `min_saving` duplicate rows: None.  Otherwise the rows are grouped on the device
        (lemon_dedup_build: radix sort of (hash, row) -> run scan -> bit-wise verification -> renumbering) and its
        counters are read back; a hash collision (two different rows, one 63-bit hash) also returns None."""
        if pend is None:
            return None
        pend["event"].synchronize()
        n, dev = p.n, self.device
        n_dup = int(pend["host"][0])
        pend.clear()
        if n_dup < max(1, min_saving * n):
            return None
        rep = torch.empty(n, dtype=torch.int32, device=dev)
        members = torch.empty(n, dtype=torch.int32, device=dev)
        offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
        counters = torch.empty(2, dtype=torch.int32, device=dev)
        ws = self._aligned_ws(self.lib.lemon_dedup_workspace_bytes(n))
        with torch.cuda.device(dev):
            self.ctx.check(self.lib.lemon_dedup_build(self.ctx.handle, _ptr(p.f32), n, p.d, _ptr(ws), _ptr(rep), _ptr(members),
                                                      _ptr(offsets), _ptr(counters), _stream()), "lemon_dedup_build")
        host = counters.cpu()                                  # second round trip, only for databases with duplicates
        n_u, collision = int(host[0]), int(host[1])
        if collision != 0 or n_u > (1.0 - min_saving) * n:
            return None
        rep = rep[:n_u]

        def gather(src, cols, dtype):
            dst = torch.empty((n_u, cols), dtype=dtype, device=dev)
            with torch.cuda.device(dev):
                self.ctx.check(self.lib.lemon_gather_rows(self.ctx.handle, _ptr(src), _ptr(rep), None, n_u,
                                                          cols * src.element_size(), _ptr(dst), _stream()), "lemon_gather_rows")
            return dst
        uniq = Prepared(gather(p.f32, p.d, torch.float32), None if p.f16 is None else gather(p.f16, p.d16, torch.float16),
                        gather(p.row_stats, 4, torch.float32), p.stats_max, n_u, p.d, p.d16)
        return Dedup(uniq, offsets[: n_u + 1], members, n_u)

    def find_duplicates(self, p: Prepared, min_saving: float = 0.1) -> "Dedup | None":
        return self.dedup_finish(p, self.dedup_start(p), min_saving)

    def prepare_db(self, x, normalize: bool = True, defer_dedup: bool = False) -> Prepared:
        """K0 + duplicate detection for one database matrix (run_lemon.py:163-164,175-176).  With defer_dedup the
        grouping is only queued (``finish_db`` completes it), so that several matrices share one host round trip."""
        p = self.prepare(x, normalize, self.knn_mode != "exact")
        if self.dedup:
            p._pending = self.dedup_start(p)
            if not defer_dedup:
                self.finish_db(p)
        return p

    def finish_db(self, p: Prepared) -> Prepared:
        pend = getattr(p, "_pending", None)
        if pend is not None:
            p._pending = None
            p.dedup = self.dedup_finish(p, pend)
        return p

    def rowwise_dist(self, a: torch.Tensor, b: torch.Tensor, metric: int) -> torch.Tensor:
        out = torch.empty(a.shape[0], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_rowwise_dist(self.ctx.handle, _ptr(a), _ptr(b), _ptr(out), a.shape[0],
                                                       a.shape[1], metric, _stream()), "lemon_rowwise_dist")
        return out

    # ------------------------------------------------------------ kNN (a4,a5)
    def tc_eligible(self, q: Prepared, db: Prepared) -> bool:
        return (q.f16 is not None and db.f16 is not None and q.d16 == db.d16 and db.d16 <= MAX_D_TC
                and db.n >= 1 and q.n >= 1)

    def knn_exact(self, q: Prepared, db: Prepared, kp: int, metric: int, top=None, rows=None, n_rows=None):
        if top is None:
            top = (torch.empty((q.n, kp), dtype=torch.float32, device=self.device),
                   torch.empty((q.n, kp), dtype=torch.int32, device=self.device))
        max_rows = q.n
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_knn_exact(self.ctx.handle, _ptr(q.f32), _ptr(db.f32), _ptr(rows), _ptr(n_rows),
                                                    max_rows, q.n, db.n, db.d, kp, metric, _ptr(top[0]), _ptr(top[1]),
                                                    _stream()), "lemon_knn_exact")
        return top

    def knn_candidates(self, q: Prepared, db: Prepared, nseg: int | None = None, cta_group: int | None = None,
                       keep: int = 0):
        """K1.  Returns (cand_keys int64[nq_pad, nlist, 256] (uint64 bit patterns), cand_cnt int32[nq_pad, nlist],
        cand_theta fp32[nq_pad, nlist], nseg) with nlist = 2*nseg; see include/lemon_b200.h."""
        cg = self.cta_group if cta_group is None else cta_group
        if nseg is None:
            nseg = plan_segments(q.n, db.n, self.num_sms, cg if cg else 2, db.d16)
        nlist = 2 * nseg
        nq_pad = -(-q.n // 256) * 256
        # K1 addresses a row's 8 KB key list with a 32-bit pointer bump: the array must be 8 KB-aligned
        raw = torch.empty(nq_pad * nlist * LIST_CAP + LIST_CAP, dtype=torch.int64, device=self.device)
        skip = (-raw.data_ptr() % (LIST_CAP * 8)) // 8
        cand_keys = raw[skip: skip + nq_pad * nlist * LIST_CAP].view(nq_pad, nlist, LIST_CAP)
        cand_cnt = torch.empty((nq_pad, nlist), dtype=torch.int32, device=self.device)
        cand_theta = torch.empty((nq_pad, nlist), dtype=torch.float32, device=self.device)
        ev = None
        if self.k1_events is not None:      # bench.py: per-launch device time of the dominant kernel
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_knn_candidates(self.ctx.handle, _ptr(q.f16), _ptr(db.f16), q.n, db.n, db.d16,
                                                         nseg, cg, keep, _ptr(cand_keys), _ptr(cand_cnt), _ptr(cand_theta),
                                                         _stream()), "lemon_knn_candidates")
        if ev is not None:
            ev[1].record()
            self.k1_events.append((ev[0], ev[1], 2.0 * q.n * db.n * db.d))
        return cand_keys, cand_cnt, cand_theta, nseg

    def rerank(self, q: Prepared, db: Prepared, cand, kp: int, metric: int, use_bound: bool = True, out=None,
               out_rows=None, acc_coef: float | None = None):
        """K2a on the output of knn_candidates (`cand` = its first three return values).  out_rows (int32 [q.n]):
        query row r belongs to row out_rows[r] of `out` (second pass on a gathered subset)."""
        cand_keys, cand_cnt, cand_theta = cand
        if out is None:
            out = (torch.empty((q.n, kp), dtype=torch.float32, device=self.device),
                   torch.empty((q.n, kp), dtype=torch.int32, device=self.device))
        top_val, top_idx = out
        uncert = torch.empty(max(q.n, 1), dtype=torch.int32, device=self.device)
        n_unc = torch.empty(1, dtype=torch.int32, device=self.device)           # zeroed by the library call
        acc = acc_eps_coef(db.d16, db.d) if acc_coef is None else acc_coef
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_rerank(
                self.ctx.handle, _ptr(q.f32), _ptr(db.f32), _ptr(cand_keys), _ptr(cand_cnt), _ptr(cand_theta),
                _ptr(q.row_stats) if use_bound else None, _ptr(db.stats_max) if use_bound else None,
                C.c_float(acc), q.n, db.n, db.d, cand_cnt.shape[1], kp, metric, _ptr(out_rows), _ptr(top_val), _ptr(top_idx),
                _ptr(uncert), _ptr(n_unc), _stream()), "lemon_rerank")
        return top_val, top_idx, uncert, n_unc

    # ------------------------------------------------ second tensor-core pass (split precision)
    def split_operands(self, p: Prepared, role: int) -> Prepared:
        """[hi | lo | hi] (role 0, queries) / [lo | hi | hi] (role 1, database) fp16 operands of 3*d16 columns and the
        statistics of the residual (include/lemon_b200.h: lemon_split_cast)."""
        d16 = p.d16
        out16 = torch.empty((p.n, 3 * d16), dtype=torch.float16, device=self.device)
        row_stats = torch.empty((p.n, 4), dtype=torch.float32, device=self.device)
        stats_max = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_split_cast(self.ctx.handle, _ptr(p.f32), _ptr(out16), _ptr(row_stats), _ptr(stats_max),
                                                     p.n, p.d, d16, role, _stream()), "lemon_split_cast")
        return Prepared(p.f32, out16, row_stats, stats_max, p.n, p.d, 3 * d16)

    def second_pass(self, q: Prepared, db: Prepared, rows: torch.Tensor, kp: int, metric: int, out):
        """Rows the first pass could not certify (their fp16 rounding bound ~6e-4 exceeds the gap below the kp-th
        neighbour) are searched again with split-precision operands (three fp16 products per pair, bound ~1e-4 at
        d = 768) and keep = 64 candidates per list, before anything is sent to the fp32 brute-force kernel (two orders
        of magnitude slower per row).  `rows`: int32 ids of the rows of q; their lists in `out` are overwritten.
        Returns (uncert, n_unc) of the rows that are still uncertified (ids of rows of q)."""
        n2 = rows.numel()
        dev = self.device

        def gather(src, cols, dtype):
            dst = torch.empty((n2, cols), dtype=dtype, device=dev)
            with torch.cuda.device(dev):
                self.ctx.check(self.lib.lemon_gather_rows(self.ctx.handle, _ptr(src), _ptr(rows), None, n2,
                                                          cols * src.element_size(), _ptr(dst), _stream()), "lemon_gather_rows")
            return dst
        q2 = Prepared(gather(q.f32, q.d, torch.float32), None, None, None, n2, q.d, q.d16)
        q2s = self.split_operands(q2, 0)
        dbs = getattr(db, "_split", None)
        if dbs is None:
            dbs = db._split = self.split_operands(db, 1)         # built on first use, kept with the staged database
        *cand, _ = self.knn_candidates(q2s, dbs, keep=MAX_KP)
        _, _, uncert, n_unc = self.rerank(q2s, dbs, cand, kp, metric, out=out, out_rows=rows,
                                          acc_coef=acc_eps_coef_split(db.d16, db.d))
        return uncert, n_unc

    def knn(self, q: Prepared, db: Prepared, kp: int, metric: int, mode: str | None = None):
        """Exact top-kp lists [nq,kp] (fp32 values, int32 DB rows), total order (best value, lower index).
        'tc': tensor-core candidates -> fp32 re-rank -> GPU exact fallback for uncertified rows."""
        mode = mode or self.knn_mode
        if kp > MAX_KP:
            raise ValueError(f"k (+1) = {kp} exceeds {MAX_KP}")
        if db.dedup is not None:
            # many identical DB rows: search the unique rows, then expand every hit to its group's members
            dd = db.dedup
            uv, ui = self.knn(q, dd.uniq, kp, metric, mode)
            info = dict(self.last_info, n_unique=dd.n_unique)
            tv = torch.empty((q.n, kp), dtype=torch.float32, device=self.device)
            ti = torch.empty((q.n, kp), dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                self.ctx.check(self.lib.lemon_expand_groups(self.ctx.handle, _ptr(uv), _ptr(ui), _ptr(dd.offsets),
                                                            _ptr(dd.members), q.n, kp, metric, _ptr(tv), _ptr(ti),
                                                            _stream()), "lemon_expand_groups")
            self.last_info = info
            return tv, ti
        use_tc = mode == "tc" or (mode == "auto" and self.tc_eligible(q, db) and db.n >= 2048)
        if mode == "tc" and not self.tc_eligible(q, db):
            raise _lib.LemonError("tensor-core path not eligible for this shape (padded d must be <= 768)")
        if not use_tc:
            tv, ti = self.knn_exact(q, db, kp, metric)
            self.last_info = {"path": "exact"}
            return tv, ti
        top_val = torch.empty((q.n, kp), dtype=torch.float32, device=self.device)
        top_idx = torch.empty((q.n, kp), dtype=torch.int32, device=self.device)
        cg = self.cta_group if self.cta_group else 2
        parts = plan_parts(q.n, db.n, self.num_sms, cg, d16=db.d16)
        pending, nsegs = [], []
        for r0, r1, ns in parts:
            qs = q if (r0 == 0 and r1 == q.n) else _slice_prepared(q, r0, r1)
            *cand, nseg = self.knn_candidates(qs, db, nseg=ns, keep=tc_keep(kp))
            tv, ti = top_val[r0:r1], top_idx[r0:r1]
            _, _, uncert, n_unc = self.rerank(qs, db, cand, kp, metric, out=(tv, ti))
            host = self._pinned_slot()
            host.copy_(n_unc, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((qs, tv, ti, uncert, n_unc, host, ev))
            nsegs.append(nseg)
        # Rows without a certificate.  Their count is read back once per launch (all launches are queued by now, so
        # the host only waits for work that has to finish anyway): it sizes the second tensor-core pass.
        first, second = [], []
        for qs, tv, ti, uncert, n_unc, host, ev in pending:
            ev.synchronize()
            n1 = int(host[0])
            first.append(n1)
            if n1 == 0:
                second.append(0)
                continue
            if self.second_pass_enabled and db.d16 * 3 <= MAX_D_TC * 3:
                uncert, n_unc = self.second_pass(qs, db, uncert[:n1], kp, metric, (tv, ti))
                second.append(n_unc)
            else:
                second.append(n1)
            # what is still uncertified: exact fp32 brute force on the GPU (row list and count stay on the device)
            self.knn_exact(qs, db, kp, metric, top=(tv, ti), rows=uncert, n_rows=n_unc)
        self.last_info = {"path": "tc", "nseg": nsegs[0] if len(nsegs) == 1 else nsegs,
                          "n_uncertified_first_pass": int(sum(first)),
                          "n_uncertified": second}      # after the second pass: ints (0) or device counters, one per launch
        return top_val, top_idx

    # ------------------------------------------------- run_lemon.py:163-176
    def set_database(self, img_db, txt_db, dist_type: str = "cosine", normalize: bool = True,
                     text_label_ids_db=None):
        metric = METRIC[dist_type]
        xdb = self.prepare_db(img_db, normalize, defer_dedup=True)
        ydb = self.prepare_db(txt_db, normalize, defer_dedup=True)
        self.finish_db(xdb)           # both duplicate counts are queued by now: one host round trip reads them
        self.finish_db(ydb)
        assert xdb.n == ydb.n and xdb.d == ydb.d
        self.db = {"x": xdb, "y": ydb, "metric": metric, "normalize": normalize,
                   "dists_tr": self.rowwise_dist(ydb.f32, xdb.f32, metric),
                   "labels": _to_dev(text_label_ids_db, self.device, torch.int32)}
        return self

    # ------------------------------- run_lemon.py:235-307 + utils.py:47-82
    def score(self, img_q, txt_q, *, k: int, query_in_db=None, hparams: dict | None = None,
              text_label_ids_q=None, return_records: bool = True, queries_are_db: bool = False,
              query_rows: tuple[int, int] | None = None, class_text_emb=None, noisy_label=None) -> dict:
        """Scores every query pair against the database set by ``set_database``.

        query_in_db: None -> val/test rule (search k).  int64[N] (DB row of the query, -1 if
        absent) -> train rule: search k+1, drop rank 0 if present else the last
        (run_lemon.py:257-263).  queries_are_db=True reuses the staged DB operands as queries
        (N == M, the BASELINE 'N x N' configs) instead of staging them twice; query_rows=(r0, r1)
        does the same for one rank's block of rows."""
        assert self.db is not None, "call set_database first"
        db = self.db
        metric = db["metric"]
        if query_rows is not None:      # queries = DB rows [r0, r1) (row-sharded multi-GPU): views, no re-staging
            xq, yq = (_slice_prepared(db[s], *query_rows) for s in ("x", "y"))
        elif queries_are_db:
            xq, yq = db["x"], db["y"]
        else:
            need16 = self.knn_mode != "exact"
            xq = self.prepare(img_q, db["normalize"], need16)
            yq = self.prepare(txt_q, db["normalize"], need16)
        nq = xq.n
        train = query_in_db is not None
        kp = k + 1 if train else k
        qid = _to_dev(query_in_db, self.device, torch.int64)
        lab_q = _to_dev(text_label_ids_q, self.device, torch.int32)
        lab_db = db["labels"] if lab_q is not None else None
        if lab_q is not None and lab_db is None:
            raise ValueError("text_label_ids_q given but the database has no text_label_ids_db")
        cls_emb = n_class = lab_noisy = None
        if class_text_emb is not None:     # --normalize_d1: class prompts are normalised like every other embedding
            cls_p = self.prepare(class_text_emb, db["normalize"], need_f16=False)
            cls_emb, n_class = cls_p.f32, cls_p.n
            lab_noisy = _to_dev(noisy_label, self.device, torch.int32)
            assert lab_noisy is not None and lab_noisy.numel() == nq and cls_p.d == db["x"].d
        topn = self.knn(xq, db["x"], kp, metric)
        info_n = self.last_info
        topm = self.knn(yq, db["y"], kp, metric)
        info_m = self.last_info
        out = self.emit(xq, yq, db["x"], db["y"], db["dists_tr"], topn, topm, k=k, kp=kp, metric=metric, qid=qid,
                        lab_q=lab_q, lab_db=lab_db, cls_emb=cls_emb, lab_noisy=lab_noisy, n_class=n_class,
                        hparams=hparams, return_records=return_records)
        self.last_info = {"img": info_n, "txt": info_m}
        return out

    def alloc_outputs(self, nq: int, k: int, hparams, return_records: bool = True, index_dtype=torch.int64) -> dict:
        dev = self.device
        out = {"d_1": torch.empty(nq, dtype=torch.float32, device=dev)}
        if return_records:
            for c in ("D_n", "dists_n", "dists_tr_n", "D_m", "dists_m", "dists_tr_m"):
                out[c] = torch.empty((nq, k), dtype=torch.float32, device=dev)
            out["I_n"] = torch.empty((nq, k), dtype=index_dtype, device=dev)
            out["I_m"] = torch.empty((nq, k), dtype=index_dtype, device=dev)
        if hparams is not None:
            for c in ("s_n", "s_m", "score"):
                out[c] = torch.empty(nq, dtype=torch.float64, device=dev)
        return out

    def emit(self, xq: Prepared, yq: Prepared, xdb: Prepared, ydb: Prepared, dists_tr, topn, topm, *, k: int, kp: int,
             metric: int, qid=None, lab_q=None, lab_db=None, cls_emb=None, lab_noisy=None, n_class=None,
             hparams: dict | None = None, return_records: bool = True, index_dtype=torch.int64, out: dict | None = None,
             rows: tuple[int, int] | None = None, sides: int = 3) -> dict:
        """K2b: per-sample records + score from exact top lists (run_lemon.py:250-307, utils.py:63-77).
        `out` = tensors from alloc_outputs to write into; rows=(a, b) processes query rows [a, b) only (callers that
        stream the results to the host launch it part by part).  sides: 1 = image-neighbour side only, 2 = text-neighbour
        side + d_1 + score (after a sides = 1 call), 3 = both (include/lemon_b200.h)."""
        nq = xq.n
        if out is None:
            out = self.alloc_outputs(nq, k, hparams, return_records, index_dtype)
        a, b = (0, nq) if rows is None else rows
        if b <= a:
            return out
        assert index_dtype in (torch.int64, torch.int32)
        hp_arr = (C.c_double * 6)(*[float(hparams[key]) for key in HP_KEYS]) if hparams is not None else None
        sl = lambda t: None if t is None else t[a:b]
        tl = lambda top, i: None if top is None else top[i][a:b]       # a side's top lists may be absent when `sides` skips it
        g = lambda name: _ptr(sl(out.get(name)))
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_score(
                self.ctx.handle, _ptr(xq.f32[a:b]), _ptr(yq.f32[a:b]), _ptr(xdb.f32), _ptr(ydb.f32), _ptr(dists_tr),
                _ptr(tl(topn, 0)), _ptr(tl(topn, 1)), _ptr(tl(topm, 0)), _ptr(tl(topm, 1)), _ptr(sl(qid)), _ptr(sl(lab_q)),
                _ptr(lab_db), _ptr(cls_emb), _ptr(sl(lab_noisy)), int(n_class or 0), b - a, xdb.n, xdb.d, k, kp, metric, hp_arr,
                g("d_1"), g("D_n"), g("dists_n"), g("dists_tr_n"), g("D_m"), g("dists_m"), g("dists_tr_m"), g("I_n"), g("I_m"),
                64 if index_dtype == torch.int64 else 32, int(sides), g("s_n"), g("s_m"), g("score"), _stream()), "lemon_score")
        return out

    def combine_scores(self, rec: dict, hparams: dict) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """lib/metrics/utils.py:63-77 on device-resident [N,k] columns."""
        dev = self.device
        cols = {c: _to_dev(rec[c], dev, torch.float32) for c in
                ("D_n", "dists_tr_n", "dists_n", "D_m", "dists_tr_m", "dists_m")}
        d1 = _to_dev(rec["d_1"], dev, torch.float64)
        n, k = cols["D_n"].shape
        sn, sm, sc = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
        hp_arr = (C.c_double * 6)(*[float(hparams[key]) for key in HP_KEYS])
        with torch.cuda.device(dev):
            self.ctx.check(self.lib.lemon_combine_scores(
                self.ctx.handle, _ptr(cols["D_n"]), _ptr(cols["dists_tr_n"]), _ptr(cols["dists_n"]), _ptr(cols["D_m"]),
                _ptr(cols["dists_tr_m"]), _ptr(cols["dists_m"]), _ptr(d1), n, k, hp_arr, _ptr(sn), _ptr(sm), _ptr(sc),
                _stream()), "lemon_combine_scores")
        return sc, sn, sm


    def keep_lowest(self, score, n_keep: int):
        """CC3M filtering step of train_clip_from_scratch.py:110-113 on the device: row ids (int64) of the `n_keep`
        lowest scores in ascending score order (ties: lower row id first) and their scores."""
        s = _to_dev(score, self.device, torch.float64).reshape(-1)
        n = s.numel()
        n_keep = int(min(max(n_keep, 0), n))
        idx = torch.empty(n_keep, dtype=torch.int64, device=self.device)
        val = torch.empty(n_keep, dtype=torch.float64, device=self.device)
        if n_keep == 0:
            return idx, val
        ws = torch.empty(int(self.lib.lemon_keep_lowest_workspace_bytes(n)) + 256, dtype=torch.uint8, device=self.device)
        ws = ws[(-ws.data_ptr()) % 256:]
        with torch.cuda.device(self.device):
            self.ctx.check(self.lib.lemon_keep_lowest(self.ctx.handle, _ptr(s), n, n_keep, _ptr(ws), _ptr(idx), _ptr(val),
                                                      _stream()), "lemon_keep_lowest")
        return idx, val


_default_scorers: dict = {}


def get_scorer(device=None, knn_mode: str = "auto") -> LemonScorer:
    key = (str(device), knn_mode)
    if key not in _default_scorers:
        _default_scorers[key] = LemonScorer(device, knn_mode)
    return _default_scorers[key]


def score_pairs(img_q, txt_q, img_db=None, txt_db=None, *, k: int, dist_type: str = "cosine",
                query_in_db=None, hparams: dict | None = None, text_label_ids_q=None, text_label_ids_db=None,
                normalize: bool = True, return_records: bool = True, to_host: bool = False, device=None,
                knn_mode: str = "auto", class_text_emb=None, noisy_label=None) -> dict:
    """Fused replacement of run_lemon.py:163-314 + :406-407 in one call (SURVEY.md §8b seam 3).

    img_db/txt_db None means DB == queries (N == M).  class_text_emb [C,d] + noisy_label [N] select the
    --normalize_d1 variant of d_1 (run_lemon.py:244-248).  Returns the df columns of
    run_lemon.py:291-307 as tensors: d_1 [N]; D_n, dists_n, dists_tr_n, D_m, dists_m, dists_tr_m
    [N,k] fp32; I_n, I_m [N,k] int64; and, with hparams, s_n, s_m, score [N] float64."""
    sc = get_scorer(device, knn_mode)
    same = img_db is None
    sc.set_database(img_q if same else img_db, txt_q if same else txt_db, dist_type, normalize, text_label_ids_db
                    if not same or text_label_ids_db is not None else text_label_ids_q)
    out = sc.score(img_q, txt_q, k=k, query_in_db=query_in_db, hparams=hparams, text_label_ids_q=text_label_ids_q,
                   return_records=return_records, queries_are_db=same, class_text_emb=class_text_emb,
                   noisy_label=noisy_label)
    if to_host:
        out = {name: t.cpu().numpy() for name, t in out.items()}
    return out
