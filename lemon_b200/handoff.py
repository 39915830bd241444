"""On-device embedding hand-off (BASELINE config 5).

The reference embeds every batch with its CLIP model, copies the features to the CPU (``.detach().cpu()``,
run_lemon.py:158-161 and :230-233), concatenates and normalises them there (:163-164) and only then builds the
index.  Here the encoder's outputs are written straight into this rank's device shard, and the replicated database
is staged WHILE the encoder is still running: as soon as a chunk of the shard is complete it is all-gathered
(NCCL, side stream) and normalised / cast by K0 into its rows of the database operands, so when the last batch
leaves the encoder only the last chunk is left to gather.  The scoring then runs on the staged operands
(``dist.score_staged``).  The encoders are the caller's (``algorithm.encode_image`` / ``encode_text``,
lib/models/downstream_models.py:30-41): any callables that return ``[b, d]`` CUDA tensors.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import dist as ldist
from .scoring import METRIC, Prepared, _ptr, get_scorer


class ShardStager:
    """Collects one modality's rows of this rank ([per, d], filled front to back, from the device or from pinned host
    memory) and stages the replicated database operands chunk by chunk on a side stream: all-gather of the chunk (NCCL),
    then K0 (normalise, fp16 copy, statistics) into the chunk's rows of the database.  With `label_cols` the shard rows
    carry that many extra fp32 columns after the embedding (column 0 of them = the int32 label id of the discrete text
    metric); they travel in the same all-gather and K0 reads the embedding columns through its row stride."""

    def __init__(self, scorer, n_total: int, bounds, d: int, normalize: bool, group, chunks: int, label_cols: int = 0,
                 side: "torch.cuda.Stream | None" = None):
        self.sc, self.n, self.group, self.normalize = scorer, n_total, group, normalize
        self.r0, self.r1, self.per = bounds
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = scorer.device
        assert d % 4 == 0 and label_cols % 4 == 0, "row widths must be multiples of 4 floats (16 B)"
        self.d, self.d16, self.d_in = d, -(-d // 64) * 64, d + label_cols
        self.shard = torch.zeros((self.per, self.d_in), dtype=torch.float32, device=dev)
        rows = self.world * self.per                       # padded row space; the operands are sliced to n_total at the end
        self.f32 = torch.empty((rows, d), dtype=torch.float32, device=dev)
        self.f16 = torch.empty((rows, self.d16), dtype=torch.float16, device=dev)
        self.stats = torch.empty((rows, 4), dtype=torch.float32, device=dev)
        self.labels = torch.empty(rows, dtype=torch.int32, device=dev) if label_cols else None
        self.chunk = max(1, -(-self.per // max(1, chunks)))
        self.maxima = []                                   # one [4] tensor per K0 call
        self.filled = 0                                    # rows written into the shard
        self.staged = 0                                    # rows handed to the side stream
        # a persistent side stream per role: the caching allocator keeps one pool per stream, a fresh stream per call
        # would turn every gather buffer into a cudaMalloc
        self.side = side if side is not None else ldist._side_stream(dev, "stage")

    def append(self, rows: torch.Tensor, cols: "slice | None" = None, stage: bool = True):
        """rows: [b, d_in] (or [b, width of `cols`]) device or pinned host tensor; copied on the CURRENT stream.
        stage=False only copies: the rows are staged by ``finish`` (used for the text shard of the host-input path: its
        all-gather must not run beside the image-side search kernel -- a collective that waits for a peer whose SMs are
        all taken by that persistent kernel would sit on this rank's SMs for the whole search)."""
        b = rows.shape[0]
        assert self.filled + b <= self.per
        dst = self.shard[self.filled:self.filled + b]
        dst = dst if cols is None else dst[:, cols]
        assert rows.shape[1] == dst.shape[1]
        dst.copy_(rows, non_blocking=True)                 # dtype conversion (bf16 autocast -> fp32) included
        self.filled += b
        while stage and self.filled - self.staged >= self.chunk:
            self._stage(self.staged, self.staged + self.chunk)

    def _k0(self, src, row0, nrows):
        sc = self.sc
        mx = torch.empty(4, dtype=torch.float32, device=sc.device)
        self.maxima.append(mx)
        with torch.cuda.device(sc.device):
            sc.ctx.check(sc.lib.lemon_normalize_cast(
                sc.ctx.handle, _ptr(src), _ptr(self.f32[row0:row0 + nrows]), _ptr(self.f16[row0:row0 + nrows]),
                _ptr(self.stats[row0:row0 + nrows]), _ptr(mx), nrows, self.d, self.d16, self.d_in, int(bool(self.normalize)),
                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lemon_normalize_cast")

    def _stage(self, c0: int, c1: int):
        """all-gather rows [c0, c1) of every rank's shard and run K0 on them, on the side stream (which first waits for
        what the current stream has queued so far, i.e. for the copies of those rows)"""
        cur = torch.cuda.current_stream(self.sc.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            part = self.shard[c0:c1]
            if self.world == 1:
                buf = part.unsqueeze(0)
            else:
                buf = torch.empty((self.world, c1 - c0, self.d_in), dtype=torch.float32, device=self.sc.device)
                dist.all_gather_into_tensor(buf.view(-1, self.d_in), part, group=self.group)
                buf.record_stream(self.side)
            for r in range(self.world):
                self._k0(buf[r], r * self.per + c0, c1 - c0)
            if self.labels is not None:
                self.labels.view(self.world, self.per)[:, c0:c1].copy_(buf[:, :, self.d].view(torch.int32))
        self.staged = c1

    def finish(self, after=None) -> Prepared:
        """Stages what is left and makes the current stream wait for the staged operands.  after: event recorded behind
        the copies that fill the remaining rows (the host->device copy stream), needed only when rows are left."""
        assert self.filled >= self.r1 - self.r0, "fewer rows than this rank owns were appended"
        cur = torch.cuda.current_stream(self.sc.device)
        if self.staged < self.per:
            if after is not None:
                cur.wait_event(after)
            self._stage(self.staged, self.per)
        cur.wait_stream(self.side)
        smax = torch.stack(self.maxima).amax(dim=0)
        n = self.n
        return Prepared(self.f32[:n], self.f16[:n], self.stats[:n], smax, n, self.d, self.d16)


def extract_and_score(batches, encode_image, encode_text, n_total: int, *, k: int, dist_type: str = "cosine",
                      hparams=None, normalize: bool = True, return_records: bool = True, scorer=None, group=None,
                      gather_chunks: int = 4, host_out: dict | None = None, index_dtype=torch.int64) -> dict:
    """Embeds this rank's pairs and scores them without the embeddings ever leaving the device.

    batches: iterable of ``(image_input, text_input)`` covering this rank's rows ``shard_bounds(n_total, world, rank)``
    in order; ``encode_image(image_input)`` / ``encode_text(text_input)`` return ``[b, d]`` CUDA tensors (any float
    dtype).  Returns what ``dist.score_pairs_sharded`` returns for the same embeddings (bit-identical), plus
    ``'shards'`` = the rank's (image, text) feature shards on the device."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = ldist.shard_bounds(n_total, world, rank)
    if scorer is None:
        scorer = get_scorer(torch.cuda.current_device())
    st_img = st_txt = None
    with torch.no_grad():
        for image_input, text_input in batches:
            fi = encode_image(image_input)
            ft = encode_text(text_input)
            if st_img is None:
                dev = scorer.device
                st_img = ShardStager(scorer, n_total, bounds, fi.shape[1], normalize, group, gather_chunks,
                                     side=ldist._side_stream(dev, "stage_img"))
                st_txt = ShardStager(scorer, n_total, bounds, ft.shape[1], normalize, group, gather_chunks,
                                     side=ldist._side_stream(dev, "stage_txt"))
            st_img.append(fi)
            st_txt.append(ft)
    if st_img is None:
        raise ValueError("no batches")
    xdb, ydb = st_img.finish(), st_txt.finish()
    if scorer.dedup:
        xdb._pending, ydb._pending = scorer.dedup_start(xdb), scorer.dedup_start(ydb)
        scorer.finish_db(xdb)
        scorer.finish_db(ydb)
    out = ldist.score_staged(scorer, xdb, ydb, bounds, k=k, metric=METRIC[dist_type], hparams=hparams,
                             return_records=return_records, host_out=host_out, index_dtype=index_dtype)
    out["shards"] = (st_img.shard, st_txt.shard)
    return out
