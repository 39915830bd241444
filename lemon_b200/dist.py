"""Row-sharded multi-GPU driver: one process per GPU, every rank owns a
contiguous block of pairs (its queries AND its slice of the database), the database is
replicated with one all-gather per modality over NCCL/NVLink, and there is no other
collective — every output row is produced by the rank that owns it.  (With the discrete text
metric the int32 label ids travel as four extra columns of the TEXT all-gather; K0 reads the
embedding columns of the gathered matrix in place through its row stride.)

The collective plumbing is backend-agnostic (tested on CPU with gloo, world_size 2, with a
scorer that implements the same staged interface on the CPU oracle); the scoring itself needs
the CUDA library.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

LABEL_COLS = 4      # label ids ride in 4 extra fp32 columns (keeps rows 16 B aligned); column 0 holds the int32 bits


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int, int]:
    """Rows [r0, r1) owned by `rank` and the padded shard length (equal on all ranks)."""
    per = -(-n // world)
    r0 = min(n, rank * per)
    r1 = min(n, r0 + per)
    return r0, r1, per


def allgather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local: [per, d] (rows past the rank's valid count are padding).  Returns the replicated
    [n_total, d] matrix; padding only ever sits at the tail, so it is sliced off."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n_total]
    per, d = local.shape
    out = torch.empty((world * per, d), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "gloo":
        chunks = list(out.view(world, per, d).unbind(0))
        dist.all_gather(chunks, local.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:n_total]


_copy_streams: dict = {}
_qid_cache: dict = {}


def _global_row_ids(scorer, r0: int, r1: int) -> torch.Tensor:
    """query_in_db of this rank's rows (their own global row ids), built once per (device, row range)."""
    key = (str(scorer.device), r0, r1)
    t = _qid_cache.get(key)
    if t is None:
        _qid_cache.clear()
        t = _qid_cache[key] = torch.arange(r0, r1, dtype=torch.int64, device=scorer.device)
    return t


def _side_stream(dev, which: str):
    key = (dev.index, which)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(dev)
    return _copy_streams[key]


def score_pairs_sharded(img_local, txt_local, n_total: int, *, k: int, dist_type: str = "cosine",
                        hparams=None, normalize: bool = True, return_records: bool = True, scorer=None,
                        group=None, text_label_ids_local=None, host_out: dict | None = None,
                        index_dtype=torch.int64, d2h_parts: int = 4, h2d_chunks: int = 4) -> dict:
    """Every rank passes its padded shard [per, d] of both modalities.  The DB is all pairs (train-split
    self-exclusion on, query_in_db = own global row ids).  Returns this rank's rows of every output (see
    lemon_b200.score_pairs) plus 'rows' = (r0, r1).

    Shards may be device tensors or (pinned) HOST tensors.  With host tensors both host->device copies are
    issued up front on a copy stream in `h2d_chunks` pieces; every piece is all-gathered and normalised on a side
    stream as soon as it has landed (lemon_b200.handoff.ShardStager), so one piece of the image copy is exposed and
    the text side is staged behind the image-side kernels.

    host_out: dict of pinned host tensors (one per output column, at least r1-r0 rows; missing ones are created
    and added).  The records are then produced in `d2h_parts` row parts and each part's device->host copy runs on
    a copy stream while the next part is being computed; the call returns host tensors once all copies have
    landed.  index_dtype=torch.int32 halves the bytes of I_n / I_m (faiss's int64 is the default)."""
    from .scoring import METRIC
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    r0, r1, per = shard_bounds(n_total, world, rank)
    assert img_local.shape[0] == per and txt_local.shape[0] == per, "shards must be padded to equal length"
    if scorer is None:
        from .scoring import get_scorer
        scorer = get_scorer(img_local.device.index if img_local.is_cuda else None)
    dev = scorer.device
    on_gpu = dev.type == "cuda"
    d = img_local.shape[1]
    with_labels = text_label_ids_local is not None
    main = torch.cuda.current_stream(dev) if on_gpu else None
    metric = METRIC[dist_type]
    if on_gpu and not img_local.is_cuda and d % 4 == 0:
        # ---- host shards: the copies are cut into chunks on a copy stream; every IMAGE chunk is all-gathered and
        # normalised on a side stream as soon as it has landed, so only one chunk of the image copy is exposed; the text
        # copy runs under the image-side kernels and the text database is staged behind them, on the main stream's order
        from .handoff import ShardStager
        cs = _side_stream(dev, "h2d")
        cs.wait_stream(main)
        st_img = ShardStager(scorer, n_total, (r0, r1, per), d, normalize, group, h2d_chunks,
                             side=_side_stream(dev, "stage_img"))
        st_txt = ShardStager(scorer, n_total, (r0, r1, per), d, normalize, group, h2d_chunks, LABEL_COLS if with_labels else 0,
                             side=_side_stream(dev, "stage_txt"))
        step = max(1, -(-per // max(1, h2d_chunks)))
        e_img, e_txt = torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(cs):
            for a in range(0, per, step):
                st_img.append(img_local[a:a + step])
            e_img.record(cs)
            if with_labels:
                lab = torch.as_tensor(text_label_ids_local).to(device=dev, dtype=torch.int32, non_blocking=True)
                st_txt.shard[:, d].copy_(lab.view(torch.float32))
            for a in range(0, per, step):
                st_txt.append(txt_local[a:a + step], cols=slice(0, d), stage=False)
            e_txt.record(cs)
        xdb = st_img.finish(after=e_img)
        xdb._pending = scorer.dedup_start(xdb) if getattr(scorer, "dedup", False) else None
        scorer.finish_db(xdb)

        def finish_text():
            ydb = st_txt.finish(after=e_txt)
            ydb._pending = scorer.dedup_start(ydb) if getattr(scorer, "dedup", False) else None
            return scorer.finish_db(ydb), st_txt.labels[:n_total] if with_labels else None
        return score_staged(scorer, xdb, None, (r0, r1, per), k=k, metric=metric, hparams=hparams, return_records=return_records,
                            lab_db=None, host_out=host_out, index_dtype=index_dtype, d2h_parts=d2h_parts, late_text=finish_text)

    # ---- device-resident shards
    if with_labels:     # the text shard as it is all-gathered: [per, d + LABEL_COLS] with the label ids
        txt_wide = torch.zeros((per, d + LABEL_COLS), dtype=torch.float32, device=dev)
        txt_wide[:, :d].copy_(txt_local, non_blocking=True)
        lab = torch.as_tensor(text_label_ids_local).to(device=dev, dtype=torch.int32, non_blocking=True)
        txt_wide[:, d].copy_(lab.view(torch.float32))
    else:
        txt_wide = txt_local.to(dev, non_blocking=True)
    img_local = img_local.to(dev, non_blocking=True)
    # image and text databases are staged before the long kernels are queued, so that the duplicate detection of both
    # matrices costs ONE host round trip (it reads their counters back)
    xdb = scorer.prepare_db(allgather_rows(img_local, n_total, group), normalize, defer_dedup=True)
    txt_all = allgather_rows(txt_wide, n_total, group)
    ydb = scorer.prepare_db(txt_all[:, :d] if with_labels else txt_all, normalize, defer_dedup=True)
    scorer.finish_db(xdb)
    scorer.finish_db(ydb)
    lab_db = txt_all[:, d].contiguous().view(torch.int32) if with_labels else None
    return score_staged(scorer, xdb, ydb, (r0, r1, per), k=k, metric=metric, hparams=hparams, return_records=return_records,
                        lab_db=lab_db, host_out=host_out, index_dtype=index_dtype, d2h_parts=d2h_parts)


def score_staged(scorer, xdb, ydb, bounds, *, k: int, metric: int, hparams=None, return_records: bool = True, lab_db=None,
                 host_out: dict | None = None, index_dtype=torch.int64, d2h_parts: int = 4, late_text=None) -> dict:
    """Second half of the sharded path on STAGED databases (K0 done, duplicate grouping finished): kNN of this rank's
    rows against both replicated databases, dists_tr, records + score.  `late_text` (host-input path) stages the text
    database after the image-side kNN has been queued."""
    from .scoring import _slice_prepared
    r0, r1, per = bounds
    dev = scorer.device
    on_gpu = dev.type == "cuda"
    main = torch.cuda.current_stream(dev) if on_gpu else None
    kp = k + 1
    xq = _slice_prepared(xdb, r0, r1)
    topn = scorer.knn(xq, xdb, kp, metric)
    info_n = scorer.last_info
    # ---- text side
    if ydb is None:
        ydb, lab_db = late_text()
    yq = _slice_prepared(ydb, r0, r1)
    dtr = scorer.rowwise_dist(ydb.f32, xdb.f32, metric)
    lab_q = lab_db[r0:r1] if lab_db is not None else None
    qid = _global_row_ids(scorer, r0, r1)
    nq = r1 - r0
    common = dict(k=k, kp=kp, metric=metric, qid=qid, lab_q=lab_q, lab_db=lab_db, hparams=hparams,
                  return_records=return_records, index_dtype=index_dtype)
    stream_out = host_out is not None and on_gpu
    early = ()
    if stream_out:
        out_dev = scorer.alloc_outputs(nq, k, hparams, return_records, index_dtype)
        ds = _side_stream(dev, "d2h")
        for name, t in out_dev.items():
            h = host_out.get(name)
            if h is None or h.shape[0] < nq or h.dtype != t.dtype or h.shape[1:] != t.shape[1:]:
                host_out[name] = torch.empty((per,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
        if return_records:
            # the image-neighbour half of the records only needs the image-side search: it is computed now and shipped
            # to the host while the text-side search runs
            scorer.emit(xq, yq, xdb, ydb, dtr, topn, None, out=out_dev, sides=1, **common)
            early = tuple(c for c in ("D_n", "dists_n", "dists_tr_n", "I_n", "s_n") if c in out_dev)
            ev = torch.cuda.Event()
            ev.record(main)
            ds.wait_event(ev)
            with torch.cuda.stream(ds):
                for name in early:
                    host_out[name][:nq].copy_(out_dev[name], non_blocking=True)
    topm = scorer.knn(yq, ydb, kp, metric)
    info_m = scorer.last_info
    if not stream_out:
        out = scorer.emit(xq, yq, xdb, ydb, dtr, topn, topm, **common)
    else:
        # the rest part by part; the copy stream ships part i to the host while part i+1 is computed
        late = [name for name in out_dev if name not in early]
        parts = max(1, min(int(d2h_parts), nq))
        step = -(-nq // parts)
        for a in range(0, nq, step):
            b = min(nq, a + step)
            scorer.emit(xq, yq, xdb, ydb, dtr, topn if not early else None, topm, out=out_dev, rows=(a, b),
                        sides=2 if early else 3, **common)
            ev = torch.cuda.Event()
            ev.record(main)
            ds.wait_event(ev)
            with torch.cuda.stream(ds):
                for name in late:
                    host_out[name][a:b].copy_(out_dev[name][a:b], non_blocking=True)
        for t in out_dev.values():
            t.record_stream(ds)
        ds.synchronize()                       # the caller holds the results on the host
        out = {name: host_out[name][:nq] for name in out_dev}
    scorer.last_info = {"img": info_n, "txt": info_m}
    out["rows"] = (r0, r1)
    return out
