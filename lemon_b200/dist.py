"""Row-sharded multi-GPU driver: one process per GPU, every rank owns a
contiguous block of pairs (its queries AND its slice of the database), the database is
replicated with one all-gather per modality over NCCL/NVLink, and there is no other
collective — every output row is produced by the rank that owns it.

The collective plumbing is backend-agnostic (tested on CPU with gloo, world_size 2); the
scoring itself needs the CUDA library.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


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


def score_pairs_sharded(img_local, txt_local, n_total: int, *, k: int, dist_type: str = "cosine",
                        hparams=None, normalize: bool = True, return_records: bool = True, scorer=None,
                        group=None, text_label_ids_local=None) -> dict:
    """Every rank passes its padded shard [per, d] of both modalities.  The DB is all pairs (train-split
    self-exclusion on, query_in_db = own global row ids).  Returns this rank's rows of every output (see
    lemon_b200.score_pairs) plus 'rows' = (r0, r1).

    Shards may be device tensors or (pinned) HOST tensors.  With host tensors both host->device copies are
    issued up front on a copy stream and the whole image side (all-gather, K0, K1, K2a) runs while the text
    shard is still in flight; the text side starts when its copy has landed."""
    from .scoring import METRIC, _slice_prepared, _to_dev
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    r0, r1, per = shard_bounds(n_total, world, rank)
    assert img_local.shape[0] == per and txt_local.shape[0] == per, "shards must be padded to equal length"
    if scorer is None:
        from .scoring import get_scorer
        scorer = get_scorer(img_local.device.index if img_local.is_cuda else None)
    if not hasattr(scorer, "prepare_db"):        # stand-in scorers (CPU tests): plain sequential path
        img_db = allgather_rows(img_local, n_total, group)
        txt_db = allgather_rows(txt_local, n_total, group)
        lab_db = None
        if text_label_ids_local is not None:
            lab_db = allgather_rows(text_label_ids_local.view(-1, 1), n_total, group).view(-1)
        scorer.set_database(img_db, txt_db, dist_type, normalize, lab_db)
        qid = torch.arange(r0, r1, dtype=torch.int64, device=img_local.device)
        out = scorer.score(None, None, k=k, query_in_db=qid, hparams=hparams, return_records=return_records,
                           query_rows=(r0, r1), text_label_ids_q=lab_db[r0:r1] if lab_db is not None else None)
        out["rows"] = (r0, r1)
        return out

    dev = scorer.device
    main = torch.cuda.current_stream(dev)
    e_txt = None
    if not img_local.is_cuda:
        cs = _copy_streams.get(dev.index)
        if cs is None:
            cs = _copy_streams[dev.index] = torch.cuda.Stream(dev)
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            img_local = img_local.to(dev, non_blocking=True)
            e_img = torch.cuda.Event()
            e_img.record(cs)
            txt_local = txt_local.to(dev, non_blocking=True)
            e_txt = torch.cuda.Event()
            e_txt.record(cs)
        img_local.record_stream(main)
        txt_local.record_stream(main)
        main.wait_event(e_img)
    metric = METRIC[dist_type]
    kp = k + 1
    # ---- image side (run_lemon.py:164,168/172,176,235) while the text shard may still be copying
    xdb = scorer.prepare_db(allgather_rows(img_local, n_total, group), normalize)
    xq = _slice_prepared(xdb, r0, r1)
    ydb = None
    if e_txt is None:
        # device-resident shards: stage the text side too before the long kernels are queued (duplicate
        # detection reads two scalars back; doing it now keeps the host ahead of the GPU)
        ydb = scorer.prepare_db(allgather_rows(txt_local, n_total, group), normalize)
    topn = scorer.knn(xq, xdb, kp, metric)
    info_n = scorer.last_info
    # ---- text side
    if ydb is None:
        main.wait_event(e_txt)
        ydb = scorer.prepare_db(allgather_rows(txt_local, n_total, group), normalize)
    yq = _slice_prepared(ydb, r0, r1)
    dtr = scorer.rowwise_dist(ydb.f32, xdb.f32, metric)
    topm = scorer.knn(yq, ydb, kp, metric)
    info_m = scorer.last_info
    lab_db = lab_q = None
    if text_label_ids_local is not None:
        lab_local = _to_dev(text_label_ids_local, dev, torch.int32)
        lab_db = allgather_rows(lab_local.view(-1, 1), n_total, group).view(-1).contiguous()
        lab_q = lab_db[r0:r1]
    qid = torch.arange(r0, r1, dtype=torch.int64, device=dev)
    out = scorer.emit(xq, yq, xdb, ydb, dtr, topn, topm, k=k, kp=kp, metric=metric, qid=qid, lab_q=lab_q, lab_db=lab_db,
                      hparams=hparams, return_records=return_records)
    scorer.last_info = {"img": info_n, "txt": info_m}
    out["rows"] = (r0, r1)
    return out
