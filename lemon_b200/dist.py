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
    world = dist.get_world_size(group)
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


def score_pairs_sharded(img_local, txt_local, n_total: int, *, k: int, dist_type: str = "cosine",
                        hparams=None, normalize: bool = True, return_records: bool = True, scorer=None,
                        group=None, text_label_ids_local=None) -> dict:
    """Every rank passes its padded shard [per, d] of both modalities (device tensors).  The DB is
    all pairs (train-split self-exclusion on, query_in_db = own global row ids).  Returns this
    rank's rows of every output (see lemon_b200.score_pairs) plus 'rows' = (r0, r1)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    r0, r1, per = shard_bounds(n_total, world, rank)
    assert img_local.shape[0] == per and txt_local.shape[0] == per, "shards must be padded to equal length"
    if world > 1:
        img_db = allgather_rows(img_local, n_total, group)
        txt_db = allgather_rows(txt_local, n_total, group)
        lab_db = None
        if text_label_ids_local is not None:
            lab_db = allgather_rows(text_label_ids_local.view(-1, 1), n_total, group).view(-1)
    else:
        img_db, txt_db = img_local[:n_total], txt_local[:n_total]
        lab_db = text_label_ids_local[:n_total] if text_label_ids_local is not None else None
    if scorer is None:
        from .scoring import get_scorer
        scorer = get_scorer(img_local.device.index)
    scorer.set_database(img_db, txt_db, dist_type, normalize, lab_db)
    qid = torch.arange(r0, r1, dtype=torch.int64, device=img_local.device)
    out = scorer.score(None, None, k=k, query_in_db=qid, hparams=hparams, return_records=return_records,
                       query_rows=(r0, r1), text_label_ids_q=lab_db[r0:r1] if lab_db is not None else None)
    out["rows"] = (r0, r1)
    return out
