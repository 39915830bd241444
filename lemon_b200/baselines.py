"""Discrepancy / diversity baseline (lib/baselines/discrepancy_baseline.py:147-230) on the same kernels
(SURVEY.md §8f-3): text kNN through the tensor-core search, second-order neighbour gathers fused into one kernel.
The reference script's faiss calls also run unmodified through lemon_b200.faiss_compat."""
from __future__ import annotations

import ctypes as C

import torch

from .scoring import _ptr, _stream, get_scorer

METHODS = ("dis_x", "dis_y", "div_x", "div_y")


def discrepancy_scores(img_q, txt_q, img_db, txt_db, *, k: int, method: str, train: bool = False, normalize: bool = True,
                       device=None, return_lists: bool = False):
    """pred_score of discrepancy_baseline.py for every query pair (float32 [N]).  `train=True` searches k+1 text
    neighbours as the script does for the train split (it keeps all k+1, :210-215)."""
    assert method in METHODS
    sc = get_scorer(device)
    ydb = sc.prepare_db(txt_db, normalize)
    yq = sc.prepare(txt_q, normalize)
    kk = k + int(train)
    _, nn = sc.knn(yq, ydb, kk, 0)                              # index_txt.search(text_embeds, k + train)  :210
    mode = 0 if method.startswith("dis") else 1
    side_x = method.endswith("_x")
    emb = sc.prepare(img_db, normalize, need_f16=False) if side_x else ydb
    qemb = (sc.prepare(img_q, normalize, need_f16=False) if side_x else yq) if mode == 0 else None
    cache = None
    kc = 0
    if mode == 0:                                               # cache of NNs for the train set  :165-168
        kc = k + 1
        _, cache = sc.knn(ydb, ydb, kc, 0)
    out = torch.empty(yq.n, dtype=torch.float32, device=sc.device)
    with torch.cuda.device(sc.device):
        sc.ctx.check(sc.lib.lemon_discrepancy(sc.ctx.handle, _ptr(emb.f32), _ptr(qemb.f32) if qemb is not None else None,
                                              _ptr(nn), _ptr(cache), yq.n, ydb.n, emb.d, kk, kc, k, mode, _ptr(out),
                                              _stream()), "lemon_discrepancy")
    if return_lists:        # the text-kNN list of every query and (dis_*) of every DB row, as searched
        return out, nn, cache
    return out
