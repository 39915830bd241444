"""Signature-compatible replacement of ``lib.metrics.utils.calc_scores_given_hparams_vectorized``
(lib/metrics/utils.py:47-82): same arguments, same return values, the weighting and reduction run
in the ``lemon_combine_scores`` kernel.  The [N,k] columns are stacked once per DataFrame and kept
on the device (the reference re-stacks the object columns on every call, which dominates the
7056-point hyper-parameter grid, run_lemon.py:332-337)."""
from __future__ import annotations

import weakref

import numpy as np
import torch

from .scoring import get_scorer

_COLS = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")
_cache: dict = {}


def _stacked(df):
    key = id(df)
    ent = _cache.get(key)
    if ent is not None and ent[0]() is df and ent[1] == len(df):
        return ent[2]
    sc = get_scorer()
    rec = {c: torch.from_numpy(np.stack(df[c].values).astype(np.float32)).to(sc.device) for c in _COLS}
    rec["d_1"] = torch.from_numpy(np.asarray(df["d_1"].values, dtype=np.float64)).to(sc.device)
    try:
        _cache[key] = (weakref.ref(df, lambda _r, k=key: _cache.pop(k, None)), len(df), rec)
    except TypeError:
        pass
    return rec


def calc_scores_given_hparams_vectorized(df, best_hparams, return_dn=False, torch_arr=False):
    rec = _stacked(df)
    sc = get_scorer()
    scores, d_ns, d_ms = sc.combine_scores(rec, best_hparams)
    if torch_arr:
        scores, d_ns, d_ms = scores.cpu(), d_ns.cpu(), d_ms.cpu()
    else:
        scores, d_ns, d_ms = scores.cpu().numpy(), d_ns.cpu().numpy(), d_ms.cpu().numpy()
    if return_dn:
        return scores, d_ns, d_ms
    return scores


def clear_cache():
    _cache.clear()
