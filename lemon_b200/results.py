"""Columnar results -> the reference's legacy record schema.

The fused path returns [N,k] arrays; run_lemon.py builds a list of per-sample dicts and a
``pd.DataFrame`` with object columns (run_lemon.py:291-314) that downstream code reads
(`res.pkl['df']`, lib/metrics/utils.py:64-69, train_clip_from_scratch.py:97-113, notebooks).
``records_to_dataframe`` produces that schema on demand, so the O(N) Python objects are only
created when a legacy consumer needs them."""
from __future__ import annotations

import numpy as np

RECORD_COLS = ("dists_n", "D_n", "dists_tr_n", "dists_m", "D_m", "dists_tr_m")


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def records_to_dataframe(out: dict, sset: str, idx_offset: int = 0, actual_label=None, actual_label_text=None,
                         noisy_label=None, noisy_label_text=None, is_mislabel=None):
    """DataFrame with the columns of run_lemon.py:291-307: sset, idx, [label columns], d_1 (Python
    float, as `.item()` gives), and the six length-k float32 arrays per row."""
    import pandas as pd
    d1 = _np(out["d_1"]).astype(np.float64)
    n = d1.shape[0]
    cols = {"sset": [sset] * n, "idx": np.arange(idx_offset, idx_offset + n)}
    for name, val in (("actual_label", actual_label), ("actual_label_text", actual_label_text),
                      ("noisy_label", noisy_label), ("noisy_label_text", noisy_label_text)):
        if val is not None:
            cols[name] = list(val)
    if is_mislabel is not None:
        mis = np.asarray(is_mislabel).astype(np.int64)
        cols["is_mislabel"] = mis
        cols["is_correct_label"] = 1 - mis
    cols["d_1"] = d1
    for c in RECORD_COLS:
        a = _np(out[c]).astype(np.float32)
        cols[c] = list(a)          # one float32 array of length k per row, as the reference stores them
    return pd.DataFrame(cols)


def dataframe_to_records(df) -> dict:
    """Inverse: stack a legacy DataFrame's object columns once ([N,k] float32 + d_1 float64)."""
    rec = {c: np.stack(df[c].values).astype(np.float32) for c in RECORD_COLS}
    rec["d_1"] = np.asarray(df["d_1"].values, dtype=np.float64)
    return rec
