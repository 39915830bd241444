"""lemon_b200 — B200-native (sm_100a) implementation of the LEMoN pair-scoring hot path.

Public surface (mirrors the reference's seams):
  score_pairs(...)                                  fused replacement of run_lemon.py:163-314,406-407
  LemonScorer                                       set_database / score / knn / combine_scores
  faiss_compat.IndexFlatIP / IndexFlatL2            drop-in for the faiss calls at run_lemon.py:167-176,235-236
  metrics_compat.calc_scores_given_hparams_vectorized   drop-in for lib/metrics/utils.py:47-82
  dist.score_pairs_sharded                          row-sharded multi-GPU driver (one NCCL all-gather per modality)
  handoff.extract_and_score                         encoder outputs -> device shard -> staged database -> scores (no host copy)
  subsample_db / query_in_db_from_indices           the reference's DB cap (run_lemon.py:48,122-127) and membership rule
  filter_lowest_scores                              CC3M consumer of the scores (train_clip_from_scratch.py:110-113)
"""
from . import _lib
from ._lib import LemonError, LIB_PATH
from .scoring import (LemonScorer, score_pairs, get_scorer, plan_segments, plan_tail, HP_KEYS, subsample_db,
                      query_in_db_from_indices)

__all__ = ["LemonScorer", "score_pairs", "get_scorer", "plan_segments", "LemonError", "LIB_PATH", "HP_KEYS",
           "install_faiss_shim", "patch_reference_metrics", "subsample_db", "query_in_db_from_indices",
           "filter_lowest_scores"]


def filter_lowest_scores(score, n_keep: int, idx=None, device=None):
    """train_clip_from_scratch.py:110-113 (`score_df.sort_values(by='score').iloc[:cc3m_filtering_n]['idx'].values`) on
    the device: the `idx` values (default: row numbers) of the n_keep lowest-scored pairs in ascending score order."""
    import torch
    order, _ = get_scorer(device).keep_lowest(score, n_keep)
    if idx is None:
        return order
    idx_t = torch.as_tensor(idx).to(order.device)
    return idx_t[order]


def install_faiss_shim():
    """``import faiss`` in the reference then resolves to lemon_b200.faiss_compat."""
    import sys
    from . import faiss_compat
    sys.modules["faiss"] = faiss_compat
    return faiss_compat


def patch_reference_metrics(ref_metrics_utils_module):
    """Monkey-patch the reference's ``lib.metrics.utils`` so its callers (run_lemon.py:406,
    optim_func at utils.py:117-121, train_clip_from_scratch.py:110) use the CUDA scorer."""
    from . import metrics_compat
    ref_metrics_utils_module.calc_scores_given_hparams_vectorized = metrics_compat.calc_scores_given_hparams_vectorized
    return ref_metrics_utils_module
