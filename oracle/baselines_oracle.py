"""CPU restatement of the discrepancy / diversity baseline scores, lib/baselines/discrepancy_baseline.py:147-230.
TEST INFRASTRUCTURE ONLY.  The reference is a flat script (argparse + dataset loading at import) with no tests, so it
cannot be imported; its arithmetic is restated here in float64 on the oracle's kNN (`discrepancy_scores`) and, as
an independent second implementation, operation for operation in fp32 torch with Python lists exactly as the script
writes it (`reference_loop_torch`, :163-230); tests/test_baselines.py checks that the two agree.  The kNN itself is
pinned by tests/test_oracle_pin.py."""
from __future__ import annotations

import numpy as np

from . import lemon_oracle as O


def discrepancy_scores(img_q, txt_q, img_db, txt_db, *, k: int, method: str, train: bool = False, normalize: bool = True):
    if normalize:
        img_q, txt_q, img_db, txt_db = (O.normalize_vectors(a) for a in (img_q, txt_q, img_db, txt_db))
    kk = k + int(train)
    _, I_m = O.knn_search(txt_q, txt_db, kk, "ip")                        # :210
    cache = None
    if method.startswith("dis"):                                          # :165-168
        _, c = O.knn_search(txt_db, txt_db, k + 1, "ip")
        cache = [[j for j in c[i].tolist() if j != i] for i in range(len(c))]
    emb = np.asarray(img_db if method.endswith("_x") else txt_db, np.float64)
    qv = np.asarray(img_q if method.endswith("_x") else txt_q, np.float64)
    out = np.empty(len(I_m))
    for i in range(len(I_m)):
        if method.startswith("dis"):                                      # :217-224
            second = [l for j in I_m[i] for l in cache[j]]
            V = 1 - emb[second] @ qv[i]
            out[i] = V.sum() / len(second)
        else:                                                             # :225-230
            E = emb[I_m[i]]
            out[i] = (1 - E @ E.T).sum() / k ** 2
    return out, I_m


def scores_given_lists(emb, qv, I_m, cache, k: int, method: str) -> np.ndarray:
    """The score formulas alone (:217-230) for GIVEN neighbour lists (float64): lets a test separate "are the lists
    right" (kNN acceptance rule) from "is the score right given the lists" (exact)."""
    emb = np.asarray(emb, np.float64)
    out = np.empty(len(I_m))
    for i in range(len(I_m)):
        if method.startswith("dis"):
            second = [l for j in I_m[i] for l in cache[j] if l != j]
            out[i] = (1 - emb[second] @ np.asarray(qv[i], np.float64)).sum() / len(second)
        else:
            E = emb[I_m[i]]
            out[i] = (1 - E @ E.T).sum() / k ** 2
    return out


def reference_loop_torch(img_q, txt_q, img_db, txt_db, *, k: int, method: str, train: bool = False):
    """discrepancy_baseline.py:147-230 operation for operation (fp32 torch, Python list comprehensions, the same
    expressions), with `torch.topk` of the fp32 similarity matrix standing in for ``index_txt.search``."""
    import torch
    F = torch.nn.functional
    emb_txt_tr = F.normalize(torch.as_tensor(txt_db, dtype=torch.float32), p=2, dim=1)
    emb_img_tr = F.normalize(torch.as_tensor(img_db, dtype=torch.float32), p=2, dim=1)
    text_embeds = F.normalize(torch.as_tensor(txt_q, dtype=torch.float32), p=2, dim=1)
    img_embeds = F.normalize(torch.as_tensor(img_q, dtype=torch.float32), p=2, dim=1)
    search = lambda q, kk: torch.topk(q @ emb_txt_tr.T, kk, dim=1)
    if "dis" in method:
        _, cache = search(emb_txt_tr, k + 1)
        cache = cache.tolist()
        for i in range(len(cache)):
            cache[i] = [j for j in cache[i] if j != i]
    D_ms, I_ms = search(text_embeds, k + int(train))
    scores = []
    for i in range(len(img_embeds)):
        img_embed = img_embeds[i, None]
        text_embed = text_embeds[i, None]
        I_m = I_ms[i].tolist()
        if method == "dis_y":
            second_nns = [l for j in I_m for l in cache[j]]
            V = 1 - emb_txt_tr[second_nns] @ (text_embed.T)
            score = V.sum() / len(second_nns)
        elif method == "dis_x":
            second_nns = [l for j in I_m for l in cache[j]]
            V = 1 - emb_img_tr[second_nns] @ (img_embed.T)
            score = V.sum() / len(second_nns)
        elif method == "div_y":
            U = 1 - emb_txt_tr[I_m] @ (emb_txt_tr[I_m].T)
            score = U.sum() / k ** 2
        else:
            U = 1 - emb_img_tr[I_m] @ (emb_img_tr[I_m].T)
            score = U.sum() / k ** 2
        scores.append(float(score))
    return np.asarray(scores), I_ms.numpy()
