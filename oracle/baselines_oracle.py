"""CPU restatement of the discrepancy / diversity baseline scores, lib/baselines/discrepancy_baseline.py:147-230.
TEST INFRASTRUCTURE ONLY.  Parity unpinned: the reference is a flat script (argparse + dataset loading at import) with
no tests; its arithmetic is restated here in float64 on the oracle's kNN."""
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
