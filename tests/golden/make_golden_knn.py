"""Generate tests/golden/knn.npz: kNN answers from implementations INDEPENDENT of the oracle.

faiss (the library the reference calls at run_lemon.py:167-176,235-236) is absent from /root/reference and from
this image, so the kNN boundary cannot be pinned to faiss itself.  It is pinned instead to two independent
brute-force searches that implement the same published contract (exact top-k by inner product, descending /
squared L2, ascending):

* scikit-learn ``NearestNeighbors(algorithm="brute")`` on float64 copies: ``metric="sqeuclidean"`` for IndexFlatL2;
  for IndexFlatIP on UNIT-NORM rows ``metric="cosine"`` (1 - <q,b> ranks like -<q,b>);
* torch float64 ``(q @ db.T).topk`` / ``cdist**2 .topk(largest=False)`` for the un-normalised inner-product case;
* exact duplicate rows (mass ties): a pure-Python ``sorted`` over (value, index) pairs, the documented total order
  (best value first, then ascending DB index).

Run in the builder container:  python tests/golden/make_golden_knn.py
The inputs are regenerated from seeds (tests/helpers.knn_pin_case); the file holds the answers only.
tests/test_oracle_pin.py checks oracle.knn_search against the file, tests/test_gpu_parity.py the CUDA path.
"""
import os
import sys

import numpy as np
import torch
from sklearn.neighbors import NearestNeighbors

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from tests.helpers import knn_pin_case, KNN_PIN_CASES
    out = {}
    for tag in KNN_PIN_CASES:
        db, q, k, kind = knn_pin_case(tag)
        if kind == "unit":        # unit-norm rows: IP via sklearn cosine, L2 via sklearn sqeuclidean
            nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(db.astype(np.float64))
            dist, idx = nn.kneighbors(q.astype(np.float64))
            out[f"{tag}_ip_I"], out[f"{tag}_ip_D"] = idx.astype(np.int64), 1.0 - dist
            nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(db.astype(np.float64))
            dist, idx = nn.kneighbors(q.astype(np.float64))
            out[f"{tag}_l2_I"], out[f"{tag}_l2_D"] = idx.astype(np.int64), dist
        elif kind == "raw":       # un-normalised rows (norms 0.5 .. 3): torch float64
            S = torch.from_numpy(q).double() @ torch.from_numpy(db).double().T
            v, i = S.topk(k, dim=1)
            out[f"{tag}_ip_I"], out[f"{tag}_ip_D"] = i.numpy(), v.numpy()
            d2 = torch.cdist(torch.from_numpy(q).double(), torch.from_numpy(db).double(),
                             compute_mode="donot_use_mm_for_euclid_dist") ** 2
            v, i = d2.topk(k, dim=1, largest=False)
            out[f"{tag}_l2_I"], out[f"{tag}_l2_D"] = i.numpy(), v.numpy()
        else:                     # mass ties: pure-Python total order (best value, then ascending DB index)
            I = np.empty((len(q), k), np.int64)
            D = np.empty((len(q), k), np.float64)
            for r in range(len(q)):
                vals = [float(np.dot(q[r].astype(np.float64), db[j].astype(np.float64))) for j in range(len(db))]
                order = sorted(range(len(db)), key=lambda j: (-vals[j], j))[:k]
                I[r], D[r] = order, [vals[j] for j in order]
            out[f"{tag}_ip_I"], out[f"{tag}_ip_D"] = I, D
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **out)
    print("wrote knn.npz", {k: v.shape for k, v in out.items() if k.endswith("_I")})


if __name__ == "__main__":
    main()
