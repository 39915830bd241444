"""Drop-in for the subset of the ``faiss`` module API the reference uses on the scoring path:
``faiss.IndexFlatIP(d)`` / ``faiss.IndexFlatL2(d)``, ``.add(x)``, ``.search(q, k) -> (D, I)``
(run_lemon.py:14,167-176,235-236; also lib/baselines/discrepancy_baseline.py:150-166,209).

Install with ``lemon_b200.install_faiss_shim()`` (``sys.modules['faiss'] = this module``) and
run_lemon.py:166-176,235-236 run unmodified on the B200 kernels.  Semantics kept from faiss:
IP results descending, L2 results are SQUARED distances ascending, D float32 / I int64,
fresh arrays returned, numpy in -> numpy out, ``ntotal < k`` pads with I=-1, D=-inf/+inf,
``add`` copies.  Tie order is defined here (faiss leaves it implementation-defined): equal
values are returned by ascending DB index.  No CPU path: a CUDA device is required.
"""
from __future__ import annotations

import numpy as np
import torch

from .scoring import get_scorer, Prepared

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


class _IndexFlat:
    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int):
        self.d = int(d)
        self.ntotal = 0
        self.is_trained = True
        self._chunks: list[torch.Tensor] = []
        self._prepared: Prepared | None = None
        self._scorer = None

    def _sc(self):
        if self._scorer is None:
            self._scorer = get_scorer()
        return self._scorer

    def add(self, x):
        x = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise RuntimeError(f"Error: 'd == x.shape[1]' failed: expected d={self.d}, got {tuple(x.shape)}")
        self._chunks.append(x.to(device=self._sc().device, dtype=torch.float32, copy=True))
        self.ntotal += x.shape[0]
        self._prepared = None

    def reset(self):
        self._chunks, self._prepared, self.ntotal = [], None, 0

    def _db(self) -> Prepared:
        if self._prepared is None:
            allx = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks)
            self._chunks = [allx]
            self._prepared = self._sc().prepare(allx, normalize=False)
            self._prepared.dedup = self._sc().find_duplicates(self._prepared)
        return self._prepared

    def search(self, x, k: int):
        was_numpy = isinstance(x, np.ndarray)
        xt = torch.as_tensor(np.ascontiguousarray(x) if was_numpy else x)
        if xt.dim() != 2 or xt.shape[1] != self.d:
            raise RuntimeError(f"Error: 'd == x.shape[1]' failed: expected d={self.d}, got {tuple(xt.shape)}")
        k = int(k)
        nq = xt.shape[0]
        sc = self._sc()
        if k > 64:
            raise RuntimeError("lemon_b200.faiss_compat: k > 64 is not supported")
        if self.ntotal > 0 and nq > 0:
            q = sc.prepare(xt, normalize=False)
            # the kernels already pad like faiss (I = -1, D = -inf / +inf when ntotal < k): the lists ARE the answer
            D, ti = sc.knn(q, self._db(), k, self.metric_type)
            I = ti.to(torch.int64)
        else:
            fill = float("-inf") if self.metric_type == METRIC_INNER_PRODUCT else float("inf")
            D = torch.full((nq, k), fill, dtype=torch.float32, device=sc.device)
            I = torch.full((nq, k), -1, dtype=torch.int64, device=sc.device)
        if was_numpy or not xt.is_cuda:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I


class IndexFlatIP(_IndexFlat):
    metric_type = METRIC_INNER_PRODUCT


class IndexFlatL2(_IndexFlat):
    metric_type = METRIC_L2
