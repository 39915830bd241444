"""Pins oracle.knn_search (the float64 restatement of faiss IndexFlatIP / IndexFlatL2, run_lemon.py:167-176,235-236)
to answers produced by INDEPENDENT brute-force implementations: scikit-learn NearestNeighbors(algorithm="brute"),
torch float64 topk and a pure-Python sort under the documented total order (tests/golden/make_golden_knn.py).
faiss itself is not installable here, so this is the strongest pin available for the kNN boundary."""
import os

import numpy as np
import pytest

from oracle import lemon_oracle as O
from tests.helpers import KNN_PIN_CASES, knn_pin_case

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "knn.npz"))


def _same_modulo_boundary_ties(I_got, D_got, I_ref, D_ref, eps):
    """Identical lists, except that entries whose value is within eps of a neighbour's may swap / be replaced at the
    k-th boundary (independent float64 summation orders differ by ~1e-16)."""
    bad = 0
    for r in range(I_ref.shape[0]):
        if (I_got[r] == I_ref[r]).all():
            continue
        sg, sr = set(I_got[r].tolist()), set(I_ref[r].tolist())
        kth = D_ref[r, -1]
        for j, i in enumerate(I_got[r]):
            if i not in sr and abs(D_got[r, j] - kth) > eps:
                bad += 1
        for j, i in enumerate(I_ref[r]):
            if i not in sg and abs(D_ref[r, j] - kth) > eps:
                bad += 1
        # same members: only the order inside a run of (nearly) equal values may differ
        if sg == sr and not np.allclose(np.sort(D_got[r]), np.sort(D_ref[r]), rtol=0, atol=eps):
            bad += 1
    return bad


@pytest.mark.parametrize("tag", KNN_PIN_CASES)
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_knn_oracle_matches_independent_implementations(tag, metric):
    if f"{tag}_{metric}_I" not in G:
        pytest.skip("case has no answer for this metric")
    db, q, k, kind = knn_pin_case(tag)
    D, I = O.knn_search(q, db, k, metric)
    I_ref, D_ref = G[f"{tag}_{metric}_I"], G[f"{tag}_{metric}_D"]
    # sklearn's cosine metric re-normalises the rows (fp32 unit rows are unit to ~3e-8), everything else is float64 rounding
    tol = 1e-7 if (kind == "unit" and metric == "ip") else 1e-9
    np.testing.assert_allclose(D, D_ref, rtol=tol, atol=tol)
    if kind == "ties":
        assert (I == I_ref).all()                                      # total order: exact duplicates by ascending index
    else:
        assert _same_modulo_boundary_ties(I, D, I_ref, D_ref, eps=2 * tol) == 0
        assert (I == I_ref).mean() > 0.995


def test_oracle_l2_matches_direct_difference_form():
    """The oracle expands ||q-b||^2 = ||q||^2 + ||b||^2 - 2<q,b> (as faiss does); in float64 that equals the direct
    sum of squared differences to ~1e-15, so both forms select the same neighbours."""
    db, q, k, _ = knn_pin_case("u")
    D, I = O.knn_search(q, db, k, "l2")
    direct = ((q[:, None, :].astype(np.float64) - db[I].astype(np.float64)) ** 2).sum(-1)
    np.testing.assert_allclose(D, direct, rtol=1e-9, atol=1e-12)
