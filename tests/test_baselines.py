"""Discrepancy baseline: fused GPU scores vs the oracle restatement."""
import numpy as np
import pytest

from tests.helpers import clustered_pairs


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["dis_x", "dis_y", "div_x", "div_y"])
@pytest.mark.parametrize("train", [False, True])
def test_discrepancy_scores_vs_oracle(method, train):
    """Two separate claims, both for EVERY row: (1) the searched neighbour lists satisfy the kNN acceptance rule
    against the float64 oracle (set equality modulo eps-ties at the boundary); (2) given the lists the GPU used, the
    score equals the oracle's formula.  Rows whose lists equal the oracle's are additionally compared end to end."""
    from lemon_b200 import baselines
    from oracle import baselines_oracle as B
    from oracle import lemon_oracle as O
    x, y, _, _ = clustered_pairs(3000, 128, n_clusters=30, seed=55)          # no duplicate captions: unambiguous lists
    nq, k = 500, 5
    got, nn, cache = baselines.discrepancy_scores(x[:nq], y[:nq], x, y, k=k, method=method, train=train, return_lists=True)
    got, nn = got.cpu().numpy(), nn.cpu().numpy().astype(np.int64)
    xn, yn = O.normalize_vectors(x), O.normalize_vectors(y)
    kk = k + int(train)
    r = O.compare_neighbor_sets(yn[:nq], yn, nn, kk, "ip")
    assert r["wrong"] == 0
    cache_l = None
    if cache is not None:
        cache = cache.cpu().numpy().astype(np.int64)
        rc = O.compare_neighbor_sets(yn, yn, cache, k + 1, "ip")
        assert rc["wrong"] == 0
        cache_l = [[j for j in row.tolist() if j != i] for i, row in enumerate(cache)]
    emb = xn if method.endswith("_x") else yn
    qv = xn[:nq] if method.endswith("_x") else yn[:nq]
    on_lists = B.scores_given_lists(emb, qv, nn, cache_l, k, method)
    np.testing.assert_allclose(got, on_lists, rtol=2e-5, atol=2e-6)           # every row, no exceptions
    ref, I_ref = B.discrepancy_scores(x[:nq], y[:nq], x, y, k=k, method=method, train=train)
    same = (np.sort(nn, axis=1) == np.sort(I_ref, axis=1)).all(axis=1)
    if cache is None:
        np.testing.assert_allclose(got[same], ref[same], rtol=2e-5, atol=2e-6)
    assert same.mean() > 0.99


@pytest.mark.parametrize("method", ["dis_x", "dis_y", "div_x", "div_y"])
@pytest.mark.parametrize("train", [False, True])
def test_baselines_oracle_agrees_with_the_script_restated_in_torch(method, train):
    """Pins oracle.baselines_oracle (float64 numpy on the oracle kNN) to an independent operation-for-operation fp32
    torch restatement of discrepancy_baseline.py:163-230."""
    from oracle import baselines_oracle as B
    x, y, _, _ = clustered_pairs(900, 48, n_clusters=12, seed=56)
    ref, I_ref = B.discrepancy_scores(x[:120], y[:120], x, y, k=5, method=method, train=train)
    alt, I_alt = B.reference_loop_torch(x[:120], y[:120], x, y, k=5, method=method, train=train)
    same = (I_ref == I_alt).all(axis=1)
    assert same.mean() > 0.97                                                 # fp32 vs float64 near-ties may swap
    np.testing.assert_allclose(alt[same], ref[same], rtol=3e-5, atol=3e-6)


def test_oracle_identities():
    from oracle import baselines_oracle as B
    x, y, _, _ = clustered_pairs(200, 32, n_clusters=8, seed=3)
    s, I = B.discrepancy_scores(x[:20], y[:20], x, y, k=4, method="div_y")
    assert s.shape == (20,) and I.shape == (20, 4) and (s >= -1e-9).all()
    s2, _ = B.discrepancy_scores(x[:20], y[:20], x, y, k=4, method="dis_x", train=True)
    assert np.isfinite(s2).all()
