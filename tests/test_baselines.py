"""Discrepancy baseline: fused GPU scores vs the oracle restatement."""
import numpy as np
import pytest

from tests.helpers import clustered_pairs


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["dis_x", "dis_y", "div_x", "div_y"])
@pytest.mark.parametrize("train", [False, True])
def test_discrepancy_scores_vs_oracle(method, train):
    from lemon_b200 import baselines
    from oracle import baselines_oracle as B
    x, y, _, _ = clustered_pairs(3000, 128, n_clusters=30, seed=55)          # no duplicate captions: unambiguous lists
    nq = 500
    got = baselines.discrepancy_scores(x[:nq], y[:nq], x, y, k=5, method=method, train=train).cpu().numpy()
    ref, _ = B.discrepancy_scores(x[:nq], y[:nq], x, y, k=5, method=method, train=train)
    close = np.isclose(got, ref, rtol=2e-5, atol=2e-6)
    assert close.mean() > 0.99, (method, train, np.abs(got - ref).max())      # eps-tied neighbour lists may differ


def test_oracle_identities():
    from oracle import baselines_oracle as B
    x, y, _, _ = clustered_pairs(200, 32, n_clusters=8, seed=3)
    s, I = B.discrepancy_scores(x[:20], y[:20], x, y, k=4, method="div_y")
    assert s.shape == (20,) and I.shape == (20, 4) and (s >= -1e-9).all()
    s2, _ = B.discrepancy_scores(x[:20], y[:20], x, y, k=4, method="dis_x", train=True)
    assert np.isfinite(s2).all()
