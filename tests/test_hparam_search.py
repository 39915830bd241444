"""The reference's hyper-parameter SEARCH (lib/metrics/utils.py:117-196) through the product's drop-ins.

CPU: the port of the search stage (oracle/hparam_search_port.py) reproduces golden answers of the live reference
(tests/golden/make_golden_hparam_search.py), which makes it a valid stand-in for ``lib.metrics.utils`` on the GPU box.
GPU: ``patch_reference_metrics`` / ``patch_reference_hparam_search`` are applied to that module exactly as they would
be applied to the reference's, and ``maximize_metric`` (scipy stages, the LBFGS stage that back-propagates through
``calc_scores_given_hparams_vectorized(torch_arr=True)``, the grid stage) runs end to end."""
import importlib
import os

import numpy as np
import pytest
import torch

from tests.helpers import hparam_search_case

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hparam_search.npz"))
X_PROBE = [1.5, 0.5, 0.3, 2.0, 0.7, 1.0]


def fresh_port():
    from oracle import hparam_search_port as P
    return importlib.reload(P)


@pytest.mark.parametrize("tag", ["plain", "ablate"])
def test_port_differentiable_stage_matches_live_reference(tag):
    P = fresh_port()
    df, grid, x0s, fz, fo = hparam_search_case(tag)
    x = torch.tensor(X_PROBE, dtype=torch.float64, requires_grad=True)
    loss = P.optim_func_torch(x, df, force_zero=fz, force_one=fo)
    loss.backward()
    assert abs(loss.item() - float(G[f"{tag}_loss"])) < 1e-12
    np.testing.assert_allclose(x.grad.numpy(), G[f"{tag}_grad"], rtol=1e-9, atol=1e-12)
    lb = P.maximize_metric_torch(df, x0s[1], P.optimize_f1_efficient, {}, force_zero=fz, force_one=fo)
    np.testing.assert_allclose(lb["x"], G[f"{tag}_lbfgs_x"], rtol=1e-6, atol=1e-8)


def test_port_maximize_metric_matches_live_reference():
    P = fresh_port()
    df, grid, x0s, fz, fo = hparam_search_case("ablate")
    bx, bv, bt = P.maximize_metric(df, grid, x0s, P.optimize_f1_efficient, {}, force_zero=fz, force_one=fo)
    assert bv == float(G["ablate_best_val"])
    np.testing.assert_allclose(np.asarray(bx, np.float64), G["ablate_best_x"], rtol=1e-9, atol=1e-12)
    assert bt == float(G["ablate_best_thr"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "ablate"])
def test_gpu_dropins_keep_the_lbfgs_stage_differentiable(tag):
    """after patch_reference_metrics the LBFGS closure (utils.py:129-141) must still be able to call
    loss.backward(); loss and gradient equal the live reference's."""
    import lemon_b200
    P = lemon_b200.patch_reference_metrics(fresh_port())
    df, grid, x0s, fz, fo = hparam_search_case(tag)
    x = torch.tensor(X_PROBE, dtype=torch.float64, requires_grad=True)
    loss = P.optim_func_torch(x, df, force_zero=fz, force_one=fo)
    loss.backward()
    assert abs(loss.item() - float(G[f"{tag}_loss"])) < 1e-6
    np.testing.assert_allclose(x.grad.numpy(), G[f"{tag}_grad"], rtol=1e-4, atol=1e-7)
    lb = P.maximize_metric_torch(df, x0s[1], P.optimize_f1_efficient, {}, force_zero=fz, force_one=fo)
    # 20 x 20 strong-Wolfe LBFGS iterations amplify last-bit differences (GPU vs CPU exp): same basin, not same digits
    assert abs(lb["fun"] - float(G[f"{tag}_lbfgs_fun"])) < 0.02 * float(G[f"{tag}_lbfgs_fun"])
    # numpy call style through the same patched function: plain floats in, float64 array out, reference values
    hp = P.unpack_vector(X_PROBE, fz, fo)
    s = P.calc_scores_given_hparams_vectorized(df, hp)
    import oracle.hparam_search_port as Q
    np.testing.assert_allclose(s, importlib.reload(Q).calc_scores_given_hparams_vectorized(df, hp), rtol=1e-5, atol=1e-6)
    # in-place edit of the cached frame (the 'd1' ablation writes df['d_1']) must not return stale scores
    df["d_1"] = 0.0
    s0 = P.calc_scores_given_hparams_vectorized(df, hp)
    assert np.abs(s0 - s).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "ablate"])
def test_gpu_maximize_metric_end_to_end(tag):
    """run_lemon.py:386-394 through patch_reference_hparam_search: scipy + LBFGS stages on the patched scoring
    function and F1 objective, grid stage in one lemon_f1_grid launch.  The GPU scores differ from the reference's
    fp32 numpy scores at ~1e-7 relative, which can flip borderline samples of the F1 objective: the optimum found
    must be as good as the live reference's up to such samples, and self-consistent."""
    from lemon_b200 import hparam_compat
    from oracle import hparam_oracle as H
    P = hparam_compat.patch_reference_hparam_search(fresh_port())
    df, grid, x0s, fz, fo = hparam_search_case(tag)
    n = len(df)
    bx, bv, bt = P.maximize_metric(df, grid, x0s, P.optimize_f1_efficient, {}, force_zero=fz, force_one=fo)
    assert bv >= float(G[f"{tag}_best_val"]) - 3.0 / n
    import oracle.hparam_search_port as Q
    Q = importlib.reload(Q)
    s = Q.calc_scores_given_hparams_vectorized(df, Q.unpack_vector(list(bx), fz, fo))
    assert abs(H.f1_at_threshold(df["is_mislabel"].values, np.asarray(s, np.float64), bt) - bv) <= 3.0 / n
    for c, name in enumerate(Q.NAMES):
        if name in fz:
            assert bx[c] == 0.0
        if name in fo:
            assert bx[c] == 1.0
    # an objective the GPU grid does not implement runs the reference loop unchanged
    other = lambda y, score, return_thres=False: Q.optimize_f1_efficient(y, score, return_thres)
    bx2, bv2, _ = P.maximize_metric(df, {k: v[:1] for k, v in grid.items()}, x0s[:1], other, {}, force_zero=fz, force_one=fo,
                                    scipy_methods=["Nelder-Mead"])
    assert bv2 > 0
