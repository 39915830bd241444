"""Port of the reference's hyper-parameter SEARCH stage, used as a stand-in for the module
``lib.metrics.utils`` where /root/reference is not mounted (the GPU box).  TEST INFRASTRUCTURE ONLY
(see oracle/lemon_oracle.py): the product's drop-ins (``lemon_b200.patch_reference_metrics``,
``lemon_b200.hparam_compat.patch_reference_hparam_search``) are exercised by patching THIS module's
attributes exactly as they would patch the reference's, and the result is compared with golden answers of
the live reference (tests/golden/make_golden_hparam_search.py).

Restated (lib/metrics/utils.py): combinations_base :18-19, calc_scores_given_hparams_vectorized :47-82 (numpy
and torch_arr branches), unpack_vector :84-103, optim_func :117-121, optim_func_torch :123-127, torch_minimize
:129-141, maximize_metric_scipy :143-146, maximize_metric_torch :148-149, maximize_metric :151-196,
optimize_f1_efficient :286-296 (through oracle.hparam_oracle, pinned bit-exact).
"""
from __future__ import annotations

import itertools

import numpy as np
import torch
from scipy.optimize import minimize

from . import hparam_oracle as H

NAMES = ("beta", "gamma", "tau_1_n", "tau_2_n", "tau_1_m", "tau_2_m")


def combinations_base(grid):
    keys = list(grid.keys())
    return [dict(zip(keys, vals)) for vals in itertools.product(*grid.values())]


def calc_scores_given_hparams_vectorized(df, best_hparams, return_dn=False, torch_arr=False):
    hp = best_hparams
    if torch_arr:
        col = lambda c: torch.stack([torch.tensor(a) for a in df[c].values])
        ex, sm = torch.exp, lambda a: torch.sum(a, dim=1)
        d1 = torch.tensor(df["d_1"].values)
    else:
        col = lambda c: np.stack(df[c].values)
        ex, sm = np.exp, lambda a: np.sum(a, axis=1)
        d1 = df["d_1"].values
    w_n = ex(-hp["tau_1_n"] * col("D_n")) * ex(-hp["tau_2_n"] * col("dists_tr_n"))
    w_m = ex(-hp["tau_1_m"] * col("D_m")) * ex(-hp["tau_2_m"] * col("dists_tr_m"))
    d_ns = sm(w_n * col("dists_n")) / col("D_n").shape[1]
    d_ms = sm(w_m * col("dists_m")) / col("D_m").shape[1]
    scores = d1 + hp["beta"] * d_ns + hp["gamma"] * d_ms
    return (scores, d_ns, d_ms) if return_dn else scores


def unpack_vector(x, force_zero=[], force_one=[]):
    cand = {name: x[i] for i, name in enumerate(NAMES)}
    for name in cand:
        if name in force_zero:
            cand[name] = 0.
    for name in cand:
        if name in force_one:
            cand[name] = 1.
    return cand


def optimize_f1_efficient(y, score, return_thres=False):
    return H.optimize_f1_efficient(np.asarray(y), np.asarray(score, dtype=np.float64), return_thres)


def optim_func(x, df, obj_func, obj_func_args, force_zero=[], force_one=[]):
    hp = unpack_vector(x, force_zero=force_zero, force_one=force_one)
    score = calc_scores_given_hparams_vectorized(df, hp, return_dn=False)
    return -obj_func(df["is_mislabel"].values, score, **obj_func_args)


def optim_func_torch(x, df, force_zero=[], force_one=[]):
    hp = unpack_vector(x, force_zero=force_zero, force_one=force_one)
    y = df["is_mislabel"].values
    score = calc_scores_given_hparams_vectorized(df, hp, return_dn=False, torch_arr=True)
    return torch.nn.SoftMarginLoss()(score, torch.from_numpy(y).double() * 2 - 1)


def torch_minimize(fn, x0, args, options={"max_iter": 20, "line_search_fn": "strong_wolfe"}):
    x = torch.tensor(x0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.LBFGS([x], lr=0.1, max_iter=options["max_iter"], line_search_fn=options["line_search_fn"])

    def closure():
        opt.zero_grad()
        loss = fn(x, args[0])
        loss.backward()
        return loss

    for _ in range(options["max_iter"]):
        opt.step(closure)
    return {"x": x.detach().numpy(), "fun": closure().item()}


def maximize_metric_scipy(df, x0, obj_func, obj_func_args, method, force_zero=[], force_one=[]):
    return minimize(optim_func, x0, method=method, args=(df, obj_func, obj_func_args, force_zero, force_one), options={})


def maximize_metric_torch(df, x0, obj_func, obj_func_args, force_zero=[], force_one=[]):
    return torch_minimize(optim_func_torch, x0, args=(df, obj_func, obj_func_args, force_zero, force_one))


def maximize_metric(df, grid, x0s, obj_func, obj_func_args, force_zero=[], force_one=[],
                    scipy_methods=["Powell", "Nelder-Mead"]):
    # the three stages look the module attributes up at call time, so that patched drop-ins are used (as in the
    # reference, whose functions resolve each other through the module globals)
    best_x, best_val = None, -1
    for x0 in x0s:
        for method in scipy_methods:
            r = maximize_metric_scipy(df, x0, obj_func, obj_func_args, method=method, force_zero=force_zero,
                                      force_one=force_one)
            if -r.fun > best_val:
                best_val, best_x = -r.fun, r.x
    for x0 in x0s:
        cx = maximize_metric_torch(df, x0, obj_func, obj_func_args, force_zero=force_zero, force_one=force_one)["x"]
        v = optim_func(cx, df, obj_func, obj_func_args, force_zero=force_zero, force_one=force_one)
        if -v > best_val:
            best_val, best_x = -v, cx
    for pt in combinations_base(grid):
        g = []
        for name in NAMES:
            if name in pt:
                g.append(pt[name])
            elif name in ("tau_1_n", "tau_1_m"):
                g.append(pt["tau_1"])
            elif name in ("tau_2_n", "tau_2_m"):
                g.append(pt["tau_2"])
            else:
                raise NotImplementedError(name)
            if name in force_zero:
                g[-1] = 0.
        v = optim_func(g, df, obj_func, obj_func_args, force_zero=force_zero, force_one=force_one)
        if -v > best_val:
            best_val, best_x = -v, g
    for c, name in enumerate(NAMES):
        if name in force_zero:
            best_x[c] = 0.
        if name in force_one:
            best_x[c] = 1.
    score = calc_scores_given_hparams_vectorized(df, unpack_vector(best_x, force_zero=force_zero, force_one=force_one))
    return best_x, best_val, obj_func(df["is_mislabel"], score, return_thres=True, **obj_func_args)[1]
