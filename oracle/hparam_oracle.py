"""CPU oracle for the F1-threshold objective and the hyper-parameter grid.  TEST
INFRASTRUCTURE ONLY (see oracle/lemon_oracle.py).

Restates, in numpy/pure Python:
* ``optimize_f1_efficient``  lib/metrics/utils.py:286-296: ``fminbound(neg_f1, score.min(), score.max(), xtol=1e-8)``
  then ``best_f1 = -neg_f1(best_thres)`` with ``neg_f1(t) = -f1_score(y, score >= t)``;
* the grid stage of ``maximize_metric``  lib/metrics/utils.py:167-186 (iteration order of ``combinations_base``,
  tau_1 -> tau_1_n = tau_1_m, tau_2 -> tau_2_n = tau_2_m, strict ``>`` so the first best grid point wins).
Third-party arithmetic restated from its published algorithm: ``scipy.optimize.fminbound`` (Brent's bounded
minimiser, Forsythe-Malcolm-Moler ``fmin``; scipy 1.18 installed here, version unpinned in requirements.txt:10) and
``sklearn.metrics.f1_score`` for binary labels (2TP / (2TP + FP + FN), 0 when the denominator is 0).
Pinned by tests/golden/hparam_*.npz, generated with the reference's own functions run live.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

from . import lemon_oracle as O


def f1_at_threshold(y: np.ndarray, score: np.ndarray, thr: float) -> float:
    pred = score >= thr
    tp = int(np.count_nonzero(pred & (y != 0)))
    den = int(np.count_nonzero(pred)) + int(np.count_nonzero(y))      # (TP+FP) + (TP+FN)
    return 2.0 * tp / den if den > 0 else 0.0


def brent_bounded(func, x1: float, x2: float, xatol: float = 1e-8, maxfun: int = 500):
    """Brent's bounded scalar minimiser as scipy.optimize.fminbound runs it.  Returns (xf, fx, nfev)."""
    sqrt_eps = math.sqrt(2.2e-16)
    golden_mean = 0.5 * (3.0 - math.sqrt(5.0))
    a, b = float(x1), float(x2)
    fulc = a + golden_mean * (b - a)
    nfc = xf = fulc
    rat = e = 0.0
    x = xf
    fx = func(x)
    num = 1
    ffulc = fnfc = fx
    xm = 0.5 * (a + b)
    tol1 = sqrt_eps * abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        golden = True
        if abs(e) > tol1:                       # try a parabolic step
            golden = False
            r = (xf - nfc) * (fx - ffulc)
            q = (xf - fulc) * (fx - fnfc)
            p = (xf - fulc) * q - (xf - nfc) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = rat
            if abs(p) < abs(0.5 * q * r) and p > q * (a - xf) and p < q * (b - xf):
                rat = (p + 0.0) / q
                x = xf + rat
                if (x - a) < tol2 or (b - x) < tol2:
                    si = np.sign(xm - xf) + ((xm - xf) == 0)
                    rat = tol1 * si
            else:
                golden = True
        if golden:
            e = (a - xf) if xf >= xm else (b - xf)
            rat = golden_mean * e
        si = np.sign(rat) + (rat == 0)
        x = xf + si * max(abs(rat), tol1)
        fu = func(x)
        num += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            fulc, ffulc = nfc, fnfc
            nfc, fnfc = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if fu <= fnfc or nfc == xf:
                fulc, ffulc = nfc, fnfc
                nfc, fnfc = x, fu
            elif fu <= ffulc or fulc == xf or fulc == nfc:
                fulc, ffulc = x, fu
        xm = 0.5 * (a + b)
        tol1 = sqrt_eps * abs(xf) + xatol / 3.0
        tol2 = 2.0 * tol1
        if num >= maxfun:
            break
    return xf, fx, num


def optimize_f1_efficient(y, score, return_thres: bool = False):
    """lib/metrics/utils.py:286-296."""
    y = np.asarray(y)
    score = np.asarray(score, dtype=np.float64)
    thr, _, _ = brent_bounded(lambda t: -f1_at_threshold(y, score, t), score.min(), score.max(), xatol=1e-8)
    f1 = f1_at_threshold(y, score, thr)
    return (f1, thr) if return_thres else f1


def grid_points(grid: dict, force_zero=()):
    """Hyper-parameter vectors [beta, gamma, tau_1_n, tau_2_n, tau_1_m, tau_2_m] in the order utils.py:167-181
    visits them (itertools.product over the grid's key order)."""
    keys = list(grid.keys())
    out = []
    for values in itertools.product(*grid.values()):
        x = dict(zip(keys, values))
        g = []
        for name in O.HP_KEYS:
            if name in x:
                g.append(x[name])
            elif name in ("tau_1_n", "tau_1_m"):
                g.append(x["tau_1"])
            elif name in ("tau_2_n", "tau_2_m"):
                g.append(x["tau_2"])
            else:
                raise NotImplementedError(name)
            if name in force_zero:
                g[-1] = 0.0
        out.append([float(v) for v in g])
    return out


def grid_search(rec: dict, y, grid: dict, force_zero=()):
    """Grid stage of maximize_metric (utils.py:167-186).  Returns (best_x, best_f1, f1 of every grid point)."""
    best_x, best_val = None, -1.0
    vals = []
    for g in grid_points(grid, force_zero):
        hp = dict(zip(O.HP_KEYS, g))
        s, _, _ = O.calc_scores_vectorized(rec, hp)
        f1 = optimize_f1_efficient(y, s)
        vals.append(f1)
        if f1 > best_val:
            best_val, best_x = f1, g
    return best_x, best_val, np.array(vals)
