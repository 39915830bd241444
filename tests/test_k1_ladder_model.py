"""CPU model of K1's threshold ladder (lemon_b200/csrc/k1_knn_tc.cu: Ladder, ladder_step, boot_select,
prune_exact) for ONE candidate list.  The CUDA kernel is checked on the GPU (tests/test_gpu_parity.py); this
model pins the ALGORITHM's invariants on the CPU over distributions the GPU tests do not sweep:

  1. every column that is not in the list has a value <= the final threshold (what the re-rank certificate uses);
  2. a finite threshold is certified: at least `cert` listed columns are >= it;
  3. the number of appended keys stays within a small factor of the ideal cert * ln(M / M0) on benign data, and
     the list never overflows its 1024 slots (adversarial orders go through the exact reduction instead).
"""
import zlib

import numpy as np
import pytest

CAP, CHUNK, BOOT_COLS = 1024, 32, 1024


def ladder_scan(vals: np.ndarray, cert: int):
    """Returns (listed column ids, final theta, total appends, exact reductions)."""
    vals = vals.astype(np.float32)
    m = len(vals)
    theta, lst, appends, prunes = -np.inf, [], 0, 0
    t = np.full(3, np.inf, np.float32)
    c = np.zeros(3, np.int64)
    delta = 0.0
    boot = m >= 16 * 128                       # items of >= 16 tiles (128 columns per group and tile)
    order = np.arange(m)
    if boot:
        mx = vals[:BOOT_COLS].reshape(-1, 8).max(1)          # 8-column group maxima of the bootstrap tiles
        srt = np.sort(mx)[::-1]
        th = srt[cert - 1]                                   # >= cert maxima (distinct columns) are >= th
        theta = np.nextafter(np.float32(th), np.float32(-np.inf))
        delta = max(float(srt[max(cert // 5, 1)] - th) / 2.2, 1e-6 * max(1.0, abs(float(theta))))
        t = np.float32(theta) + np.float32(delta) * np.arange(1, 4, dtype=np.float32)
        order = np.concatenate([np.arange(BOOT_COLS, m), np.arange(BOOT_COLS)])   # bootstrap tiles are re-scanned last
    for s0 in range(0, m, CHUNK):
        cols = order[s0:s0 + CHUNK]
        v = vals[cols]
        mxv = v.max()
        if mxv > theta:
            c += mxv > t
            if c[0] >= cert:                                  # certified level: theta moves up, ladder shifts
                theta = max(theta, float(t[0]))
                if c[1] >= 3 * cert // 4:
                    delta *= 1.5
                elif c[1] < 5 * cert // 16:
                    delta *= 0.75
                delta = max(delta, 1e-6 * max(1.0, abs(float(t[2]))))
                t = np.array([t[1], t[2], t[2] + np.float32(delta)], np.float32)
                c = np.array([c[1], c[2], 0])
            keep = cols[v > theta]
            lst.extend(keep.tolist())
            appends += len(keep)
        if len(lst) > CAP - CHUNK:                            # exact reduction to the best 64; re-seeds the ladder
            prunes += 1
            lv = vals[lst]
            best = np.argsort(-lv, kind="stable")[:64]
            sv = lv[best]
            lst = [lst[i] for i in best]
            theta = max(theta, float(sv[cert - 1]))
            t = np.array([sv[3 * cert // 4 - 1], sv[cert // 2 - 1], sv[cert // 4 - 1]], np.float32)
            c = np.array([3 * cert // 4, cert // 2, cert // 4])
            delta = max(float(t[2] - t[0]) * 0.625, 1e-6 * max(1.0, abs(theta)))
    return np.array(lst, np.int64), theta, appends, prunes


def _check(vals, cert):
    lst, theta, appends, prunes = ladder_scan(vals, cert)
    assert len(lst) <= CAP and len(set(lst.tolist())) == len(lst)
    unlisted = np.ones(len(vals), bool)
    unlisted[lst] = False
    if unlisted.any():
        assert vals[unlisted].max() <= theta                                   # invariant 1
    if np.isfinite(theta):
        assert (vals[lst] >= theta).sum() >= cert                              # invariant 2
    top = np.argsort(-vals, kind="stable")[:cert]
    assert set(top[vals[top] > theta].tolist()) <= set(lst.tolist())           # nothing above theta is missing
    return appends, prunes


@pytest.mark.parametrize("cert", [32, 40, 64])
@pytest.mark.parametrize("kind", ["gauss", "clustered", "ties", "ascending", "descending", "short"])
def test_ladder_invariants(kind, cert):
    rng = np.random.RandomState(zlib.crc32(f"{kind}-{cert}".encode()) % 2**31)
    m = 59_000
    if kind == "gauss":
        v = rng.standard_normal(m) / 22.6
    elif kind == "clustered":
        v = rng.standard_normal(m) * 0.04
        v[rng.choice(m, 60, replace=False)] += 0.6          # the query's own cluster
    elif kind == "ties":
        v = np.round(rng.standard_normal(m), 1) * 0.05      # mass ties at every level
    elif kind == "ascending":
        v = np.sort(rng.standard_normal(m)) * 0.05          # every column beats all earlier ones
    elif kind == "descending":
        v = np.sort(rng.standard_normal(m))[::-1] * 0.05
    else:
        v = rng.standard_normal(1500) * 0.05                # no bootstrap: the list fills up once
    appends, prunes = _check(v.astype(np.float32), cert)
    if kind in ("gauss", "clustered"):
        ideal = cert * (1.0 + np.log(m / BOOT_COLS))
        assert appends <= 2.5 * ideal and prunes == 0, (appends, ideal, prunes)
    if kind == "ascending":
        assert prunes >= 1                                   # adversarial order: handled by the exact reduction
