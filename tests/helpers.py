"""Seeded synthetic inputs shared by CPU and GPU tests (SURVEY.md §8d geometry)."""
import numpy as np


def clustered_pairs(n, d, n_clusters=50, seed=0, noise_frac=0.0, dup_text_classes=0):
    """Clustered unit-norm image/text embeddings mimicking CLIP geometry.
    noise_frac: fraction of captions replaced by another caption of the same cluster
    (exact duplicate text rows -> ties).  dup_text_classes>0: text side drawn from only
    that many distinct vectors (classification datasets); returns label ids too."""
    rng = np.random.RandomState(seed)
    c = rng.standard_normal((n_clusters, d))
    c2 = rng.standard_normal((n_clusters, d))
    z = rng.randint(0, n_clusters, n)
    x = c[z] + 0.6 * rng.standard_normal((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if dup_text_classes:
        protos = rng.standard_normal((dup_text_classes, d))
        lab = z % dup_text_classes
        y = protos[lab]
    else:
        lab = None
        y = 0.5 * x * np.sqrt(d) / 1.0 + 0.5 * c2[z] + 0.6 * rng.standard_normal((n, d))
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    mislabel = np.zeros(n, bool)
    if noise_frac > 0:
        idx = rng.choice(n, int(noise_frac * n), replace=False)
        for i in idx:
            peers = np.nonzero(z == z[i])[0]
            peers = peers[peers != i]
            if len(peers):
                y[i] = y[rng.choice(peers)]
                mislabel[i] = True
    return x.astype(np.float32), y.astype(np.float32), lab, mislabel


def iid_pairs(n, d, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((n, d)); x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = rng.standard_normal((n, d)); y /= np.linalg.norm(y, axis=1, keepdims=True)
    return x.astype(np.float32), y.astype(np.float32)


def hparam_search_case(tag):
    """Small LEMoN-like validation DataFrame + search arguments (run_lemon.py:331-394 in miniature) shared by
    tests/golden/make_golden_hparam_search.py and tests/test_hparam_search.py."""
    import pandas as pd
    from oracle import lemon_oracle as O
    n, k = 240, 6
    x, y, _, mis = clustered_pairs(n, 40, n_clusters=8, seed=77, noise_frac=0.35)
    o = O.lemon_oracle(x, y, x, y, k=k, query_in_db=np.arange(n))
    cols = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")
    rec = {c: o[c].astype(np.float32) for c in cols}
    d1 = o["d_1"].astype(np.float32).astype(np.float64)
    df = pd.DataFrame([{**{c: rec[c][i] for c in cols}, "d_1": float(d1[i]), "is_mislabel": int(mis[i])} for i in range(n)])
    grid = {"beta": np.arange(0, 20.01, 10), "gamma": np.arange(0, 20.01, 10), "tau_1": [0, 5], "tau_2": [0, 5]}
    x0s = [[0] * 6, [0.5] * 6]
    if tag == "ablate":          # run_lemon.py:364-377: --ablation d1_gamma
        return df, grid, x0s, ["gamma"], ["beta"]
    return df, grid, x0s, [], []


KNN_PIN_CASES = ("a", "b", "c", "u", "t")


def knn_pin_case(tag):
    """Seeded inputs of the independent kNN answers in tests/golden/knn.npz (make_golden_knn.py).
    Returns (db, q, k, kind); numpy's legacy RandomState stream is frozen, so the arrays are reproducible."""
    rng = np.random.RandomState({"a": 101, "b": 102, "c": 103, "u": 104, "t": 105}[tag])
    unit = lambda a: (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)
    if tag in ("a", "b", "c"):
        m, nq, d, k = {"a": (3000, 96, 64, 31), "b": (2500, 64, 96, 30), "c": (4096, 48, 512, 51)}[tag]
        cen = rng.standard_normal((20, d))
        db = unit(cen[rng.randint(0, 20, m)] + 0.7 * rng.standard_normal((m, d)))
        q = unit(db[rng.choice(m, nq, replace=False)] + 0.05 * rng.standard_normal((nq, d)))
        return db, q, k, "unit"
    if tag == "u":
        m, nq, d, k = 2200, 80, 128, 31
        db = (rng.standard_normal((m, d)) * rng.uniform(0.5, 3.0, (m, 1)) / np.sqrt(d)).astype(np.float32)
        q = (rng.standard_normal((nq, d)) * rng.uniform(0.5, 3.0, (nq, 1)) / np.sqrt(d)).astype(np.float32)
        return db, q, k, "raw"
    base = unit(rng.standard_normal((40, 64)))
    db = np.repeat(base, 25, axis=0)[rng.permutation(1000)]     # every vector 25 times, shuffled: mass ties
    return db, base[:16].copy(), 31, "ties"


def check_against_oracle(out, xq, yq, xdb, ydb, *, k, dist_type="cosine", query_in_db=None, hparams=None,
                         lab_q=None, lab_db=None, eps_tie=None, rtol=1e-5, normalize=True, strict=True):
    """Acceptance check of SURVEY.md §8c for a score_pairs() result `out` (numpy arrays):
    neighbour SETS equal the float64 oracle's modulo eps-ties at the boundary; record arrays and
    scores within `rtol` of the oracle — rows whose sets differ by an excused tie are compared
    with the oracle re-evaluated on the returned index sets.  Returns counts.  strict=False (bench.py's
    parity gate) counts wrong rows / mismatching values in stats["wrong"] instead of raising."""
    from oracle import lemon_oracle as O
    metric = "ip" if dist_type == "cosine" else "l2"
    if eps_tie is None:
        eps_tie = O.EPS_TIE_COSINE * (1 if metric == "ip" else 2)
    if normalize:
        xq, yq, xdb, ydb = (O.normalize_vectors(a) for a in (xq, yq, xdb, ydb))
    ref = O.lemon_oracle(xq, yq, xdb, ydb, k=k, dist_type=dist_type, query_in_db=query_in_db, hparams=hparams,
                         normalize=False, text_label_ids_q=lab_q, text_label_ids_db=lab_db)
    N = xq.shape[0]
    stats = {"wrong": 0}
    tie_rows = np.zeros(N, bool)
    bad_rows = np.zeros(N, bool)

    def close(a, b, atol, what):
        """assert_allclose, or (strict=False) mark the rows that violate it"""
        if strict:
            np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)
            return
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        viol = ~(np.abs(a - b) <= atol + rtol * np.abs(b))
        viol |= np.isnan(a) != np.isnan(b)
        return viol.any(axis=tuple(range(1, viol.ndim))) if viol.ndim > 1 else viol
    for side, q, db in (("n", xq, xdb), ("m", yq, ydb)):
        I_got = out[f"I_{side}"]
        assert I_got.shape == (N, k) and I_got.dtype == np.int64
        # kNN acceptance is defined on the searched list (k or k+1); with self-exclusion compare the kept k
        D_ref = O.pair_values(q, db, ref[f"I_{side}"], metric)
        # boundary value: the worst kept neighbour of the oracle
        top = None
        if query_in_db is not None:   # value of the dropped rank-0 entry for rows that are in the DB
            raw_D = ref["raw"][0 if side == "n" else 2]
            top = np.where(np.asarray(query_in_db) >= 0, raw_D[:, 0], np.nan)
        r = O.compare_neighbor_sets(q, db, I_got, k, metric, eps_tie=eps_tie, D_ref=D_ref, I_ref=ref[f"I_{side}"],
                                    top_boundary=top)
        if strict:
            assert r["wrong"] == 0, f"side {side}: {r['wrong']} rows with wrong neighbour sets, e.g. {r['wrong_rows'][:5]}"
        bad_rows[r["wrong_rows"]] = True
        stats[f"exact_{side}"], stats[f"tie_excused_{side}"] = r["exact"], r["tie_excused"]
        tie_rows[r["excused_rows"]] = True
    # rows with tie-excused sets: oracle re-evaluated on the returned sets
    ref2 = O.lemon_oracle(xq, yq, xdb, ydb, k=k, dist_type=dist_type, hparams=hparams, normalize=False,
                          text_label_ids_q=lab_q, text_label_ids_db=lab_db, given_I=(out["I_n"], out["I_m"]))
    if query_in_db is not None:
        pass  # given_I lists are already self-excluded
    cols = ("D_n", "dists_n", "dists_tr_n", "D_m", "dists_m", "dists_tr_m")
    for c in cols:
        side = c[-1]
        # align by neighbour id (near-ties may permute the order inside a row)
        og = np.argsort(out[f"I_{side}"], axis=1, kind="stable")
        orf = np.argsort(ref[f"I_{side}"], axis=1, kind="stable")
        a = np.take_along_axis(out[c].astype(np.float64), og, axis=1)
        b = np.take_along_axis(ref[c], orf, axis=1)
        ok = ~tie_rows & ~bad_rows
        tr = tie_rows & ~bad_rows
        v1 = close(a[ok], b[ok], 2e-6, c)
        v2 = close(out[c][tr].astype(np.float64), ref2[c][tr], 2e-6, c + " (tie rows)")
        if not strict:
            bad_rows[np.nonzero(ok)[0][v1]] = True
            bad_rows[np.nonzero(tr)[0][v2]] = True
    v = close(out["d_1"], ref["d_1"], 2e-6, "d_1")
    if not strict:
        bad_rows |= v
    if hparams is not None:
        for c in ("s_n", "s_m", "score"):
            ok = ~tie_rows & ~bad_rows
            tr = tie_rows & ~bad_rows
            v1 = close(out[c][ok], ref[c][ok], 1e-6, c)
            v2 = close(out[c][tr], ref2[c][tr], 1e-6, c + " (tie rows)")
            if not strict:
                bad_rows[np.nonzero(ok)[0][v1]] = True
                bad_rows[np.nonzero(tr)[0][v2]] = True
    stats["tie_rows"] = int(tie_rows.sum())
    stats["wrong"] = int(bad_rows.sum())
    stats["wrong_rows"] = np.nonzero(bad_rows)[0][:16].tolist()
    return stats
