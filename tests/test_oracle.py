"""CPU tests: the oracle against the reference's golden vectors and known answers."""
import os

import numpy as np
import pytest

from oracle import lemon_oracle as O
from oracle import ref_live
from tests.helpers import clustered_pairs, iid_pairs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
COLS = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")


def _hp(row):
    return dict(zip(O.HP_KEYS, (float(v) for v in row)))


@pytest.mark.parametrize("tag", ["k30", "k5", "k1"])
def test_scoring_matches_reference_golden(tag):
    g = np.load(os.path.join(GOLD, f"scores_{tag}.npz"))
    rec = {c: g[c] for c in COLS + ("d_1",)}
    for h, row in enumerate(g["hparams"]):
        s, dn, dm = O.calc_scores_vectorized(rec, _hp(row))
        for kind in ("vec", "loop"):
            np.testing.assert_allclose(s, g[f"{kind}_scores_{h}"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(dn, g[f"{kind}_dn_{h}"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(dm, g[f"{kind}_dm_{h}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(s, g[f"torch_scores_{h}"], rtol=1e-5, atol=1e-6)
        s2, dn2, dm2 = O.calc_scores_loop(rec, _hp(row))
        np.testing.assert_allclose(s2, s, rtol=1e-12)


def test_normalize_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "normalize.npz"))
    y = O.normalize_vectors(g["x"])
    np.testing.assert_allclose(y, g["y"], rtol=3e-7, atol=1e-30)
    assert np.all(y[3] == 0)


def test_known_answers():
    rng = np.random.RandomState(1)
    n, k = 32, 7
    rec = {c: rng.rand(n, k) for c in COLS}
    rec["d_1"] = rng.rand(n)
    z = dict.fromkeys(O.HP_KEYS, 0.0)
    s, dn, dm = O.calc_scores_vectorized(rec, z)
    np.testing.assert_allclose(s, rec["d_1"])                       # beta=gamma=0 -> score == d_1
    np.testing.assert_allclose(dn, rec["dists_n"].mean(1))          # tau=0 -> s_n == mean(dists_n)
    np.testing.assert_allclose(dm, rec["dists_m"].mean(1))
    hp = dict(z, beta=2.0, gamma=3.0)
    s, _, _ = O.calc_scores_vectorized(rec, hp)
    np.testing.assert_allclose(s, rec["d_1"] + 2 * rec["dists_n"].mean(1) + 3 * rec["dists_m"].mean(1))


def test_self_exclusion_rule():
    D = np.arange(12, dtype=float).reshape(2, 6)
    I = np.arange(12).reshape(2, 6)
    Dk, Ik = O.apply_self_exclusion(D, I, np.array([True, False]))
    assert Ik[0].tolist() == [1, 2, 3, 4, 5]     # in DB: rank 0 dropped
    assert Ik[1].tolist() == [6, 7, 8, 9, 10]    # not in DB: last dropped
    assert Dk.shape == (2, 5)


def test_knn_orthonormal_and_ties():
    d = 16
    db = np.eye(d, dtype=np.float32)
    q = np.zeros((1, d), np.float32)
    q[0, [5, 2, 9]] = [0.9, 0.3, 0.1]
    D, I = O.knn_search(q, db, 3, "ip")
    assert I[0].tolist() == [5, 2, 9]
    np.testing.assert_allclose(D[0], [0.9, 0.3, 0.1], rtol=1e-6)
    # exact duplicates: ties resolve to ascending DB index
    db2 = np.repeat(np.eye(4, dtype=np.float32), 5, axis=0)      # rows 0-4 identical, 5-9 identical ...
    q2 = np.array([[0, 1, 0, 0]], np.float32)
    D2, I2 = O.knn_search(q2, db2, 3, "ip")
    assert I2[0].tolist() == [5, 6, 7]
    # l2: squared, ascending
    D3, I3 = O.knn_search(q2, db2, 6, "l2")
    assert I3[0].tolist() == [5, 6, 7, 8, 9, 0]
    np.testing.assert_allclose(D3[0], [0, 0, 0, 0, 0, 2.0], atol=1e-12)
    # ntotal < k pads with -1 / -inf
    D4, I4 = O.knn_search(q2, db2[:2], 4, "ip")
    assert I4[0].tolist()[2:] == [-1, -1] and np.isinf(D4[0, 2:]).all()


def test_knn_mass_ties_and_chunking_follow_the_total_order():
    """Bit-identical DB rows far beyond the candidate padding, straddling the k-th boundary and DB-chunk borders:
    the result is the full stable sort (best value, then ascending index) whatever the chunking."""
    rng = np.random.RandomState(3)
    db = rng.standard_normal((5000, 24)).astype(np.float32)
    db[100:300] = db[7]            # 201 copies of one row
    q = np.concatenate([db[7:8], db[:6]])
    for metric in ("ip", "l2"):
        q64, db64 = q.astype(np.float64), db.astype(np.float64)
        full = O._det_values(q64, np.broadcast_to(db64, (len(q),) + db64.shape), metric)
        ref = np.argsort(-full if metric == "ip" else full, axis=1, kind="stable")[:, :40]
        for kw in ({}, {"db_chunk": 150}, {"db_chunk": 64, "block": 3}):
            D, I = O.knn_search(q, db, 40, metric, **kw)
            assert (I == ref).all(), (metric, kw)
    assert O.knn_search(q, db, 40, "ip")[1][0, :5].tolist() == [7, 100, 101, 102, 103]


def test_conventions_cosine_and_euclid():
    x, y = iid_pairs(60, 32, seed=2)
    out = O.lemon_oracle(x, y, x, y, k=4, dist_type="cosine", query_in_db=np.arange(60), hparams=O.CC3M_HPARAMS)
    assert (out["D_n"] <= 0.5).all()                       # D = -<a,b>  (run_lemon.py:270,286)
    ip = O.pair_values(x, x, out["I_n"], "ip")
    np.testing.assert_allclose(out["D_n"], -ip, atol=1e-12)
    assert not (out["I_n"] == np.arange(60)[:, None]).any()  # self dropped (rank 0)
    np.testing.assert_allclose(out["dists_m"], 1 - O.pair_values(x, x, out["I_m"], "ip"), atol=1e-12)
    oe = O.lemon_oracle(x, y, x, y, k=4, dist_type="euclidean", query_in_db=np.arange(60), hparams=O.CC3M_HPARAMS)
    assert (oe["D_n"] >= 0).all()                          # squared L2, not negated (run_lemon.py:173,273)
    np.testing.assert_allclose(oe["D_n"], O.pair_values(x, x, oe["I_n"], "l2"), atol=1e-9)
    # unit vectors: L2 ranking == IP ranking
    assert (oe["I_n"] == out["I_n"]).all()
    np.testing.assert_allclose(oe["d_1"], 2 * out["d_1"], atol=1e-6)


def test_discrete_text_metric_skips_negation():
    x, y, lab, _ = clustered_pairs(80, 32, n_clusters=8, seed=4, dup_text_classes=4)
    out = O.lemon_oracle(x, y, x, y, k=5, dist_type="cosine", query_in_db=np.arange(80),
                         text_label_ids_q=lab, text_label_ids_db=lab)
    assert (out["D_n"] > 0).any()                          # +<x_i,x_j>: negation skipped (run_lemon.py:266-270)
    assert set(np.unique(out["dists_n"])) <= {0.0, 1.0}
    assert (out["D_m"] <= 1e-9).all()                      # D_m still negated (:285-286)


def test_reference_cpu_port_matches_oracle():
    x, y, _, _ = clustered_pairs(300, 48, n_clusters=12, seed=5)
    k = 6
    tr_idx = np.arange(300)
    df = O.reference_cpu_scorer(x, y, x, y, k=k, dist_type="cosine", train_indices_in_compr=tr_idx,
                                hparams=O.CC3M_HPARAMS)
    out = O.lemon_oracle(x, y, x, y, k=k, dist_type="cosine", query_in_db=np.arange(300), hparams=O.CC3M_HPARAMS)
    I_n = np.stack(df["I_n"].values)
    same = np.array([set(a) == set(b) for a, b in zip(I_n, out["I_n"])])
    assert same.mean() > 0.98                              # fp32 topk vs f64: only eps-ties may differ
    sc = df["score"].values
    np.testing.assert_allclose(sc[same], out["score"][same], rtol=1e-5)
    # val/test-style split: no exclusion, search k
    df2 = O.reference_cpu_scorer(x[:40], y[:40], x, y, k=k, dist_type="euclidean", hparams=O.CC3M_HPARAMS)
    o2 = O.lemon_oracle(x[:40], y[:40], x, y, k=k, dist_type="euclidean", hparams=O.CC3M_HPARAMS)
    np.testing.assert_allclose(df2["score"].values, o2["score"], rtol=2e-5)


def test_compare_neighbor_sets_tie_rule():
    db = np.eye(8, dtype=np.float32)
    db[3] = db[2]                                          # exact duplicate rows 2,3
    q = np.array([[0.5, 0.4, 0.3, 0, 0, 0, 0, 0]], np.float32)
    _, I = O.knn_search(q, db, 3, "ip")
    assert I[0].tolist() == [0, 1, 2]
    r = O.compare_neighbor_sets(q, db, np.array([[0, 1, 3]]), 3)   # picked the other duplicate
    assert r["tie_excused"] == 1 and r["wrong"] == 0
    r = O.compare_neighbor_sets(q, db, np.array([[0, 1, 5]]), 3)
    assert r["wrong"] == 1


@pytest.mark.skipif(not ref_live.available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_vs_live_reference_on_oracle_records():
    import pandas as pd
    mu = ref_live.import_reference_metrics()
    x, y, _, _ = clustered_pairs(200, 64, n_clusters=10, seed=6, noise_frac=0.3)
    out = O.lemon_oracle(x, y, x, y, k=8, dist_type="cosine", query_in_db=np.arange(200), hparams=O.CC3M_HPARAMS)
    df = pd.DataFrame([{**{c: out[c][i].astype(np.float32) for c in COLS}, "d_1": float(out["d_1"][i])}
                       for i in range(200)])
    s, dn, dm = mu.calc_scores_given_hparams_vectorized(df, O.CC3M_HPARAMS, return_dn=True)
    np.testing.assert_allclose(out["score"], s, rtol=1e-5)
    np.testing.assert_allclose(out["s_n"], dn, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["s_m"], dm, rtol=1e-5, atol=1e-7)
