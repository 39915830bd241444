"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
liblemon_b200.so; the oracle is only the checker."""
import os

import numpy as np
import pytest

from tests.helpers import clustered_pairs, iid_pairs, check_against_oracle
from lemon_b200.scoring import count_uncertified

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
HP = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}
COLS = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")


@pytest.fixture(scope="module")
def lb():
    import torch
    import lemon_b200
    assert torch.cuda.is_available()
    assert os.path.exists(lemon_b200.LIB_PATH), "liblemon_b200.so must be built in-tree"
    return lemon_b200


def _np(out):
    return {k: v.cpu().numpy() for k, v in out.items()}


def test_normalize_matches_reference_golden(lb):
    g = np.load(os.path.join(GOLD, "normalize.npz"))
    sc = lb.get_scorer()
    p = sc.prepare(g["x"], normalize=True)
    y = p.f32.cpu().numpy()
    np.testing.assert_allclose(y, g["y"], rtol=5e-7, atol=1e-30)
    assert (y[3] == 0).all()
    # fp16 operand copy + statistics
    y16 = p.f16.cpu().numpy()[:, : y.shape[1]]
    assert (y16 == y.astype(np.float16)).all()
    st = p.row_stats.cpu().numpy()
    np.testing.assert_allclose(st[:, 2], np.linalg.norm(y - y16.astype(np.float32), axis=1), rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(st[:, 0], np.linalg.norm(y, axis=1), rtol=1e-5)


@pytest.mark.parametrize("tag", ["k30", "k5", "k1"])
def test_combine_scores_matches_reference_golden(lb, tag):
    import pandas as pd
    from lemon_b200 import metrics_compat
    g = np.load(os.path.join(GOLD, f"scores_{tag}.npz"))
    n = len(g["d_1"])
    df = pd.DataFrame([{**{c: g[c][i] for c in COLS}, "d_1": float(g["d_1"][i])} for i in range(n)])
    for h, row in enumerate(g["hparams"]):
        hp = dict(zip(lb.HP_KEYS, (float(v) for v in row)))
        s, dn, dm = metrics_compat.calc_scores_given_hparams_vectorized(df, hp, return_dn=True)
        assert isinstance(s, np.ndarray) and s.dtype == np.float64 and s.shape == (n,)
        np.testing.assert_allclose(s, g[f"vec_scores_{h}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dn, g[f"vec_dn_{h}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(dm, g[f"vec_dm_{h}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(s, g[f"loop_scores_{h}"], rtol=1e-5, atol=1e-6)
        s2 = metrics_compat.calc_scores_given_hparams_vectorized(df, hp)
        assert (s2 == s).all()


@pytest.mark.parametrize("d,k", [(64, 1), (96, 5), (512, 30), (768, 50), (50, 7)])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_knn_exact_kernel_vs_oracle(lb, d, k, metric):
    from oracle import lemon_oracle as O
    x, _ = iid_pairs(1500, d, seed=d + k)
    q = x[:333] + 0.05 * iid_pairs(333, d, seed=7)[0]
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(q, normalize=False), sc.prepare(x, normalize=False)
    tv, ti = sc.knn(qp, dbp, k, 0 if metric == "ip" else 1, mode="exact")
    tv, ti = tv.cpu().numpy(), ti.cpu().numpy().astype(np.int64)
    D, I = O.knn_search(q, x, k, metric)
    r = O.compare_neighbor_sets(q, x, ti, k, metric, eps_tie=4e-6, D_ref=D, I_ref=I)
    assert r["wrong"] == 0
    assert r["exact"] >= 0.99 * r["rows"]
    same = (ti == I).all(axis=1)
    np.testing.assert_allclose(tv[same], D[same], rtol=1e-5, atol=2e-6)
    # sortedness (best first)
    assert ((np.diff(tv, axis=1) <= 0) if metric == "ip" else (np.diff(tv, axis=1) >= 0)).all()


def test_knn_exact_ties_resolve_by_index(lb):
    from oracle import lemon_oracle as O
    base, _ = iid_pairs(40, 64, seed=11)
    db = np.repeat(base, 25, axis=0)            # 1000 rows, every vector 25 times
    perm = np.random.RandomState(0).permutation(len(db))
    db = db[perm]
    q = base[:16]
    sc = lb.get_scorer()
    tv, ti = sc.knn(sc.prepare(q, False), sc.prepare(db, False), 31, 0, mode="exact")
    D, I = O.knn_search(q, db, 31, "ip")
    assert (ti.cpu().numpy() == I).all()        # exact duplicates: identical order, ascending DB index


def test_knn_exact_small_db_pads_like_faiss(lb):
    x, _ = iid_pairs(5, 64, seed=3)
    sc = lb.get_scorer()
    tv, ti = sc.knn(sc.prepare(x, False), sc.prepare(x[:3], False), 6, 0, mode="exact")
    ti, tv = ti.cpu().numpy(), tv.cpu().numpy()
    assert (ti[:, 3:] == -1).all() and np.isneginf(tv[:, 3:]).all() and (ti[:, :3] >= 0).all()


@pytest.mark.parametrize("dist_type", ["cosine", "euclidean"])
@pytest.mark.parametrize("train", [True, False])
def test_score_pairs_exact_mode_vs_oracle(lb, dist_type, train):
    x, y, _, _ = clustered_pairs(1200, 128, n_clusters=24, seed=21, noise_frac=0.3)
    k = 10
    raw = lambda a: (a * np.random.RandomState(5).uniform(0.5, 3.0, (a.shape[0], 1))).astype(np.float32)
    xr, yr = raw(x), raw(y)                      # un-normalised inputs: exercises K0
    if train:
        out = _np(lb.score_pairs(xr, yr, k=k, dist_type=dist_type, query_in_db=np.arange(1200), hparams=HP,
                                 knn_mode="exact"))
        st = check_against_oracle(out, xr, yr, xr, yr, k=k, dist_type=dist_type, query_in_db=np.arange(1200), hparams=HP)
        assert not (out["I_n"] == np.arange(1200)[:, None]).any()
    else:
        out = _np(lb.score_pairs(xr[:300], yr[:300], xr[300:], yr[300:], k=k, dist_type=dist_type, hparams=HP,
                                 knn_mode="exact"))
        st = check_against_oracle(out, xr[:300], yr[:300], xr[300:], yr[300:], k=k, dist_type=dist_type, hparams=HP)
    assert st["exact_n"] + st["tie_excused_n"] == out["I_n"].shape[0]


def test_score_pairs_query_not_in_db_drops_last(lb):
    x, y, _, _ = clustered_pairs(600, 64, n_clusters=12, seed=22)
    qid = np.arange(600)
    qid[::3] = -1                                 # a third of the train queries were not sampled into the DB
    out = _np(lb.score_pairs(x, y, x, y, k=6, query_in_db=qid, hparams=HP, knn_mode="exact"))
    check_against_oracle(out, x, y, x, y, k=6, query_in_db=qid, hparams=HP)
    # rows with qid == -1 keep rank 0 (the sample itself, since it IS physically in this DB)
    assert (out["I_n"][::3, 0] == np.arange(600)[::3]).all()


def test_score_pairs_discrete_text_metric(lb):
    x, y, lab, _ = clustered_pairs(900, 64, n_clusters=20, seed=23, dup_text_classes=10)
    out = _np(lb.score_pairs(x, y, k=8, query_in_db=np.arange(900), hparams=HP, text_label_ids_q=lab,
                             text_label_ids_db=lab, knn_mode="exact"))
    check_against_oracle(out, x, y, x, y, k=8, query_in_db=np.arange(900), hparams=HP, lab_q=lab, lab_db=lab)
    assert (out["D_n"] > 0).any() and set(np.unique(out["dists_n"])) <= {0.0, 1.0}


def test_faiss_compat_api(lb):
    from lemon_b200 import faiss_compat as faiss
    from oracle import lemon_oracle as O
    x, _ = iid_pairs(700, 96, seed=31)
    q = x[:50] * 1.7
    for cls, metric in ((faiss.IndexFlatIP, "ip"), (faiss.IndexFlatL2, "l2")):
        index = cls(96)
        index.add(x[:400]); index.add(x[400:])
        assert index.ntotal == 700
        D, I = index.search(q, 9)
        assert isinstance(D, np.ndarray) and D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (50, 9)
        Dr, Ir = O.knn_search(q, x, 9, metric)
        r = O.compare_neighbor_sets(q, x, I, 9, metric, eps_tie=4e-6, D_ref=Dr, I_ref=Ir)
        assert r["wrong"] == 0
        np.testing.assert_allclose(D[(I == Ir).all(1)], Dr[(I == Ir).all(1)], rtol=1e-5, atol=3e-6)
    small = faiss.IndexFlatIP(96); small.add(x[:4])
    D, I = small.search(q[:2], 6)
    assert (I[:, 4:] == -1).all() and np.isneginf(D[:, 4:]).all()
    with pytest.raises(RuntimeError):
        small.add(x[:, :50])


# ------------------------------------------------------------------ tensor-core (tcgen05) path
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("nq,m,d,nseg", [(300, 5000, 128, 1), (1000, 20000, 512, 3), (700, 9000, 768, 1),
                                         (128, 256, 64, 1), (257, 3001, 200, 2), (5, 70, 512, 1)])
def test_tc_candidates_match_fp16_matmul(lb, cg, nq, m, d, nseg):
    """The fused kernel's candidate lists contain the top-64 of (fp16 operands, fp32 accumulate), every reported
    value is that pair's product, and every column outside the lists is below the published threshold."""
    import torch
    from lemon_b200.scoring import decode_candidates
    x, _, _, _ = clustered_pairs(m, d, n_clusters=max(4, m // 100), seed=m)
    q = iid_pairs(nq, d, seed=nq)[0] * 0.2 + x[np.arange(nq) % m]
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(q, True), sc.prepare(x, True)
    ck, cc, ct, nseg_out = sc.knn_candidates(qp, dbp, nseg=nseg, cta_group=cg)
    assert nseg_out == nseg and ck.shape[1:] == (2 * nseg, 1024) and ck.shape[0] % 256 == 0
    S = (qp.f16.float() @ dbp.f16.float().T).cpu().numpy()
    cv, ci = decode_candidates(ck, cc, nq)
    valid = ci >= 0
    assert (ci < m).all() and (cc[:nq].cpu().numpy() <= 1024).all()
    for r in (0, nq // 2, nq - 1):                                # no duplicate columns inside a row
        v = ci[r][valid[r]]
        assert len(np.unique(v)) == len(v)
    got = np.take_along_axis(S, np.where(valid, ci, 0), 1)
    np.testing.assert_allclose(cv[valid], got[valid], rtol=0, atol=3e-5)   # reported value == that pair's product
    kk = min(64, m)
    tv = -np.sort(-S, axis=1)[:, :kk]
    np.testing.assert_allclose(cv[:, :kk], tv, rtol=0, atol=3e-5)          # union of the lists holds the global top-64
    if m <= 64:
        assert (valid.sum(1) == m).all()
    th = ct[:nq].max(dim=1).values.cpu().numpy()                  # columns in no list are <= the lists' thresholds
    Sn = S.copy()
    np.put_along_axis(Sn, np.where(valid, ci, 0), -np.inf, 1)
    assert (Sn.max(1) <= th + 3e-5).all() or m <= 64


@pytest.mark.parametrize("normalize", [True, False])
def test_tc_error_bound_is_rigorous(lb, normalize):
    """|fp16 tensor-core inner product - float64 inner product| <= eps_row used by the certificate; with
    un-normalised rows (norms 0.5 .. 6) the accumulation term scales with the norms."""
    from lemon_b200.scoring import acc_eps_coef
    x, y, _, _ = clustered_pairs(6000, 768, n_clusters=40, seed=41)
    if not normalize:
        x = (x * np.random.RandomState(3).uniform(0.5, 6.0, (len(x), 1))).astype(np.float32)
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(x[:900], normalize), sc.prepare(x, normalize)
    from lemon_b200.scoring import decode_candidates
    ck, cc, ct, _ = sc.knn_candidates(qp, dbp, nseg=1)
    cv, ci = decode_candidates(ck, cc, 900)
    cv, ci = cv[:, :64].astype(np.float64), ci[:, :64]
    assert (ci >= 0).all()
    q64, db64 = qp.f32.cpu().numpy().astype(np.float64), dbp.f32.cpu().numpy().astype(np.float64)
    exact = np.einsum("nd,nkd->nk", q64, db64[ci])
    rs, smax = qp.row_stats.cpu().numpy(), dbp.stats_max.cpu().numpy()
    eps = rs[:, 2] * smax[1] + rs[:, 0] * smax[2] + acc_eps_coef(768, 768) * np.maximum(rs[:, 0], rs[:, 1]) * max(smax[0], smax[1])
    err = np.abs(cv - exact).max(axis=1)
    assert (err <= eps).all(), (err.max(), eps.min())
    assert (err < 0.25 * eps).all()              # the bound is comfortably loose, not marginal


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("d,kp", [(512, 31), (768, 51), (96, 6)])
def test_knn_tc_path_equals_exact_path_bitwise(lb, metric, d, kp):
    """Tensor-core candidates + fp32 re-rank + certificate/fallback give the SAME lists, bit for bit,
    as the fp32 brute-force kernel (both report values through the same fp32 summation order)."""
    x, y, _, _ = clustered_pairs(9000, d, n_clusters=60, seed=d, noise_frac=0.2)
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(y[:1500], True), sc.prepare(y, True)      # text side: contains exact duplicates
    tv, ti = sc.knn(qp, dbp, kp, metric, mode="tc")
    info = dict(sc.last_info)
    ev, ei = sc.knn(qp, dbp, kp, metric, mode="exact")
    assert (ti == ei).all()
    assert (tv == ev).all()
    assert count_uncertified(info) < 0.2 * 1500


def test_knn_tc_mass_duplicates_fall_back_exactly(lb):
    """Classification-style text side (10 distinct vectors): ties span far beyond 64 candidates, every
    row is uncertified and must be served by the exact GPU fallback with index-ascending ties."""
    from oracle import lemon_oracle as O
    x, y, lab, _ = clustered_pairs(3000, 128, n_clusters=30, seed=51, dup_text_classes=10)
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(y[:200], True), sc.prepare(y, True)
    tv, ti = sc.knn(qp, dbp, 31, 0, mode="tc")
    assert count_uncertified(sc.last_info) == 200
    D, I = O.knn_search(qp.f32.cpu().numpy(), dbp.f32.cpu().numpy(), 31, "ip")
    assert (ti.cpu().numpy() == I).all()


@pytest.mark.parametrize("dist_type", ["cosine", "euclidean"])
def test_score_pairs_tc_mode_vs_oracle(lb, dist_type):
    x, y, _, mis = clustered_pairs(6000, 512, n_clusters=40, seed=61, noise_frac=0.4)
    k = 30
    out = _np(lb.score_pairs(x, y, k=k, dist_type=dist_type, query_in_db=np.arange(6000), hparams=HP, knn_mode="tc"))
    st = check_against_oracle(out, x, y, x, y, k=k, dist_type=dist_type, query_in_db=np.arange(6000), hparams=HP)
    assert st["exact_n"] + st["tie_excused_n"] == 6000
    # sanity: the score separates the injected caption noise (not a parity claim)
    from sklearn.metrics import roc_auc_score
    assert roc_auc_score(mis, out["score"]) > 0.7


def test_score_pairs_tc_val_split_and_768(lb):
    x, y, _, _ = clustered_pairs(5000, 768, n_clusters=50, seed=62)
    out = _np(lb.score_pairs(x[:700], y[:700], x[700:], y[700:], k=15, hparams=HP, knn_mode="tc"))
    check_against_oracle(out, x[:700], y[:700], x[700:], y[700:], k=15, hparams=HP)


def test_full_size_properties_c1_shape(lb):
    """BASELINE config-1 shape (50k x 512, k=30) through the default path: size-independent properties."""
    import torch
    x, y, _, _ = clustered_pairs(50000, 512, n_clusters=1000, seed=71, noise_frac=0.4)
    n = 50000
    out = lb.score_pairs(x, y, k=30, query_in_db=np.arange(n), hparams=HP)
    torch.cuda.synchronize()
    I_n, I_m = out["I_n"], out["I_m"]
    ar = torch.arange(n, device=I_n.device)[:, None]
    assert int((I_n == ar).sum()) == 0                            # self excluded on the image side (no duplicates there)
    assert bool((I_n >= 0).all()) and bool((I_n < n).all()) and bool((I_m >= 0).all())
    # D_n = -<x_i,x_j> ascending (best first); dists_m = 1 - <x_i, x_m> in [0,2]
    assert bool((out["D_n"][:, 1:] >= out["D_n"][:, :-1]).all())
    assert bool((out["D_m"][:, 1:] >= out["D_m"][:, :-1]).all())
    assert bool(((out["dists_m"] > -1e-5) & (out["dists_m"] < 2 + 1e-5)).all())
    # no duplicate neighbours per row
    srt = torch.sort(I_n, dim=1).values
    assert int((srt[:, 1:] == srt[:, :-1]).sum()) == 0
    # idempotence / determinism: a second run is bit-identical
    out2 = lb.score_pairs(x, y, k=30, query_in_db=np.arange(n), hparams=HP)
    for c in ("score", "I_n", "I_m", "D_n", "dists_n"):
        assert bool((out[c] == out2[c]).all()), c
    # score == d_1 + beta*s_n + gamma*s_m and matches the oracle's formula on the returned records
    from oracle import lemon_oracle as O
    sub = slice(0, 4000)
    rec = {c: out[c][sub].cpu().numpy() for c in ("d_1", "D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")}
    s, _, _ = O.calc_scores_vectorized(rec, HP)
    np.testing.assert_allclose(out["score"][sub].cpu().numpy(), s, rtol=1e-6)
    # neighbour sets of a sample of rows against the float64 oracle
    rows = np.arange(0, n, 97)[:400]
    xn = O.normalize_vectors(x)
    D, I = O.knn_search(xn[rows], xn, 31, "ip")
    r = O.compare_neighbor_sets(xn[rows], xn, I_n[rows].cpu().numpy(), 30, "ip", D_ref=D[:, 1:], I_ref=I[:, 1:],
                                top_boundary=D[:, 0])
    assert r["wrong"] == 0


@pytest.mark.parametrize("dist_type", ["cosine", "euclidean"])
def test_normalize_d1_variant(lb, dist_type):
    """--normalize_d1 (run_lemon.py:244-248): d_1 = softmax over class prompts read at the noisy label."""
    from oracle import lemon_oracle as O
    x, y, lab, _ = clustered_pairs(700, 64, n_clusters=20, seed=81, dup_text_classes=10)
    protos = np.stack([y[np.nonzero(lab == c)[0][0]] for c in range(10)]) * 1.3     # un-normalised class prompts
    noisy = (lab + (np.arange(700) % 3 == 0)) % 10
    out = _np(lb.score_pairs(x, y, k=6, dist_type=dist_type, query_in_db=np.arange(700), hparams=HP, knn_mode="exact",
                             class_text_emb=protos, noisy_label=noisy, text_label_ids_q=lab, text_label_ids_db=lab))
    ref = O.lemon_oracle(x, y, x, y, k=6, dist_type=dist_type, query_in_db=np.arange(700), hparams=HP,
                         class_text_emb=protos, noisy_label=noisy, text_label_ids_q=lab, text_label_ids_db=lab)
    np.testing.assert_allclose(out["d_1"], ref["d_1"], rtol=1e-5, atol=1e-7)
    assert ((out["d_1"] > 0) & (out["d_1"] < 1)).all()
    same = (out["I_n"] == ref["I_n"]).all(1) & (out["I_m"] == ref["I_m"]).all(1)
    np.testing.assert_allclose(out["score"][same], ref["score"][same], rtol=1e-5)


@pytest.mark.parametrize("seed,n,d,k", [(1, 1500, 512, 1), (2, 1999, 768, 2), (3, 777, 512, 5), (4, 2048, 768, 10),
                                        (5, 1300, 512, 15), (6, 1800, 768, 20), (7, 900, 512, 30), (8, 2000, 768, 50)])
def test_property_grid_default_path_vs_oracle(lb, seed, n, d, k):
    """SURVEY.md §4 property grid: random N <= 2k, d in {512,768}, k in the reference's k grid
    (experiments.py:86); default (tensor-core) path; sets == float64 oracle modulo eps-ties, scores <= 1e-5."""
    x, y, _, _ = clustered_pairs(n, d, n_clusters=max(4, n // 60), seed=100 + seed, noise_frac=0.25)
    out = _np(lb.score_pairs(x, y, k=k, query_in_db=np.arange(n), hparams=HP))
    st = check_against_oracle(out, x, y, x, y, k=k, query_in_db=np.arange(n), hparams=HP)
    assert st["exact_n"] + st["tie_excused_n"] == n and st["exact_m"] + st["tie_excused_m"] == n


def test_results_adapter_feeds_metrics_compat(lb):
    """fused path -> legacy DataFrame (run_lemon.py:291-314 schema) -> drop-in scoring function: same scores."""
    from lemon_b200 import results, metrics_compat
    x, y, _, mis = clustered_pairs(1200, 128, n_clusters=20, seed=91, noise_frac=0.3)
    out = lb.score_pairs(x, y, k=9, query_in_db=np.arange(1200), hparams=HP)
    df = results.records_to_dataframe(out, "train", is_mislabel=mis)
    s = metrics_compat.calc_scores_given_hparams_vectorized(df, HP)
    np.testing.assert_allclose(s, out["score"].cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("kind", ["ten_classes", "dup25"])
def test_tc_long_db_with_mass_ties(lb, kind):
    """DB long enough for K1's threshold bootstrap (>= 64 tiles) and full of exact duplicates: the bootstrap
    threshold must not lose the tied columns; results equal the fp32 brute-force kernel bit for bit."""
    if kind == "ten_classes":
        x, y, lab, _ = clustered_pairs(20000, 128, n_clusters=40, seed=95, dup_text_classes=10)
        mat = y
    else:
        base, _ = iid_pairs(800, 128, seed=96)
        mat = np.repeat(base, 25, axis=0)[np.random.RandomState(1).permutation(20000)]
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(mat[:600], True), sc.prepare(mat, True)
    ck, cc, ct, nseg = sc.knn_candidates(qp, dbp)
    assert int(cc[:600].sum(1).min()) >= 64              # at least 64 real candidates per row, nothing lost
    tv, ti = sc.knn(qp, dbp, 31, 0, mode="tc")
    ev, ei = sc.knn(qp, dbp, 31, 0, mode="exact")
    assert bool((ti == ei).all()) and bool((tv == ev).all())


def test_duplicate_rows_are_searched_once_and_expanded_exactly(lb):
    """Classification-style DB (10 distinct text rows) and caption-noise duplicates: the deduplicated search
    returns bit-for-bit what the plain search returns, and both match the oracle."""
    import torch
    x, y, lab, _ = clustered_pairs(6000, 128, n_clusters=30, seed=97, dup_text_classes=10)
    n = 6000
    plain = lb.LemonScorer(dedup=False)
    dd = lb.LemonScorer(dedup=True)
    outs = []
    for sc in (plain, dd):
        sc.set_database(x, y, "cosine", True, lab)
        outs.append(_np(sc.score(None, None, k=30, query_in_db=np.arange(n), hparams=HP, text_label_ids_q=lab,
                                 queries_are_db=True)))
    assert dd.db["y"].dedup is not None and dd.db["y"].dedup.n_unique == 10 and dd.db["x"].dedup is None
    for c in outs[0]:
        assert (outs[0][c] == outs[1][c]).all(), c
    check_against_oracle(outs[1], x, y, x, y, k=30, query_in_db=np.arange(n), hparams=HP, lab_q=lab, lab_db=lab)
    # caption-noise duplicates (pairs / triples of identical rows) on the general path
    x2, y2, _, _ = clustered_pairs(5000, 512, n_clusters=40, seed=98, noise_frac=0.4)
    a = _np(plain.set_database(x2, y2).score(None, None, k=10, query_in_db=np.arange(5000), hparams=HP, queries_are_db=True))
    b = _np(dd.set_database(x2, y2).score(None, None, k=10, query_in_db=np.arange(5000), hparams=HP, queries_are_db=True))
    assert dd.db["y"].dedup is not None and dd.db["y"].dedup.n_unique < 4600
    for c in a:
        assert (a[c] == b[c]).all(), c


def test_faiss_compat_unnormalised_vectors_on_the_tensor_core_path(lb):
    """faiss semantics do not normalise: vectors of assorted norms, DB large enough for K1.  The certificate uses
    the rows' actual norms; whatever it cannot certify (all of L2 here: the norms are far from 1) goes through the
    exact GPU fallback — results must equal the oracle either way."""
    from lemon_b200 import faiss_compat as faiss
    from oracle import lemon_oracle as O
    rng = np.random.RandomState(5)
    x, _, _, _ = clustered_pairs(6000, 96, n_clusters=40, seed=33)
    x = (x * rng.uniform(0.3, 2.5, (6000, 1))).astype(np.float32)
    q = (x[:400] + 0.02 * rng.standard_normal((400, 96))).astype(np.float32)
    for cls, metric in ((faiss.IndexFlatIP, "ip"), (faiss.IndexFlatL2, "l2")):
        index = cls(96)
        index.add(x)
        D, I = index.search(q, 12)
        Dr, Ir = O.knn_search(q, x, 12, metric)
        r = O.compare_neighbor_sets(q, x, I, 12, metric, eps_tie=2e-5, D_ref=Dr, I_ref=Ir)
        assert r["wrong"] == 0, (metric, r["wrong_rows"][:5])
        same = (I == Ir).all(1)
        assert same.mean() > 0.98
        np.testing.assert_allclose(D[same], Dr[same], rtol=2e-5, atol=1e-5)


def test_tail_launch_and_segments_equal_exact_path(lb):
    """Row count just above one full round of CTA pairs: the scorer sends the last rows to a second,
    DB-segmented launch (plan_tail); results must equal the brute-force kernel bit for bit."""
    from lemon_b200 import plan_tail
    sc = lb.get_scorer()
    units = sc.num_sms // 2
    nq, m, d = units * 256 + 700, 20000, 64
    n_main, nseg = plan_tail(nq, m, sc.num_sms, 2)
    assert n_main == units * 256 and nseg >= 2
    x, _, _, _ = clustered_pairs(m, d, n_clusters=80, seed=123)
    q = np.concatenate([x, x[: nq - m] * 0.5 + x[7:7 + nq - m] * 0.5]) if nq > m else x[:nq]
    qp, dbp = sc.prepare(q, True), sc.prepare(x, True)
    tv, ti = sc.knn(qp, dbp, 31, 0, mode="tc")
    assert isinstance(sc.last_info["nseg"], list) and sc.last_info["nseg"][1] == nseg
    ev, ei = sc.knn(qp, dbp, 31, 0, mode="exact")
    assert bool((ti == ei).all()) and bool((tv == ev).all())


def test_sharded_driver_host_inputs_equal_device_inputs(lb):
    """Pinned host shards (copy stream, image side overlapped with the text copy) == device shards, bitwise."""
    import torch
    from lemon_b200 import dist as ldist
    x, y, _, _ = clustered_pairs(5000, 512, n_clusters=40, seed=77, noise_frac=0.3)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    a = ldist.score_pairs_sharded(xd, yd, 5000, k=12, hparams=HP)
    b = ldist.score_pairs_sharded(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(), 5000, k=12, hparams=HP)
    torch.cuda.synchronize()
    assert a["rows"] == b["rows"] == (0, 5000)
    for c in ("score", "I_n", "I_m", "D_n", "dists_m", "d_1"):
        assert bool((a[c] == b[c]).all()), c
    check_against_oracle(_np({k: v for k, v in b.items() if k != "rows"}), x, y, x, y, k=12, query_in_db=np.arange(5000), hparams=HP)


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_shapes_default_path_equals_exact_and_oracle(lb, seed):
    """Random shapes through knn(): tiny and mid-size DBs (exact kernel), TC-sized DBs, odd dims, k up to 63, both
    metrics.  The default path must equal the brute-force kernel bit for bit, and the oracle modulo eps-ties."""
    from oracle import lemon_oracle as O
    rng = np.random.RandomState(1000 + seed)
    m = int(rng.choice([40, 300, 1900, 2500, 7000, 17000]))
    nq = int(rng.choice([1, 33, 200, 600]))
    d = int(rng.choice([20, 100, 300, 512, 640, 768]))
    kp = int(rng.choice([1, 2, 6, 31, 51, 63]))
    metric = int(rng.randint(0, 2))
    x, _, _, _ = clustered_pairs(m, d, n_clusters=max(2, m // 50), seed=seed)
    q = (x[rng.randint(0, m, nq)] + 0.05 * rng.standard_normal((nq, d))).astype(np.float32)
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(q, True), sc.prepare(x, True)
    tv, ti = sc.knn(qp, dbp, kp, metric)                      # default ("auto") path
    ev, ei = sc.knn(qp, dbp, kp, metric, mode="exact")
    assert bool((ti == ei).all()) and bool((tv == ev).all()), (m, nq, d, kp, metric)
    qn, xn = qp.f32.cpu().numpy()[:, :d], dbp.f32.cpu().numpy()[:, :d]
    name = "ip" if metric == 0 else "l2"
    D, I = O.knn_search(qn, xn, kp, name)
    got = ti.cpu().numpy().astype(np.int64)
    if m >= kp:
        r = O.compare_neighbor_sets(qn, xn, got, kp, name, eps_tie=4e-6, D_ref=D, I_ref=I)
        assert r["wrong"] == 0, (m, nq, d, kp, metric, r["wrong_rows"][:3])
    else:
        assert (got[:, m:] == -1).all() and (np.sort(got[:, :m], 1) == np.arange(m)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("m,nseg,keep", [(40_000, 8, 40), (20_000, 8, 40), (40_000, 8, 64), (30_000, 16, 32)])
def test_many_segments_stay_certified(lb, m, nseg, keep):
    """A tail launch splits the DB into up to 8 segments = 16 candidate lists per row, each certifying only its own
    `keep` best.  The re-rank selects over the UNION of the lists, so the rows stay certified (no exact fallback)
    and equal the brute-force kernel bit for bit.  (A per-list bound left > 256 candidates per row here and sent
    every tail row of a 2-GPU C2 run to the exact kernel.)  m = 20000 runs without the threshold bootstrap, so
    every list fills up once and goes through the exact reduction."""
    import torch
    sc = lb.get_scorer()
    d, kp, nq = 128, 31, 900
    x, _, _, _ = clustered_pairs(m, d, n_clusters=m // 150, seed=5)
    qp, dbp = sc.prepare(x[:nq].copy(), True), sc.prepare(x, True)
    ck, cc, ct, ns = sc.knn_candidates(qp, dbp, nseg=nseg, cta_group=2, keep=keep)
    assert ns == nseg and ck.shape[1] == 2 * nseg and int(cc[:nq].max()) <= 1024
    tv, ti, uncert, n_unc = sc.rerank(qp, dbp, (ck, cc, ct), kp, 0)
    ev, ei = sc.knn_exact(qp, dbp, kp, 0)
    torch.cuda.synchronize()
    assert int(n_unc.item()) == 0
    assert bool((ti == ei).all()) and bool((tv == ev).all())
    # the two lists of a segment (one per epilogue group; they adopt each other's thresholds) together hold at
    # least `keep` columns at or above the larger of their two thresholds
    k64 = ck[:nq].cpu().numpy().view(np.uint64)
    hi = (k64 >> np.uint64(32)).astype(np.uint32)
    vals = hi.view(np.float32)                                                # K1 stores the raw fp32 bit pattern
    cnt, th = cc[:nq].cpu().numpy(), ct[:nq].cpu().numpy()
    valid = np.arange(1024)[None, None, :] < cnt[:, :, None]
    th_seg = th.reshape(nq, nseg, 2).max(-1)                                  # [nq, nseg]
    above = (valid & (vals >= np.repeat(th_seg, 2, axis=1)[:, :, None])).sum(-1).reshape(nq, nseg, 2).sum(-1)
    assert (above[np.isfinite(th_seg)] >= keep).all()
