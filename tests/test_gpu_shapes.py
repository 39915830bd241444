"""GPU parity at the BENCHED shapes and on hard distributions, through the default path
(tensor-core candidates -> fp32 re-rank -> certificate -> exact fallback), against the float64 oracle on sampled
rows; plus the kNN answers of independent implementations (tests/golden/knn.npz) and the device-side helpers that
replaced torch plumbing in round 2 (duplicate grouping, keep_lowest)."""
import os

import numpy as np
import pytest

from tests.helpers import KNN_PIN_CASES, check_against_oracle, knn_pin_case

pytestmark = pytest.mark.gpu
HP = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "knn.npz"))


@pytest.fixture(scope="module")
def lb():
    import torch
    import lemon_b200
    assert torch.cuda.is_available()
    return lemon_b200


def _sampled_check(out, x, y, rows, k, qid_rows=None, xdb=None, ydb=None, **kw):
    sub = {c: t[rows].cpu().numpy() for c, t in out.items()}
    xh, yh = x.cpu().numpy(), y.cpu().numpy()
    xd = xh if xdb is None else xdb.cpu().numpy()
    yd = yh if ydb is None else ydb.cpu().numpy()
    qid = rows if qid_rows is None else qid_rows
    return check_against_oracle(sub, xh[rows], yh[rows], xd, yd, k=k, query_in_db=qid, hparams=HP, **kw)


def _run_full(lb, name, n_rows, seed=3):
    import torch
    from bench import WORKLOADS, workload_pairs
    from lemon_b200.scoring import count_uncertified
    wl = WORKLOADS[name]
    dev = torch.device("cuda", 0)
    x, y, _ = workload_pairs(wl, dev)
    out = lb.score_pairs(x, y, k=wl["k"], query_in_db=torch.arange(wl["n"], device=dev), hparams=HP, device=0)
    info = lb.get_scorer(0).last_info
    rows = np.sort(np.random.RandomState(seed).choice(wl["n"], n_rows, replace=False))
    st = _sampled_check(out, x, y, rows, wl["k"])
    unc = {s: count_uncertified(info[s]) for s in ("img", "txt")}
    return st, unc, info, out


def test_full_c2_shape_sampled_rows_vs_oracle(lb):
    """C2 as benched: 118 000 x 118 000 x 512, cat noise 0.4 (de-duplicated text side, tail launch); 2048 sampled
    rows, the LAST rows (served by the tail launch) included."""
    import torch
    from bench import WORKLOADS, workload_pairs
    wl = WORKLOADS["c2"]
    dev = torch.device("cuda", 0)
    x, y, _ = workload_pairs(wl, dev)
    n = wl["n"]
    out = lb.score_pairs(x, y, k=30, query_in_db=torch.arange(n, device=dev), hparams=HP, device=0)
    info = lb.get_scorer(0).last_info
    assert info["img"]["path"] == "tc" and info["txt"]["path"] == "tc"
    assert info["txt"].get("n_unique", n) < 0.95 * n            # caption noise -> exact duplicate rows are searched once
    rows = np.unique(np.concatenate([np.random.RandomState(1).choice(n, 1792, replace=False), np.arange(n - 256, n)]))
    st = _sampled_check(out, x, y, rows, 30)
    assert st["wrong"] == 0 and st["exact_n"] + st["tie_excused_n"] == len(rows)
    print("C2 parity:", {k: v for k, v in st.items() if k != "wrong_rows"})


def test_c3_shape_sampled_rows_vs_oracle(lb):
    st, unc, info, _ = _run_full(lb, "c3", 768)
    assert st["wrong"] == 0
    assert unc["img"] < 370 and unc["txt"] < 370                # < 0.1 % of the rows need the exact kernel
    print("C3 parity:", {k: v for k, v in st.items() if k != "wrong_rows"}, "uncertified", unc)


def test_iid_gaussian_stress_vs_oracle(lb):
    """SURVEY.md 8d worst case: iid unit vectors (neighbour similarities only ~4 sigma above the bulk)."""
    st, unc, info, _ = _run_full(lb, "c2iid", 1024)
    assert st["wrong"] == 0
    assert unc["img"] < 1180 and unc["txt"] < 1180              # < 1 % uncertified
    print("iid parity:", {k: v for k, v in st.items() if k != "wrong_rows"}, "uncertified", unc)


def test_d768_million_row_db_sampled_rows_vs_oracle(lb):
    """d = 768 with M >= 1 M rows: the streamed-query-chunk path of K1 (kres < kchunks), multi-round main launch and
    a tail launch (40 000 queries = 156.25 row tiles on 74 CTA pairs)."""
    import torch
    from bench import synth_pairs
    from lemon_b200.scoring import count_uncertified
    dev = torch.device("cuda", 0)
    m, nq, d, k = 1_050_000, 40_000, 768, 30
    x, y, _ = synth_pairs(m, d, 0.0, 77, dev)
    qid = torch.arange(nq, device=dev) * 26 + 3                 # queries are DB rows 3, 29, 55, ...
    out = lb.score_pairs(x[qid], y[qid], x, y, k=k, query_in_db=qid, hparams=HP, device=0)
    info = lb.get_scorer(0).last_info
    assert info["img"]["path"] == "tc"
    pick = np.sort(np.random.RandomState(2).choice(nq, 192, replace=False))
    pick = np.unique(np.concatenate([pick, np.arange(nq - 32, nq)]))
    sub = {c: t[pick].cpu().numpy() for c, t in out.items()}
    xh, yh = x.cpu().numpy(), y.cpu().numpy()
    g = qid.cpu().numpy()[pick]
    st = check_against_oracle(sub, xh[g], yh[g], xh, yh, k=k, query_in_db=g, hparams=HP)
    assert st["wrong"] == 0
    print("d768 1M parity:", {k_: v for k_, v in st.items() if k_ != "wrong_rows"},
          "uncertified", {s: count_uncertified(info[s]) for s in ("img", "txt")})


@pytest.mark.parametrize("tag", KNN_PIN_CASES)
@pytest.mark.parametrize("metric", ["ip", "l2"])
@pytest.mark.parametrize("mode", ["exact", "tc"])
def test_knn_matches_independent_implementations(lb, tag, metric, mode):
    """The CUDA kNN (both the fp32 brute-force kernel and the tensor-core path) against answers of scikit-learn /
    torch float64 / a pure-Python sort (tests/golden/knn.npz), i.e. implementations that share no code with the
    oracle: same index sets modulo eps-ties at the k-th boundary, same values."""
    from oracle import lemon_oracle as O
    if f"{tag}_{metric}_I" not in G:
        pytest.skip("case has no answer for this metric")
    db, q, k, kind = knn_pin_case(tag)
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(q, normalize=False), sc.prepare(db, normalize=False)
    tv, ti = sc.knn(qp, dbp, k, 0 if metric == "ip" else 1, mode=mode)
    tv, ti = tv.cpu().numpy(), ti.cpu().numpy().astype(np.int64)
    I_ref, D_ref = G[f"{tag}_{metric}_I"], G[f"{tag}_{metric}_D"]
    scale = float(np.abs(D_ref).max()) if kind == "raw" else 1.0
    if kind == "ties":
        assert (ti == I_ref).all()                               # exact duplicates: ascending DB index
    r = O.compare_neighbor_sets(q, db, ti, k, metric, eps_tie=4e-6 * max(1.0, scale), D_ref=D_ref, I_ref=I_ref)
    assert r["wrong"] == 0
    same = (ti == I_ref).all(axis=1)
    assert same.mean() > 0.9
    np.testing.assert_allclose(tv[same], D_ref[same], rtol=2e-5, atol=4e-6 * max(1.0, scale))


@pytest.mark.parametrize("n,d,groups", [(70_001, 64, 900), (5000, 512, 10), (3000, 48, 3000), (130_000, 32, 40_000)])
def test_dedup_build_matches_numpy_grouping(lb, n, d, groups):
    """lemon_dedup_build (hash -> radix sort -> scans -> verification -> renumbering, no host round trip) against
    numpy's unique-rows grouping: representatives = lowest row of each group in ascending order, members ascending."""
    import torch
    rng = np.random.RandomState(n)
    base = rng.standard_normal((groups, d)).astype(np.float32)
    assign = rng.randint(0, groups, n)
    assign[:groups] = rng.permutation(groups)                    # every group occurs (when groups <= n)
    x = base[assign]
    sc = lb.get_scorer()
    p = sc.prepare(x, normalize=False)
    dd = sc.dedup_finish(p, sc.dedup_start(p), min_saving=0.0)
    _, first, inverse = np.unique(assign, return_index=True, return_inverse=True)
    order = np.argsort(first)                                    # groups by ascending first (= lowest) row
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    gid = rank[inverse]
    n_u = len(first)
    if n_u == n:
        assert dd is None or dd.n_unique == n
        return
    assert dd is not None and dd.n_unique == n_u
    off = dd.offsets.cpu().numpy()
    mem = dd.members.cpu().numpy()
    assert off[0] == 0 and off[-1] == n and len(off) == n_u + 1
    ref_off = np.concatenate([[0], np.cumsum(np.bincount(gid, minlength=n_u))])
    assert (off == ref_off).all()
    ref_mem = np.argsort(gid, kind="stable")
    assert (mem == ref_mem).all()
    assert (dd.uniq.f32.cpu().numpy() == x[np.sort(first)]).all()
    assert (dd.uniq.f16.cpu().numpy() == p.f16.cpu().numpy()[np.sort(first)]).all()


def test_expand_groups_merges_exactly_tied_distinct_rows(lb):
    """two DISTINCT unique rows that tie exactly in value must interleave their members by ascending DB
    index.  Rows +-e_1 mirrored in the unused coordinates give distinct rows with bit-equal similarities."""
    import torch
    d, reps = 64, 40
    a = np.zeros(d, np.float32); a[0] = 0.6; a[1] = 0.8
    b = np.zeros(d, np.float32); b[0] = 0.6; b[1] = -0.8        # <q, a> == <q, b> for q = e_0
    filler = np.random.RandomState(0).standard_normal((3000, d)).astype(np.float32) * 0.01
    db = filler.copy()
    pos = np.random.RandomState(1).choice(3000, 2 * reps, replace=False)
    db[pos[:reps]] = a
    db[pos[reps:]] = b
    q = np.zeros((4, d), np.float32); q[:, 0] = 1.0
    sc = lb.get_scorer()
    qp, dbp = sc.prepare(q, False), sc.prepare_db(db, False)
    assert dbp.dedup is None or dbp.dedup.n_unique < 3000
    dbp.dedup = sc.dedup_finish(dbp, sc.dedup_start(dbp), min_saving=0.0)
    tv, ti = sc.knn(qp, dbp, 31, 0, mode="exact")
    assert (ti.cpu().numpy() == np.sort(pos)[:31]).all()


@pytest.mark.parametrize("n,keep", [(100_000, 30_000), (5000, 5000), (70_000, 1), (3_300_000, 1_000_000)])
def test_keep_lowest_matches_sorted_order(lb, n, keep):
    """CC3M consumer (train_clip_from_scratch.py:110-113): ids of the `keep` lowest scores, ascending, ties by id."""
    import torch
    rng = np.random.RandomState(n)
    s = rng.standard_normal(n)
    s[rng.choice(n, n // 10, replace=False)] = np.round(s[:n // 10], 1)      # ties
    s[:7] = [-0.0, 0.0, -1e300, 1e300, 5e-324, -5e-324, 0.25]
    idx, val = lb.get_scorer().keep_lowest(torch.from_numpy(s), keep)
    ref = np.lexsort((np.arange(n), s))[:keep]
    got = idx.cpu().numpy()
    # -0.0 and 0.0 compare equal for numpy but not for the bit-pattern order: compare values, and ids off exact ties
    assert (s[got] == s[ref]).all()
    assert (val.cpu().numpy() == s[got]).all()
    differ = got != ref
    assert (s[got][differ] == 0.0).all() if differ.any() else True
    ids = lb.filter_lowest_scores(torch.from_numpy(s), min(keep, 100), idx=np.arange(n) * 2 + 1)
    assert (ids.cpu().numpy() == got[:100] * 2 + 1).all()


def _narrow_cone(n, d, shared, seed, dev):
    """CLIP-like 'cone' geometry: every embedding shares a large common component, which compresses all similarities
    (and the gaps between neighbours) by 1 - shared: the distribution on which an fp16 first pass loses certificates."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    c0 = torch.nn.functional.normalize(torch.randn(1, d, generator=g, device=dev), dim=1)
    z = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1)
    return torch.nn.functional.normalize(shared ** 0.5 * c0 + (1 - shared) ** 0.5 * z, dim=1).contiguous()


def test_split_precision_bound_is_rigorous(lb):
    """|q_hi.b_hi + q_hi.b_lo + q_lo.b_hi (fp32 TMEM accumulation) - float64 inner product| <= the bound the second
    pass certifies with, and that bound is ~an order of magnitude below the one-word fp16 bound."""
    import torch
    from lemon_b200.scoring import acc_eps_coef, acc_eps_coef_split, decode_candidates
    dev = torch.device("cuda", 0)
    x = _narrow_cone(20_000, 768, 0.5, 3, dev)
    sc = lb.get_scorer(0)
    dbp = sc.prepare(x, True)
    qp = lemon_slice(dbp, 0, 1024)
    qs, dbs = sc.split_operands(qp, 0), sc.split_operands(dbp, 1)
    assert qs.f16.shape == (1024, 3 * 768) and dbs.d16 == 2304
    ck, cc, ct, _ = sc.knn_candidates(qs, dbs, nseg=1, keep=64)
    cv, ci = decode_candidates(ck, cc, 1024)
    cv, ci = cv[:, :64].astype(np.float64), ci[:, :64]
    assert (ci >= 0).all()
    q64, db64 = qp.f32.cpu().numpy().astype(np.float64), dbp.f32.cpu().numpy().astype(np.float64)
    exact = np.einsum("nd,nkd->nk", q64, db64[ci])
    rs, smax = qs.row_stats.cpu().numpy(), dbs.stats_max.cpu().numpy()
    acc = acc_eps_coef_split(768, 768)
    eps2 = rs[:, 2] * smax[1] + rs[:, 0] * smax[2] + acc * np.maximum(rs[:, 0], rs[:, 1]) * max(smax[0], smax[1])
    err = np.abs(cv - exact).max(axis=1)
    assert (err <= eps2).all(), (err.max(), eps2.min())
    rs1, smax1 = qp.row_stats.cpu().numpy(), dbp.stats_max.cpu().numpy()
    eps1 = rs1[:, 2] * smax1[1] + rs1[:, 0] * smax1[2] + acc_eps_coef(768, 768) * rs1[:, 0] * smax1[0]
    assert np.median(eps2) < 0.2 * np.median(eps1)
    print("eps first pass %.2e, second pass %.2e, worst observed error %.2e" % (np.median(eps1), np.median(eps2), err.max()))


def lemon_slice(p, a, b):
    from lemon_b200.scoring import _slice_prepared
    return _slice_prepared(p, a, b)


def test_second_pass_rescues_rows_the_fp16_pass_cannot_certify(lb):
    """Narrow-cone embeddings: a large share of the rows fails the first-pass certificate; the split-precision second
    pass certifies (nearly) all of them, so (nearly) nothing reaches the fp32 brute-force kernel, and the lists are
    BITWISE those of the exact kernel."""
    import torch
    from lemon_b200.scoring import count_uncertified
    dev = torch.device("cuda", 0)
    n, d, kp = 100_000, 512, 31
    x = _narrow_cone(n, d, 0.98, 11, dev)
    sc = lb.get_scorer(0)
    dbp = sc.prepare(x, True)
    qp = lemon_slice(dbp, 0, 8192)
    tv, ti = sc.knn(qp, dbp, kp, 0, mode="tc")
    info = dict(sc.last_info)
    first, left = info["n_uncertified_first_pass"], count_uncertified(info)
    print("first pass uncertified %d of 8192, after the second pass %d" % (first, left))
    assert first > 400                      # the distribution is hard for one fp16 word ...
    assert left <= first // 20              # ... and easy for the split-precision pass
    ev, ei = sc.knn(lemon_slice(dbp, 0, 2048), dbp, kp, 0, mode="exact")
    assert (ti[:2048] == ei).all() and (tv[:2048] == ev).all()
    # with the second pass switched off the same rows go to the exact kernel: same lists
    sc.second_pass_enabled = False
    try:
        tv0, ti0 = sc.knn(lemon_slice(dbp, 0, 2048), dbp, kp, 0, mode="tc")
    finally:
        sc.second_pass_enabled = True
    assert (ti0 == ei).all() and (tv0 == ev).all()


@pytest.mark.parametrize("metric,kp", [(1, 31), (0, 51), (1, 51)])
def test_second_pass_other_metrics_and_list_lengths(lb, metric, kp):
    """The split-precision pass under the squared-L2 metric and with the longest lists LEMoN uses (k = 50, train split):
    hard rows get certified (or fall through to the exact kernel) and the lists stay bitwise those of the exact kernel."""
    import torch
    from lemon_b200.scoring import count_uncertified
    dev = torch.device("cuda", 0)
    x = _narrow_cone(60_000, 512, 0.98, 13, dev)
    sc = lb.get_scorer(0)
    dbp = sc.prepare(x, True)
    qp = lemon_slice(dbp, 1000, 3048)
    tv, ti = sc.knn(qp, dbp, kp, metric, mode="tc")
    info = dict(sc.last_info)
    ev, ei = sc.knn(qp, dbp, kp, metric, mode="exact")
    assert (ti == ei).all() and (tv == ev).all()
    print("metric %d kp %d: first pass uncertified %d of 2048, to the exact kernel %d" %
          (metric, kp, info["n_uncertified_first_pass"], count_uncertified(info)))
    if metric == 0:
        assert info["n_uncertified_first_pass"] > 100 and count_uncertified(info) <= info["n_uncertified_first_pass"] // 10
