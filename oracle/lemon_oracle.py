"""CPU oracle for the LEMoN pair-scoring hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in numpy (float64 arithmetic on fp32 inputs), what the
reference computes on the path named by BASELINE.json:north_star.  It is the
checker for the CUDA path; it is never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.

Reference lines restated (all under /root/reference, none copied):

* ``normalize_vectors``            lib/utils/utils.py:39-40
* DB subsample                     run_lemon.py:122-127
* ``dists_tr``                     run_lemon.py:169 (cosine), :173 (euclidean)
* ``index.search``                 run_lemon.py:235-236  (faiss IndexFlatIP/L2)
* per-sample loop                  run_lemon.py:238-307
* ``calc_scores_given_hparams``    lib/metrics/utils.py:21-45   (row loop)
* ``..._vectorized``               lib/metrics/utils.py:47-82

PARITY PINNING STATUS
---------------------
* score combination (utils.py:21-82): PINNED.  ``tests/golden/*.npz`` hold
  outputs of the reference's own ``calc_scores_given_hparams(_vectorized)``
  imported live from /root/reference (``tests/golden/make_golden.py``); the
  oracle is checked against them in ``tests/test_oracle.py``.
* ``normalize_vectors``: PINNED the same way (reference function run live).
* kNN search: **parity unpinned** at the faiss boundary.  faiss-gpu
  (requirements.txt:20, unpinned version) is a third-party dependency absent
  from /root/reference and from this image, and the reference has no tests or
  golden vectors.  The published algorithm of ``IndexFlatIP`` / ``IndexFlatL2``
  (exact brute-force top-k by inner product, descending / squared L2,
  ascending) is restated here in float64 with the documented total order
  (value best-first, then DB index ascending).  Acceptance is defined as set
  equality modulo eps-ties at the k-th boundary (see ``compare_neighbor_sets``).
"""
from __future__ import annotations

import numpy as np

EPS_TIE_COSINE = 2e-6   # fp32 accumulation noise at d=768
HP_KEYS = ("beta", "gamma", "tau_1_n", "tau_2_n", "tau_1_m", "tau_2_m")
CC3M_HPARAMS = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0,
                "tau_1_m": 0.1, "tau_2_m": 5.0}   # train_clip_from_scratch.py:102-109


# --------------------------------------------------------------------------- a1
def normalize_vectors(v: np.ndarray) -> np.ndarray:
    """Row-wise ``x / max(||x||_2, 1e-12)`` in fp32 (lib/utils/utils.py:39-40 ->
    ``torch.nn.functional.normalize(p=2, dim=1)``)."""
    v = np.asarray(v, dtype=np.float32)
    out = np.empty_like(v)
    for s in range(0, v.shape[0], 1 << 18):        # row blocks: no full-size float64 temporary
        b = v[s:s + (1 << 18)]
        nrm = np.sqrt((b.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
        out[s:s + (1 << 18)] = b / np.maximum(nrm, np.float32(1e-12))[:, None]
    return out


# --------------------------------------------------------------------------- a2
def subsample_db(n_train: int, limit: int, rng: np.random.RandomState | None = None) -> np.ndarray:
    """run_lemon.py:122-127: DB = all train rows, or ``limit`` rows drawn without
    replacement (unsorted) when the train split is larger."""
    if n_train > limit:
        rng = rng if rng is not None else np.random
        return rng.choice(np.arange(n_train), limit, replace=False)
    return np.arange(n_train)


def query_in_db_from_indices(n_queries: int, train_indices_in_compr: np.ndarray) -> np.ndarray:
    """For each train-split query ``sample_idx`` return the DB row holding it, or -1
    (run_lemon.py:258,278 only test membership; the row id is what the fused API takes)."""
    out = np.full(n_queries, -1, dtype=np.int64)
    idx = np.asarray(train_indices_in_compr, dtype=np.int64)
    ok = idx < n_queries
    out[idx[ok]] = np.nonzero(ok)[0]
    return out


# --------------------------------------------------------------------------- a3
def dists_tr(emb_txt_tr: np.ndarray, emb_img_tr: np.ndarray, dist_type: str) -> np.ndarray:
    """Per-DB-row cross-modal distance (run_lemon.py:169 / :173), float64."""
    t = emb_txt_tr.astype(np.float64)
    x = emb_img_tr.astype(np.float64)
    if dist_type == "cosine":
        return 1.0 - (t * x).sum(axis=1)
    if dist_type == "euclidean":
        return ((t - x) ** 2).sum(axis=1)
    raise ValueError(dist_type)


# --------------------------------------------------------------------------- a5
def _det_values(q64: np.ndarray, g64: np.ndarray, metric: str) -> np.ndarray:
    """Deterministic float64 values of query rows [n,d] against gathered DB rows [n,c,d]: numpy's row-wise pairwise
    summation depends only on the row CONTENT, so bit-identical DB rows get bit-identical values (a BLAS gemm does
    not guarantee that: its result can depend on where a column sits in the matrix)."""
    if metric == "ip":
        return (q64[:, None, :] * g64).sum(-1)
    return ((q64[:, None, :] - g64) ** 2).sum(-1)


def _topk_chunk(q64: np.ndarray, db64: np.ndarray, S: np.ndarray, k: int, metric: str, pad: int = 32):
    """Exact top-k of every row of the float64 similarity block S = f(q64, db64) under the total order (best value
    first, then column index ascending).  The gemm values only pre-select k+pad candidates; their values are then
    recomputed deterministically and ordered.  A row whose k-th value ties with the weakest candidate (mass ties
    reaching past the candidate set) is redone with a full deterministic evaluation."""
    nq, m = S.shape
    largest = metric == "ip"
    k_eff = min(k, m)
    ncand = min(m, k_eff + pad)
    key = -S if largest else S
    if ncand < m:
        cand = np.argpartition(key, ncand - 1, axis=1)[:, :ncand]
    else:
        cand = np.broadcast_to(np.arange(m), (nq, m)).copy()
    vals = np.empty((nq, ncand))
    for s in range(0, nq, 256):
        vals[s:s + 256] = _det_values(q64[s:s + 256], db64[cand[s:s + 256]], metric)
    o = np.lexsort((cand, -vals if largest else vals), axis=1)
    cand, vals = np.take_along_axis(cand, o, 1), np.take_along_axis(vals, o, 1)
    if ncand < m:
        for r in np.nonzero(np.abs(vals[:, k_eff - 1] - vals[:, -1]) <= 1e-12)[0]:
            full = _det_values(q64[r:r + 1], db64[None], metric)[0]
            oo = np.lexsort((np.arange(m), -full if largest else full))[:ncand]
            cand[r], vals[r] = oo, full[oo]
    return vals[:, :k_eff], cand[:, :k_eff].astype(np.int64)


def knn_search(q: np.ndarray, db: np.ndarray, k: int, metric: str = "ip",
               block: int = 2048, db_chunk: int = 131072) -> tuple[np.ndarray, np.ndarray]:
    """Exact brute-force kNN, float64 (stands in for faiss ``IndexFlatIP.search`` /
    ``IndexFlatL2.search`` at run_lemon.py:235-236).

    metric 'ip': D = <q,b>, descending.  metric 'l2': D = ||q-b||^2, ascending.
    Returns (D float64 [nq,k], I int64 [nq,k]); ``ntotal < k`` pads I=-1, D=-inf/+inf as faiss does.  The DB is
    walked in chunks of `db_chunk` rows (only a chunk is ever held in float64, so multi-million-row databases fit
    in host memory); every chunk yields its exact top-k under the total order (value best-first, then DB index
    ascending) and the chunk lists are merged under the same order, which is exact."""
    if metric not in ("ip", "l2"):
        raise ValueError(metric)
    q = np.asarray(q)
    db = np.asarray(db)
    nq, m = q.shape[0], db.shape[0]
    largest = metric == "ip"
    D = np.full((nq, k), -np.inf if largest else np.inf, dtype=np.float64)
    I = np.full((nq, k), -1, dtype=np.int64)
    block = max(1, min(block, int(2.5e8 // max(1, min(m, db_chunk)))))      # similarity block <= ~2 GB of float64
    for s in range(0, nq, block):
        e = min(nq, s + block)
        q64 = np.ascontiguousarray(q[s:e], dtype=np.float64)
        best_v = best_i = None
        for c0 in range(0, m, db_chunk):
            c1 = min(m, c0 + db_chunk)
            db64 = np.ascontiguousarray(db[c0:c1], dtype=np.float64)
            S = q64 @ db64.T
            if metric == "l2":      # pre-selection only: ||q||^2 + ||b||^2 - 2<q,b>, the expansion faiss uses
                S = (q64 ** 2).sum(axis=1)[:, None] + (db64 ** 2).sum(axis=1)[None, :] - 2.0 * S
            v, i = _topk_chunk(q64, db64, S, k, metric)
            i = i + c0
            if best_v is None:
                best_v, best_i = v, i
            else:                                  # merge two exact lists under (value best-first, index ascending)
                av = np.concatenate([best_v, v], axis=1)
                ai = np.concatenate([best_i, i], axis=1)
                o = np.lexsort((ai, -av if largest else av), axis=1)[:, :k]
                best_v, best_i = np.take_along_axis(av, o, 1), np.take_along_axis(ai, o, 1)
        if best_v is not None:
            D[s:e, :best_v.shape[1]], I[s:e, :best_v.shape[1]] = best_v, best_i
    return D, I


# --------------------------------------------------------------------------- a7
def apply_self_exclusion(D: np.ndarray, I: np.ndarray, in_db: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Train-split rule of run_lemon.py:257-263 / :277-283: the search returned k+1
    neighbours; drop rank 0 when the sample is in the DB, otherwise drop the last.
    (The reference never checks that rank 0 *is* the sample.)"""
    in_db = np.asarray(in_db, dtype=bool)
    Dk = np.where(in_db[:, None], D[:, 1:], D[:, :-1])
    Ik = np.where(in_db[:, None], I[:, 1:], I[:, :-1])
    return Dk, Ik


# ----------------------------------------------------------------------- a6-a10
def build_records(img_q, txt_q, img_db, txt_db, D_n, I_n, D_m, I_m, dist_type: str,
                  text_label_ids_q=None, text_label_ids_db=None, class_text_emb=None, noisy_label=None) -> dict:
    """Vectorised restatement of the per-sample loop body run_lemon.py:250-307 for
    given (already self-excluded) neighbour lists.  float64 arithmetic.

    D_n/D_m come in as the search returned them (inner product, or squared L2).
    Returns the df columns: d_1 [N], D_n, dists_n, dists_tr_n, D_m, dists_m,
    dists_tr_m (each [N,k])."""
    xq = np.asarray(img_q, np.float64)
    yq = np.asarray(txt_q, np.float64)
    xdb = np.asarray(img_db)          # DB rows are gathered first and widened to float64 afterwards
    ydb = np.asarray(txt_db)
    dtr_rows = lambda I: dists_tr(ydb[I.reshape(-1)], xdb[I.reshape(-1)], dist_type).reshape(I.shape)
    discrete = text_label_ids_q is not None
    cos = dist_type == "cosine"
    N, k = I_n.shape
    d_1 = (1.0 - (xq * yq).sum(1)) if cos else ((xq - yq) ** 2).sum(1)       # :250-253
    if class_text_emb is not None:                                            # --normalize_d1, :244-248
        T = np.asarray(class_text_emb, np.float64)
        dc = (1.0 - xq @ T.T) if cos else ((xq[:, None, :] - T[None]) ** 2).sum(-1)
        e = np.exp(dc - dc.max(axis=1, keepdims=True))                        # scipy.special.softmax
        d_1 = (e / e.sum(axis=1, keepdims=True))[np.arange(N), np.asarray(noisy_label)]
    dists_n = np.empty((N, k))
    dists_m = np.empty((N, k))
    bs = 1024
    for s in range(0, N, bs):
        e = min(N, s + bs)
        if discrete:                                                          # :266-267
            dists_n[s:e] = 1.0 - (np.asarray(text_label_ids_db)[I_n[s:e]]
                                  == np.asarray(text_label_ids_q)[s:e, None])
        else:
            y_n = ydb[I_n[s:e]].astype(np.float64)                            # :264
            dists_n[s:e] = (1.0 - np.einsum("nd,nkd->nk", yq[s:e], y_n)) if cos else \
                ((yq[s:e, None, :] - y_n) ** 2).sum(-1)                       # :270-273
        x_m = xdb[I_m[s:e]].astype(np.float64)                                # :284
        dists_m[s:e] = (1.0 - np.einsum("nd,nkd->nk", xq[s:e], x_m)) if cos else \
            ((xq[s:e, None, :] - x_m) ** 2).sum(-1)                           # :286-289
    Dn = np.asarray(D_n, np.float64).copy()
    Dm = np.asarray(D_m, np.float64).copy()
    if cos:
        if not discrete:
            Dn = -Dn          # :270 (sits in the else-branch: skipped for the discrete metric)
        Dm = -Dm              # :286
    return {"d_1": d_1, "D_n": Dn, "dists_n": dists_n, "dists_tr_n": dtr_rows(I_n),
            "D_m": Dm, "dists_m": dists_m, "dists_tr_m": dtr_rows(I_m)}


# ------------------------------------------------------------------------- a11
def calc_scores_vectorized(rec: dict, hp: dict) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """lib/metrics/utils.py:63-77 in float64.  Returns (scores, d_ns, d_ms)."""
    f = lambda a: np.asarray(a, np.float64)
    w_n = np.exp(-hp["tau_1_n"] * f(rec["D_n"])) * np.exp(-hp["tau_2_n"] * f(rec["dists_tr_n"]))
    w_m = np.exp(-hp["tau_1_m"] * f(rec["D_m"])) * np.exp(-hp["tau_2_m"] * f(rec["dists_tr_m"]))
    d_ns = (w_n * f(rec["dists_n"])).sum(axis=1) / f(rec["D_n"]).shape[1]
    d_ms = (w_m * f(rec["dists_m"])).sum(axis=1) / f(rec["D_m"]).shape[1]
    scores = f(rec["d_1"]) + hp["beta"] * d_ns + hp["gamma"] * d_ms
    return scores, d_ns, d_ms


def calc_scores_loop(rec: dict, hp: dict) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Row-by-row twin (lib/metrics/utils.py:21-45); pure-Python loop, small cases only."""
    N = len(rec["d_1"])
    d_ns = np.empty(N)
    d_ms = np.empty(N)
    for i in range(N):
        sf = np.ones(len(rec["D_n"][i]))
        sf = sf * np.exp(-hp["tau_1_n"] * np.asarray(rec["D_n"][i], np.float64))
        sf = sf * np.exp(-hp["tau_2_n"] * np.asarray(rec["dists_tr_n"][i], np.float64))
        d_ns[i] = np.dot(sf, np.asarray(rec["dists_n"][i], np.float64)) / len(rec["D_n"][i])
        sf = np.ones(len(rec["D_m"][i]))
        sf = sf * np.exp(-hp["tau_1_m"] * np.asarray(rec["D_m"][i], np.float64))
        sf = sf * np.exp(-hp["tau_2_m"] * np.asarray(rec["dists_tr_m"][i], np.float64))
        d_ms[i] = np.dot(sf, np.asarray(rec["dists_m"][i], np.float64)) / len(rec["D_m"][i])
    scores = np.asarray(rec["d_1"], np.float64) + hp["beta"] * d_ns + hp["gamma"] * d_ms
    return scores, d_ns, d_ms


# ----------------------------------------------------------------- whole path
def lemon_oracle(img_q, txt_q, img_db, txt_db, *, k: int, dist_type: str = "cosine",
                 query_in_db=None, hparams: dict | None = None, normalize: bool = True,
                 text_label_ids_q=None, text_label_ids_db=None, given_I=None, class_text_emb=None,
                 noisy_label=None) -> dict:
    """Truth oracle for the whole path (run_lemon.py:163-176, 235-307 + utils.py:47-82).

    query_in_db: None for val/test-style queries (search k, keep all), or int64[N]
    holding the DB row of each train query (-1 = not in DB): search k+1 and apply the
    reference's drop rule.  given_I=(I_n, I_m): skip the search and evaluate the
    records on those neighbour lists (used to check tie-excused rows)."""
    if normalize:
        img_q, txt_q = normalize_vectors(img_q), normalize_vectors(txt_q)
        img_db, txt_db = normalize_vectors(img_db), normalize_vectors(txt_db)
    metric = "ip" if dist_type == "cosine" else "l2"
    train = query_in_db is not None
    kk = k + 1 if train else k
    out = {}
    if given_I is None:
        D_n, I_n = knn_search(img_q, img_db, kk, metric)
        D_m, I_m = knn_search(txt_q, txt_db, kk, metric)
        out["raw"] = (D_n, I_n, D_m, I_m)
        if train:
            in_db = np.asarray(query_in_db) >= 0
            D_n, I_n = apply_self_exclusion(D_n, I_n, in_db)
            D_m, I_m = apply_self_exclusion(D_m, I_m, in_db)
    else:
        I_n, I_m = (np.asarray(a, np.int64) for a in given_I)
        D_n = pair_values(img_q, img_db, I_n, metric)
        D_m = pair_values(txt_q, txt_db, I_m, metric)
    if class_text_emb is not None and normalize:
        class_text_emb = normalize_vectors(class_text_emb)
    rec = build_records(img_q, txt_q, img_db, txt_db, D_n, I_n, D_m, I_m, dist_type,
                        text_label_ids_q, text_label_ids_db, class_text_emb, noisy_label)
    rec["I_n"], rec["I_m"] = I_n, I_m
    if hparams is not None:
        rec["score"], rec["s_n"], rec["s_m"] = calc_scores_vectorized(rec, hparams)
    out.update(rec)
    return out


def pair_values(q, db, I, metric: str) -> np.ndarray:
    """float64 similarity / squared distance of each query to the listed DB rows."""
    q64 = np.asarray(q, np.float64)
    db = np.asarray(db)
    out = np.empty(I.shape, np.float64)
    bs = 1024
    for s in range(0, I.shape[0], bs):
        e = min(I.shape[0], s + bs)
        g = db[I[s:e]].astype(np.float64)
        if metric == "ip":
            out[s:e] = np.einsum("nd,nkd->nk", q64[s:e], g)
        else:
            out[s:e] = ((q64[s:e, None, :] - g) ** 2).sum(-1)
    return out


# ------------------------------------------------------------- acceptance rule
def compare_neighbor_sets(q, db, I_got: np.ndarray, k: int, metric: str = "ip",
                          eps_tie: float = EPS_TIE_COSINE, D_ref=None, I_ref=None, top_boundary=None) -> dict:
    """SURVEY.md §8c acceptance: per row the returned index SET must equal the
    float64 oracle's, except members whose float64 value lies within ``eps_tie`` of
    the oracle's k-th value (eps-ties at the boundary).  top_boundary [N] (NaN = none):
    value of the rank-0 entry the train-split rule dropped (run_lemon.py:257-263 drops
    rank 0 without checking that it IS the sample, so with duplicate rows which of the
    tied entries goes is a tie at that second boundary as well).  Returns counts and the
    rows that were tie-excused / wrong."""
    if I_ref is None:
        D_ref, I_ref = knn_search(q, db, k, metric)
    got_vals = pair_values(q, db, np.where(I_got < 0, 0, I_got), metric)
    kth = D_ref[:, k - 1]
    exact = excused = 0
    wrong_rows, excused_rows = [], []
    for r in range(I_got.shape[0]):
        sg, sr = set(I_got[r].tolist()), set(I_ref[r].tolist())
        if len(sg) != I_got.shape[1]:
            wrong_rows.append(r)          # duplicates / padding where none expected
            continue
        if sg == sr:
            exact += 1
            continue
        bounds = [kth[r]]
        if top_boundary is not None and np.isfinite(top_boundary[r]):
            bounds.append(top_boundary[r])
        near = lambda v: any(abs(v - b) <= eps_tie for b in bounds)
        extra = [j for j, i in enumerate(I_got[r]) if i not in sr]
        ok = all(near(got_vals[r, j]) for j in extra)
        miss = [j for j, i in enumerate(I_ref[r]) if i not in sg]
        ok = ok and all(near(D_ref[r, j]) for j in miss)
        if ok:
            excused += 1
            excused_rows.append(r)
        else:
            wrong_rows.append(r)
    return {"rows": int(I_got.shape[0]), "exact": exact, "tie_excused": excused,
            "wrong": len(wrong_rows), "wrong_rows": wrong_rows, "excused_rows": excused_rows}


# ------------------------------------------------------- timed CPU baseline port
def reference_cpu_scorer(img_q, txt_q, img_db, txt_db, *, k: int, dist_type: str = "cosine",
                         train_indices_in_compr=None, sample_offset: int = 0,
                         hparams: dict | None = None, batch_size: int = 128,
                         score_fn=None, faiss_module=None, timings: dict | None = None):
    """Operation-for-operation port of the reference CPU scorer, used ONLY as the timed
    CPU baseline (bench.py cpu_baseline / --impl reference).

    Mirrors: fp32 ``F.normalize`` (utils.py:39-40) -> index build + dists_tr
    (run_lemon.py:166-176) -> per-128-batch search of both indices (:235-236; faiss
    CPU IndexFlat is BLAS sgemm + heap, stood in for by torch ``matmul`` + ``topk`` —
    same BLAS class, same complexity) -> the per-sample Python loop with its tiny
    torch ops and dict/array allocations (:238-307) -> ``pd.DataFrame`` (:314) ->
    ``calc_scores_given_hparams_vectorized`` (utils.py:47-82; ``score_fn`` may be the
    live reference function).  train_indices_in_compr=None means a val/test split.

    faiss_module: a module with the faiss API (``IndexFlatIP/IndexFlatL2.add/search``); when given, the index build
    and the per-batch searches go through it exactly as run_lemon.py:166-176,235-236 call faiss (the seam test
    passes ``lemon_b200.faiss_compat``).  timings: dict that receives 'prep_s' (normalise + index build: paid once
    per database) and 'query_s' (everything per query batch + DataFrame + scoring)."""
    import time
    import torch
    import pandas as pd

    t_start = time.perf_counter()

    F = torch.nn.functional
    emb_txt_tr = F.normalize(torch.as_tensor(txt_db, dtype=torch.float32), p=2, dim=1)
    emb_img_tr = F.normalize(torch.as_tensor(img_db, dtype=torch.float32), p=2, dim=1)
    cos = dist_type == "cosine"
    if cos:
        tr_d = 1 - (emb_txt_tr * emb_img_tr).sum(axis=1)
    else:
        tr_d = ((emb_txt_tr - emb_img_tr) ** 2).sum(axis=1)
    db_img_t = emb_img_tr.t().contiguous()
    db_txt_t = emb_txt_tr.t().contiguous()
    if not cos:
        n_img = (emb_img_tr ** 2).sum(1)
        n_txt = (emb_txt_tr ** 2).sum(1)
    train = train_indices_in_compr is not None
    kk = k + int(train)
    if faiss_module is not None:
        d = emb_img_tr.shape[1]
        index_img = faiss_module.IndexFlatIP(d) if cos else faiss_module.IndexFlatL2(d)      # run_lemon.py:167-168 / 171-172
        index_txt = faiss_module.IndexFlatIP(d) if cos else faiss_module.IndexFlatL2(d)
        index_img.add(emb_img_tr.numpy())                                                    # :175-176
        index_txt.add(emb_txt_tr.numpy())
    t_prep = time.perf_counter()

    def search(qb, db_t, nrm):
        if faiss_module is not None:
            index = index_img if db_t is db_img_t else index_txt
            Dq, Iq = index.search(qb.numpy(), kk)                                            # :235-236
            return torch.from_numpy(Dq), torch.from_numpy(Iq)
        ip = qb @ db_t
        if cos:
            return torch.topk(ip, kk, dim=1)
        d2 = (qb ** 2).sum(1, keepdim=True) + nrm[None, :] - 2 * ip
        v, i = torch.topk(d2, kk, dim=1, largest=False)
        return v, i

    rows = []
    all_img = torch.as_tensor(img_q, dtype=torch.float32)
    all_txt = torch.as_tensor(txt_q, dtype=torch.float32)
    for b0 in range(0, all_img.shape[0], batch_size):
        qi = F.normalize(all_img[b0:b0 + batch_size], p=2, dim=1)
        qt = F.normalize(all_txt[b0:b0 + batch_size], p=2, dim=1)
        Dn_b, In_b = search(qi, db_img_t, None if cos else n_img)
        Dm_b, Im_b = search(qt, db_txt_t, None if cos else n_txt)
        Dn_b, In_b, Dm_b, Im_b = Dn_b.numpy(), In_b.numpy(), Dm_b.numpy(), Im_b.numpy()
        for j in range(qi.shape[0]):
            sidx = sample_offset + b0 + j
            xi, yi = qi[j, None], qt[j, None]
            if cos:
                d1 = 1 - torch.dot(xi.flatten(), yi.flatten())
            else:
                d1 = ((xi.flatten() - yi.flatten()) ** 2).sum()
            dn, nn_i = Dn_b[j], In_b[j]
            if train:
                if sidx in train_indices_in_compr:      # O(M) scan, as in the reference
                    nn_i, dn = nn_i[1:], dn[1:]
                else:
                    nn_i, dn = nn_i[:-1], dn[:-1]
            nb_txt = emb_txt_tr[nn_i]
            if cos:
                dn = -dn
                dist_n = 1 - (yi * nb_txt).sum(axis=1)
            else:
                dist_n = ((yi - nb_txt) ** 2).sum(axis=1)
            dm, mm_i = Dm_b[j], Im_b[j]
            if train:
                if sidx in train_indices_in_compr:
                    mm_i, dm = mm_i[1:], dm[1:]
                else:
                    mm_i, dm = mm_i[:-1], dm[:-1]
            nb_img = emb_img_tr[mm_i]
            if cos:
                dm = -dm
                dist_m = 1 - (xi * nb_img).sum(axis=1)
            else:
                dist_m = ((xi - nb_img) ** 2).sum(axis=1)
            rows.append({"sset": "train" if train else "test", "idx": sidx,
                         "d_1": d1.item(), "dists_n": dist_n.numpy(), "D_n": dn.flatten(),
                         "dists_tr_n": tr_d[nn_i].numpy(), "dists_m": dist_m.numpy(),
                         "D_m": dm.flatten(), "dists_tr_m": tr_d[mm_i].numpy(),
                         "I_n": nn_i, "I_m": mm_i})
    df = pd.DataFrame(rows)
    if hparams is not None:
        if score_fn is None:
            rec = {c: np.stack(df[c].values) for c in
                   ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")}
            rec["d_1"] = df["d_1"].values
            df["score"] = calc_scores_vectorized(rec, hparams)[0]
        else:
            df["score"] = score_fn(df, hparams)
    if timings is not None:
        timings["prep_s"] = t_prep - t_start
        timings["query_s"] = time.perf_counter() - t_prep
    return df
