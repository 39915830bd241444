"""End-to-end seam tests (SURVEY.md §4 "integration" row): the port of the reference's own
scoring loop (oracle.reference_cpu_scorer restates run_lemon.py:163-176,235-307,314,406) is run with
``lemon_b200.faiss_compat`` installed as ``faiss`` and the patched ``calc_scores_given_hparams_vectorized``, 128
queries per search call as run_lemon.py:45,202,235-236 do, against a SHUFFLED 50 000-of-N database
(run_lemon.py:48,122-127), and must agree with the fused ``score_pairs`` call on the same inputs."""
import sys
import types

import numpy as np
import pytest

from tests.helpers import clustered_pairs

pytestmark = pytest.mark.gpu
HP = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}
COLS = ("D_n", "dists_n", "dists_tr_n", "D_m", "dists_m", "dists_tr_m")


@pytest.mark.parametrize("dist_type", ["cosine", "euclidean"])
def test_reference_loop_through_faiss_shim_equals_score_pairs(dist_type):
    import lemon_b200
    from oracle import lemon_oracle as O
    n_train, cap, nq, d, k = 60_000, 50_000, 2048, 128, 30
    x, y, _, _ = clustered_pairs(n_train, d, n_clusters=300, seed=5, noise_frac=0.0)
    rng = np.random.RandomState(11)
    raw = lambda a: (a * rng.uniform(0.5, 2.0, (a.shape[0], 1))).astype(np.float32)      # the loop normalises (utils.py:39-40)
    x, y = raw(x), raw(y)
    idx = lemon_b200.subsample_db(n_train, cap, np.random.RandomState(3))                # run_lemon.py:122-127
    assert len(idx) == cap and not (np.diff(idx) > 0).all()                             # unsorted subsample
    qid = lemon_b200.query_in_db_from_indices(n_train, idx)
    assert ((qid >= 0).sum() == cap) and (idx[qid[qid >= 0]] == np.nonzero(qid >= 0)[0]).all()
    assert (qid == O.query_in_db_from_indices(n_train, idx)).all()

    # ---- the reference's loop, its `faiss` and its scoring function replaced through the documented seams
    saved = sys.modules.get("faiss")
    try:
        shim = lemon_b200.install_faiss_shim()
        import faiss                                               # what run_lemon.py:14 does
        assert faiss is shim
        ref_utils = lemon_b200.patch_reference_metrics(types.SimpleNamespace())
        df = O.reference_cpu_scorer(x[:nq], y[:nq], x[idx], y[idx], k=k, dist_type=dist_type,
                                    train_indices_in_compr=idx, hparams=HP, batch_size=128, faiss_module=faiss,
                                    score_fn=ref_utils.calc_scores_given_hparams_vectorized)
    finally:
        if saved is not None:
            sys.modules["faiss"] = saved
        else:
            sys.modules.pop("faiss", None)

    # ---- the fused call
    out = lemon_b200.score_pairs(x[:nq], y[:nq], x[idx], y[idx], k=k, dist_type=dist_type, query_in_db=qid[:nq], hparams=HP)
    out = {c: t.cpu().numpy() for c, t in out.items()}
    I_n, I_m = np.stack(df["I_n"].values), np.stack(df["I_m"].values)
    same = (I_n == out["I_n"]).all(axis=1) & (I_m == out["I_m"]).all(axis=1)
    assert same.mean() > 0.99          # K0 on the GPU vs F.normalize on the CPU differ in the last bit: near-ties may swap
    for c in COLS:
        np.testing.assert_allclose(np.stack(df[c].values)[same], out[c][same], rtol=1e-5, atol=3e-6, err_msg=c)
    np.testing.assert_allclose(df["d_1"].values[same], out["d_1"][same], rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(df["score"].values[same], out["score"][same], rtol=1e-5, atol=1e-5)
    # rows outside the subsample keep rank 0 and lose the last neighbour (run_lemon.py:261-263)
    absent = np.nonzero(qid[:nq] < 0)[0]
    assert len(absent) > 100 and not (out["I_n"][qid[:nq] >= 0, 0] == qid[:nq][qid[:nq] >= 0]).any()


def test_host_output_streaming_equals_device_outputs():
    """score_pairs_sharded(host shards, host_out=..., index_dtype=int32): records computed in parts, each part's
    device->host copy overlapping the next part; must be bit-identical to the device-resident call."""
    import torch
    import lemon_b200
    from lemon_b200 import dist as ldist
    n, d, k = 30_011, 512, 30
    x, y, _, _ = clustered_pairs(n, d, n_clusters=200, seed=8, noise_frac=0.3)
    dev = torch.device("cuda", 0)
    xd, yd = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    sc = lemon_b200.get_scorer(0)
    ref = ldist.score_pairs_sharded(xd, yd, n, k=k, hparams=HP, scorer=sc)
    host_out = {}
    got = ldist.score_pairs_sharded(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(), n, k=k, hparams=HP,
                                    scorer=sc, host_out=host_out, index_dtype=torch.int32, d2h_parts=5)
    assert got["rows"] == (0, n)
    for c, t in ref.items():
        if c == "rows":
            continue
        g = got[c]
        assert not g.is_cuda and g.is_pinned()
        if c in ("I_n", "I_m"):
            assert g.dtype == torch.int32
            assert (g.to(torch.int64) == t.cpu()).all()
        else:
            assert torch.equal(g, t.cpu()), c
    # the pinned buffers are reused by the next call
    ptr = host_out["score"].data_ptr()
    ldist.score_pairs_sharded(xd, yd, n, k=k, hparams=HP, scorer=sc, host_out=host_out, index_dtype=torch.int32)
    assert host_out["score"].data_ptr() == ptr


def test_sharded_driver_with_label_ids_in_the_text_gather():
    """Discrete text metric: the label ids travel as extra columns of the text shard (one all-gather per modality, no
    third collective) and K0 reads the embedding columns through its row stride; equals score_pairs with labels."""
    import torch
    import lemon_b200
    from lemon_b200 import dist as ldist
    x, y, lab, _ = clustered_pairs(9000, 96, n_clusters=40, seed=9, dup_text_classes=10)
    dev = torch.device("cuda", 0)
    sc = lemon_b200.get_scorer(0)
    lab_t = torch.from_numpy(lab.astype(np.int32))
    got = ldist.score_pairs_sharded(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), 9000, k=12, hparams=HP, scorer=sc,
                                    text_label_ids_local=lab_t.to(dev))
    ref = lemon_b200.score_pairs(x, y, k=12, query_in_db=np.arange(9000), hparams=HP, text_label_ids_q=lab, text_label_ids_db=lab)
    for c, t in ref.items():
        assert torch.equal(got[c], t), c
    assert set(np.unique(got["dists_n"].cpu().numpy())) <= {0.0, 1.0}


def test_embedding_handoff_equals_score_pairs_on_the_same_embeddings():
    """encoder outputs are written into the device shard and the database is staged chunk by chunk
    while the encoder runs (lemon_b200.handoff); the result is bit-identical to score_pairs_sharded on the embeddings
    the same encoders produce in one go.  The 'encoders' are fixed random projections of structured inputs, so the
    embeddings are CLIP-like (clustered, non-degenerate) and the certificate path is exercised, not bypassed."""
    import torch
    import lemon_b200
    from lemon_b200 import dist as ldist, handoff
    from lemon_b200.scoring import count_uncertified
    dev = torch.device("cuda", 0)
    n, latent, d, k, bs = 20_000, 48, 512, 30, 384
    g = torch.Generator(device=dev).manual_seed(5)
    cen = torch.randn(200, latent, generator=g, device=dev)
    z = torch.randint(0, 200, (n,), generator=g, device=dev)
    pix = cen[z] + 0.5 * torch.randn(n, latent, generator=g, device=dev)              # "images"
    tok = cen[z] + 0.5 * torch.randn(n, latent, generator=g, device=dev)              # "captions" of the same concept
    Wi = torch.randn(latent, d, generator=g, device=dev)
    Wt = torch.randn(latent, d, generator=g, device=dev)
    enc_img = lambda p: torch.tanh(p @ Wi).to(torch.bfloat16)                          # encoders emit bf16 (autocast), like CLIP
    enc_txt = lambda t: torch.tanh(t @ Wt + 0.3 * (t @ Wi)).to(torch.bfloat16)
    batches = [(pix[b:b + bs], tok[b:b + bs]) for b in range(0, n, bs)]
    sc = lemon_b200.get_scorer(0)
    got = handoff.extract_and_score(batches, enc_img, enc_txt, n, k=k, hparams=HP, scorer=sc, gather_chunks=5)
    info = sc.last_info
    assert info["img"]["path"] == "tc" and count_uncertified(info["img"]) < 0.05 * n     # certified, not all-fallback
    img = torch.cat([enc_img(p) for p, _ in batches]).float()
    txt = torch.cat([enc_txt(t) for _, t in batches]).float()
    assert torch.equal(got["shards"][0][:n], img) and torch.equal(got["shards"][1][:n], txt)
    ref = ldist.score_pairs_sharded(img, txt, n, k=k, hparams=HP, scorer=sc)
    for c, t in ref.items():
        if c != "rows":
            assert torch.equal(got[c], t), c


def test_sharded_driver_host_inputs_with_labels_and_host_outputs():
    """The C1-shaped bench path end to end: pinned HOST shards, label ids, mass-duplicate text side (10 distinct
    prompts -> de-duplicated search), streamed host outputs; equals the device-resident call bit for bit."""
    import torch
    import lemon_b200
    from lemon_b200 import dist as ldist
    x, y, lab, _ = clustered_pairs(12_000, 512, n_clusters=60, seed=10, dup_text_classes=10)
    dev = torch.device("cuda", 0)
    sc = lemon_b200.get_scorer(0)
    lab_t = torch.from_numpy(lab.astype(np.int32))
    ref = ldist.score_pairs_sharded(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), 12_000, k=30, hparams=HP, scorer=sc,
                                    text_label_ids_local=lab_t.to(dev))
    assert sc.last_info["txt"].get("n_unique") == 10
    got = ldist.score_pairs_sharded(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(), 12_000, k=30, hparams=HP,
                                    scorer=sc, text_label_ids_local=lab_t, host_out={})
    for c, t in ref.items():
        if c != "rows":
            assert torch.equal(got[c], t.cpu()), c
