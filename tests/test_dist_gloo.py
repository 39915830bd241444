"""CPU tests of the row-sharding plumbing with the gloo backend (world_size 2).  The scorer is replaced
by a CPU scorer with the same staged interface that evaluates the oracle, so the N>1 host path (shard bounds, padding, the two
all-gathers, global query ids) is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lemon_b200 import dist as ldist


def test_shard_bounds_cover_everything():
    for n in (1, 7, 118000, 3300000):
        for world in (1, 2, 3, 8):
            spans = [ldist.shard_bounds(n, world, r) for r in range(world)]
            per = spans[0][2]
            assert all(s[2] == per for s in spans) and per * world >= n
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(s[1] - s[0] <= per for s in spans)


class OracleScorer:
    """CPU scorer with the staged interface score_pairs_sharded drives (prepare_db / finish_db / knn / rowwise_dist /
    emit on ``Prepared`` operands), every stage evaluated by the oracle."""
    device = torch.device("cpu")

    def __init__(self):
        self.last_info = {}

    def prepare_db(self, x, normalize=True, defer_dedup=False):
        from lemon_b200.scoring import Prepared
        from oracle import lemon_oracle as O
        a = x.numpy()
        a = O.normalize_vectors(a) if normalize else a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
        return Prepared(t, None, torch.zeros(len(a), 4), torch.zeros(4), t.shape[0], t.shape[1], t.shape[1])

    def finish_db(self, p):
        return p

    def knn(self, q, db, kp, metric):
        from oracle import lemon_oracle as O
        D, I = O.knn_search(q.f32.numpy(), db.f32.numpy(), kp, "ip" if metric == 0 else "l2")
        self.last_info = {"path": "oracle"}
        return torch.from_numpy(D), torch.from_numpy(I.astype(np.int32))

    def rowwise_dist(self, a, b, metric):
        from oracle import lemon_oracle as O
        return torch.from_numpy(O.dists_tr(a.numpy(), b.numpy(), "cosine" if metric == 0 else "euclidean"))

    def emit(self, xq, yq, xdb, ydb, dists_tr, topn, topm, *, k, kp, metric, qid, hparams, lab_q=None, lab_db=None, **_):
        from oracle import lemon_oracle as O
        dist_type = "cosine" if metric == 0 else "euclidean"
        in_db = qid.numpy() >= 0
        Dn, In = O.apply_self_exclusion(topn[0].numpy(), topn[1].numpy().astype(np.int64), in_db)
        Dm, Im = O.apply_self_exclusion(topm[0].numpy(), topm[1].numpy().astype(np.int64), in_db)
        rec = O.build_records(xq.f32.numpy(), yq.f32.numpy(), xdb.f32.numpy(), ydb.f32.numpy(), Dn, In, Dm, Im, dist_type,
                              None if lab_q is None else lab_q.numpy(), None if lab_db is None else lab_db.numpy())
        score, sn, sm = O.calc_scores_vectorized(rec, hparams)
        return {"score": torch.from_numpy(score), "I_n": torch.from_numpy(In), "I_m": torch.from_numpy(Im),
                "d_1": torch.from_numpy(rec["d_1"])}


def _worker(rank, world, port, n, d, k, q, with_labels=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(0)
        x = rng.standard_normal((n, d)).astype(np.float32)
        y = rng.standard_normal((n, d)).astype(np.float32)
        r0, r1, per = ldist.shard_bounds(n, world, rank)
        pad = lambda a: torch.from_numpy(np.concatenate([a[r0:r1], np.full((per - (r1 - r0), d), 7.0, np.float32)]))
        hp = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}
        g = ldist.allgather_rows(pad(x), n)
        assert g.shape == (n, d) and np.array_equal(g.numpy(), x)       # padding never enters the DB
        lab_local = None
        if with_labels:      # label ids travel as extra columns of the text all-gather (no third collective)
            lab = (np.arange(n) * 7 % 5).astype(np.int32)
            lab_local = torch.from_numpy(np.concatenate([lab[r0:r1], np.full(per - (r1 - r0), -9, np.int32)]))
        out = ldist.score_pairs_sharded(pad(x), pad(y), n, k=k, hparams=hp, scorer=OracleScorer(), text_label_ids_local=lab_local)
        q.put((rank, out["rows"], out["score"].numpy(), out["I_n"].numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("with_labels", [False, True])
def test_sharded_equals_single_rank_gloo(with_labels):
    from oracle import lemon_oracle as O
    n, d, k, world = 101, 16, 4, 2           # odd n: the last shard is padded
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, k, q, with_labels)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    x = rng.standard_normal((n, d)).astype(np.float32)
    y = rng.standard_normal((n, d)).astype(np.float32)
    hp = {"beta": 5.0, "gamma": 5.0, "tau_1_n": 0.1, "tau_2_n": 5.0, "tau_1_m": 0.1, "tau_2_m": 5.0}
    lab = (np.arange(n) * 7 % 5).astype(np.int32) if with_labels else None
    full = O.lemon_oracle(x, y, x, y, k=k, query_in_db=np.arange(n), hparams=hp, text_label_ids_q=lab, text_label_ids_db=lab)
    score = np.concatenate([r[2] for r in res])
    I_n = np.concatenate([r[3] for r in res])
    assert res[0][1] == (0, 51) and res[1][1] == (51, 101)
    assert np.array_equal(score, full["score"])          # bitwise: rows are independent of the sharding
    assert np.array_equal(I_n, full["I_n"])
