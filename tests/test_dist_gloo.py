"""CPU tests of the row-sharding plumbing with the gloo backend (world_size 2).  The scorer is replaced
by a stand-in that evaluates the oracle, so the N>1 host path (shard bounds, padding, the two
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
    """Stand-in with the LemonScorer interface used by score_pairs_sharded."""

    def set_database(self, img_db, txt_db, dist_type, normalize, labels):
        self.db = (img_db.numpy(), txt_db.numpy(), dist_type, normalize)

    def score(self, img_q, txt_q, *, k, query_in_db, hparams, return_records, query_rows, text_label_ids_q):
        from oracle import lemon_oracle as O
        x, y, dist_type, normalize = self.db
        r0, r1 = query_rows
        assert query_in_db.tolist() == list(range(r0, r1))
        out = O.lemon_oracle(x[r0:r1], y[r0:r1], x, y, k=k, dist_type=dist_type, query_in_db=query_in_db.numpy(),
                             hparams=hparams, normalize=normalize)
        return {c: torch.from_numpy(np.asarray(out[c])) for c in ("score", "I_n", "I_m", "d_1")}


def _worker(rank, world, port, n, d, k, q):
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
        out = ldist.score_pairs_sharded(pad(x), pad(y), n, k=k, hparams=hp, scorer=OracleScorer())
        q.put((rank, out["rows"], out["score"].numpy(), out["I_n"].numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_equals_single_rank_gloo():
    from oracle import lemon_oracle as O
    n, d, k, world = 101, 16, 4, 2           # odd n: the last shard is padded
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, k, q)) for r in range(world)]
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
    full = O.lemon_oracle(x, y, x, y, k=k, query_in_db=np.arange(n), hparams=hp)
    score = np.concatenate([r[2] for r in res])
    I_n = np.concatenate([r[3] for r in res])
    assert res[0][1] == (0, 51) and res[1][1] == (51, 101)
    assert np.array_equal(score, full["score"])          # bitwise: rows are independent of the sharding
    assert np.array_equal(I_n, full["I_n"])
