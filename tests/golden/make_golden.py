"""Generate tests/golden/*.npz by running the reference's OWN functions live.

Run in the builder container only (needs /root/reference):
    python tests/golden/make_golden.py
Functions executed from the reference (imported through oracle/ref_live.py, nothing copied):
    lib.metrics.utils.calc_scores_given_hparams             (utils.py:21-45)
    lib.metrics.utils.calc_scores_given_hparams_vectorized  (utils.py:47-82, numpy and torch_arr)
    lib.utils.utils.normalize_vectors                       (utils.py:39-40)
The fixtures hold the inputs and the reference's outputs; tests/test_oracle.py pins the
oracle to them, tests/test_gpu_parity.py pins the CUDA path to them.
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_live  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
COLS = ("D_n", "D_m", "dists_tr_n", "dists_tr_m", "dists_n", "dists_m")

HPARAM_SETS = [
    {"beta": 5, "gamma": 5, "tau_1_n": 0.1, "tau_2_n": 5, "tau_1_m": 0.1, "tau_2_m": 5},  # train_clip_from_scratch.py:102-109
    {"beta": 0.0, "gamma": 0.0, "tau_1_n": 0.0, "tau_2_n": 0.0, "tau_1_m": 0.0, "tau_2_m": 0.0},
    {"beta": 1.0, "gamma": 1.0, "tau_1_n": 0.0, "tau_2_n": 0.0, "tau_1_m": 0.0, "tau_2_m": 0.0},
    {"beta": 17.5, "gamma": 0.25, "tau_1_n": 10.0, "tau_2_n": 0.5, "tau_1_m": 2.0, "tau_2_m": 20.0},
]


def random_records(rng, n, k):
    # value ranges as the cosine path produces them: D = -<a,b> in [-1,1], dists in [0,2]
    rec = {
        "D_n": -rng.uniform(0.2, 1.0, (n, k)), "D_m": -rng.uniform(0.2, 1.0, (n, k)),
        "dists_tr_n": rng.uniform(0.0, 1.5, (n, k)), "dists_tr_m": rng.uniform(0.0, 1.5, (n, k)),
        "dists_n": rng.uniform(0.0, 2.0, (n, k)), "dists_m": rng.uniform(0.0, 2.0, (n, k)),
    }
    rec = {c: v.astype(np.float32) for c, v in rec.items()}
    rec["d_1"] = rng.uniform(0.0, 2.0, n).astype(np.float32).astype(np.float64)  # `.item()` of an fp32 tensor
    return rec


def to_df(rec):
    n = len(rec["d_1"])
    return pd.DataFrame([{**{c: rec[c][i] for c in COLS}, "d_1": float(rec["d_1"][i])} for i in range(n)])


def main():
    mu = ref_live.import_reference_metrics()
    uu = ref_live.import_reference_utils()
    rng = np.random.RandomState(20261018)

    for tag, (n, k) in {"k30": (96, 30), "k5": (64, 5), "k1": (16, 1)}.items():
        rec = random_records(rng, n, k)
        df = to_df(rec)
        out = dict(rec)
        for h, hp in enumerate(HPARAM_SETS):
            s, dn, dm = mu.calc_scores_given_hparams_vectorized(df, hp, return_dn=True)
            out[f"vec_scores_{h}"], out[f"vec_dn_{h}"], out[f"vec_dm_{h}"] = (np.asarray(a) for a in (s, dn, dm))
            s, dn, dm = mu.calc_scores_given_hparams(df, hp, return_dn=True)
            out[f"loop_scores_{h}"], out[f"loop_dn_{h}"], out[f"loop_dm_{h}"] = (np.asarray(a) for a in (s, dn, dm))
            s = mu.calc_scores_given_hparams_vectorized(df, hp, torch_arr=True)
            out[f"torch_scores_{h}"] = s.numpy()
        out["hparams"] = np.array([[hp[key] for key in ("beta", "gamma", "tau_1_n", "tau_2_n", "tau_1_m", "tau_2_m")]
                                   for hp in HPARAM_SETS], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"scores_{tag}.npz"), **out)

    x = rng.standard_normal((40, 96)).astype(np.float32) * rng.uniform(0.01, 30, (40, 1)).astype(np.float32)
    x[3] = 0.0      # zero row: eps clamp of F.normalize
    x[4] *= 1e-20   # tiny row
    y = uu.normalize_vectors(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(HERE, "normalize.npz"), x=x, y=y)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
