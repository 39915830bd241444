"""CPU test of the columnar -> legacy DataFrame adapter against the reference's own consumer of that
schema (calc_scores_given_hparams_vectorized reads the object columns, lib/metrics/utils.py:64-69)."""
import numpy as np
import pytest

from lemon_b200 import results
from oracle import lemon_oracle as O
from oracle import ref_live
from tests.helpers import clustered_pairs


def _oracle_out(n=150, k=7):
    x, y, _, mis = clustered_pairs(n, 48, n_clusters=9, seed=8, noise_frac=0.3)
    out = O.lemon_oracle(x, y, x, y, k=k, query_in_db=np.arange(n), hparams=O.CC3M_HPARAMS)
    return out, mis


def test_schema_and_roundtrip():
    out, mis = _oracle_out()
    df = results.records_to_dataframe(out, "train", idx_offset=10, is_mislabel=mis, noisy_label=np.arange(150))
    assert list(df.columns) == ["sset", "idx", "noisy_label", "is_mislabel", "is_correct_label", "d_1",
                                "dists_n", "D_n", "dists_tr_n", "dists_m", "D_m", "dists_tr_m"]
    assert df["idx"].iloc[0] == 10 and (df["sset"] == "train").all()
    assert isinstance(df["D_n"].iloc[3], np.ndarray) and df["D_n"].iloc[3].dtype == np.float32 and df["D_n"].iloc[3].shape == (7,)
    assert (df["is_mislabel"] + df["is_correct_label"] == 1).all()
    rec = results.dataframe_to_records(df)
    for c in results.RECORD_COLS:
        np.testing.assert_array_equal(rec[c], np.asarray(out[c], np.float32))
    s, _, _ = O.calc_scores_vectorized(rec, O.CC3M_HPARAMS)
    np.testing.assert_allclose(s, out["score"], rtol=1e-6)


@pytest.mark.skipif(not ref_live.available(), reason="/root/reference not mounted (GPU box)")
def test_reference_consumes_the_adapter_output():
    mu = ref_live.import_reference_metrics()
    out, mis = _oracle_out()
    df = results.records_to_dataframe(out, "val", is_mislabel=mis)
    s, dn, dm = mu.calc_scores_given_hparams_vectorized(df, O.CC3M_HPARAMS, return_dn=True)
    np.testing.assert_allclose(s, out["score"], rtol=1e-5)
    s2 = mu.calc_scores_given_hparams(df, O.CC3M_HPARAMS)
    np.testing.assert_allclose(np.asarray(s2), out["score"], rtol=1e-5)
    # the reference's F1 objective runs on it (hyper-parameter search consumer, utils.py:286-296)
    f1 = mu.optimize_f1_efficient(df["is_mislabel"].values, s)
    assert 0.0 <= f1 <= 1.0
