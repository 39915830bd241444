"""CPU tests: the C-ABI library loads and exports every symbol include/lemon_b200.h declares, the host
planner behaves, and the product path fails loudly (no CPU fallback) without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()                                                   # nvcc cross-compiles without a GPU
    from lemon_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "lemon_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lemon_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 12
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lemon_version() >= 100


def test_ctypes_signatures_match_the_header():
    """The ctypes table mirrors include/lemon_b200.h prototype by prototype: same number of parameters, and the same
    kind (pointer / 64-bit integer / 32-bit integer / floating point) in every position -- a drifted binding would
    pass garbage to the kernels without any error."""
    import ctypes as C
    from lemon_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "lemon_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b[\w\s\*]+?\b(lemon_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == len(_lib.SIGNATURES)

    def kind_c(decl: str) -> str:
        decl = " ".join(decl.split())
        if "*" in decl:
            return "ptr"
        if decl.startswith(("int64_t", "long long")):
            return "i64"
        if decl.startswith(("float", "double")):
            return decl.split()[0]
        assert decl.startswith("int"), decl
        return "i32"

    def kind_py(t) -> str:
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int64: "i64", C.c_int: "i32", C.c_float: "float", C.c_double: "double"}[t]

    for name, params in protos:
        params = [p for p in params.split(",") if p.strip() and p.strip() != "void"]
        argtypes = _lib.SIGNATURES[name][1]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for i, (pc, pt) in enumerate(zip(params, argtypes)):
            assert kind_c(pc) == kind_py(pt), (name, i, pc.strip(), pt)


def test_sass_has_blackwell_tensor_and_tma_ops():
    import subprocess
    from lemon_b200 import LIB_PATH
    sass = subprocess.run(["cuobjdump", "-sass", LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass      # tcgen05.mma, TMA, tcgen05.ld
    assert "HMMA.16816" not in sass                                        # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import lemon_b200
    with pytest.raises(lemon_b200.LemonError):
        lemon_b200.LemonScorer()
    with pytest.raises(lemon_b200.LemonError):
        lemon_b200.score_pairs(torch.zeros(4, 8), torch.zeros(4, 8), k=2)


def test_plan_segments():
    from lemon_b200 import plan_segments
    assert plan_segments(3_300_000, 3_300_000, 148, 2, 768) == 1      # many waves: no split
    assert plan_segments(412_500, 3_300_000, 148, 2, 768) == 1
    assert plan_segments(118_000, 118_000, 148, 2, 512) == 1          # per-item start-up outweighs the wave gain
    assert plan_segments(2_048, 3_300_000, 148, 2, 768) > 1           # 8 pair tiles for 74 pairs, long DB: split it
    assert plan_segments(100, 2000, 148, 2, 512) == 1                 # tiny DB: never split below 4096 columns
    for nq in (1, 1000, 50_000):
        for m in (10, 50_000, 3_300_000):
            n = plan_segments(nq, m, 148, 2, 512)
            assert 1 <= n <= 32 and (n == 1 or m // n >= 4096)
    # calibration points (profiles/r02_plan_calibrate.log): one row tile wants one full wave of short items
    assert plan_segments(128, 370_000, 148, 2, 512) == 32 and plan_segments(1024, 370_000, 148, 2, 512) in (16, 18)
    assert plan_segments(4096, 370_000, 148, 2, 512) == 4 and plan_segments(8448, 370_000, 148, 2, 512) == 2
    assert plan_segments(128, 50_000, 148, 2, 512) in (8, 12)


def test_faiss_shim_installs():
    import sys
    import lemon_b200
    old = sys.modules.get("faiss")
    try:
        mod = lemon_b200.install_faiss_shim()
        import faiss
        assert faiss is mod and hasattr(faiss, "IndexFlatIP") and hasattr(faiss, "IndexFlatL2")
    finally:
        if old is not None:
            sys.modules["faiss"] = old
        else:
            sys.modules.pop("faiss", None)


def test_plan_tail():
    from lemon_b200 import plan_tail
    # C2 on one GPU: 461 pair tiles on 74 pairs = 6 full rounds + 17 tiles -> the 17 go to a 4-segment launch
    n_main, nseg = plan_tail(118_000, 118_000, 148, 2)
    assert n_main == 6 * 74 * 256 and nseg == 4
    assert plan_tail(74 * 256 * 3, 118_000, 148, 2) == (74 * 256 * 3, 1)       # exact multiple: nothing to fix
    assert plan_tail(1000, 118_000, 148, 2) == (1000, 1)                       # less than one round: plan_segments' job
    assert plan_tail(74 * 256 + 60 * 256, 118_000, 148, 2)[1] == 1             # tail nearly a full round: leave it
    n_main, nseg = plan_tail(118_000, 10_000, 148, 2)
    assert nseg == 2 and 10_000 // nseg >= 4096                                # short DB: fewer segments


def test_plan_parts_bounds_candidate_memory():
    """K1 launches of one kNN call: contiguous cover of the rows, whole rounds of query tiles per cut, at most
    2**19 rows (8.6 GB of candidate lists) per launch; small problems stay one launch as planned by plan_tail."""
    from lemon_b200.scoring import plan_parts, plan_tail
    assert plan_parts(118_000, 118_000, 148, 2) == [(0, 6 * 74 * 256, 1), (6 * 74 * 256, 118_000, 4)]   # C2 unchanged
    assert plan_parts(1000, 5000, 148, 2) == [(0, 1000, None)]
    assert plan_parts(412_500, 3_300_000, 148, 2) == [(0, 412_500, None)]                              # C4 on 8 GPUs
    for nq, m in ((3_300_000, 3_300_000), (18944 * 27 + 5, 100_000), (18944 * 30, 1_000_000), (600_000, 50_000)):
        parts = plan_parts(nq, m, 148, 2)
        assert parts[0][0] == 0 and parts[-1][1] == nq
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(b - a <= (1 << 19) + 18944 for a, b, _ in parts)
        assert all((b - a) % 18944 == 0 for a, b, _ in parts[:-2])                                     # whole rounds
        n_main, nseg_tail = plan_tail(nq, m, 148, 2)
        if n_main < nq:
            assert parts[-1] == (n_main, nq, nseg_tail)


def test_db_cap_helpers_follow_the_reference():
    """run_lemon.py:122-127 (np.random.choice without replacement when the split exceeds the cap, arange otherwise) and
    :258,278 (membership of the sample in train_indices_in_compr) as one scatter; same answers as the oracle's."""
    import numpy as np
    import lemon_b200
    from oracle import lemon_oracle as O
    assert (lemon_b200.subsample_db(100, 50_000) == np.arange(100)).all()
    a = lemon_b200.subsample_db(1000, 300, np.random.RandomState(4))
    b = O.subsample_db(1000, 300, np.random.RandomState(4))
    assert (a == b).all() and len(set(a.tolist())) == 300 and not (np.diff(a) > 0).all()
    np.random.seed(81)                                            # run_lemon.py:81 seeds the GLOBAL numpy RNG
    c = lemon_b200.subsample_db(1000, 300)
    np.random.seed(81)
    assert (c == np.random.choice(np.arange(1000), 300, replace=False)).all()
    qid = lemon_b200.query_in_db_from_indices(1000, a)
    assert (qid == O.query_in_db_from_indices(1000, a)).all()
    for s in (0, 17, 999):
        assert (qid[s] >= 0) == (s in a) and (qid[s] < 0 or a[qid[s]] == s)


def test_accumulation_bounds():
    from lemon_b200.scoring import acc_eps_coef, acc_eps_coef_split
    assert abs(acc_eps_coef(768, 768) - (768 * 2.0 ** -23 + 30 * 2.0 ** -24)) < 1e-12       # gamma_n with truncating adds
    assert acc_eps_coef(512, 512) < acc_eps_coef(768, 768)
    # the split pass accumulates the two small cross terms first: its bound stays that of ONE d16-long product
    assert acc_eps_coef_split(768, 768) < 1.02 * acc_eps_coef(768, 768) + 2.0 ** -21
    assert acc_eps_coef_split(768, 768) < 0.4 * acc_eps_coef(3 * 768, 768)
