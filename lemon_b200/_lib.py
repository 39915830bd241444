"""ctypes binding of liblemon_b200.so (include/lemon_b200.h).

There is no CPU fallback: if the shared library is missing or the device is not
sm_100, every entry point raises.  The library is built in-tree by
``python -m lemon_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LEMON_B200_LIB: developer override (e.g. a build with -DLEMON_TC_EXPERIMENT for kernel experiments)
LIB_PATH = os.environ.get("LEMON_B200_LIB") or os.path.join(_HERE, "liblemon_b200.so")

c_f32p = C.c_void_p
c_i32p = C.c_void_p
c_i64p = C.c_void_p
c_f64p = C.c_void_p

# name -> (restype, argtypes); mirrors include/lemon_b200.h declaration by declaration
SIGNATURES = {
    "lemon_version": (C.c_int, []),
    "lemon_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "lemon_ctx_destroy": (C.c_int, [C.c_void_p]),
    "lemon_last_error": (C.c_char_p, [C.c_void_p]),
    "lemon_launch_count": (C.c_int64, [C.c_void_p]),
    "lemon_normalize_cast": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, c_f32p, c_f32p, C.c_int64,
                                       C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "lemon_rowwise_dist": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "lemon_knn_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                       C.c_int, C.c_int, c_i64p, c_i32p, c_f32p, C.c_void_p]),
    "lemon_rerank": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i64p, c_i32p, c_f32p, c_f32p, c_f32p, C.c_float,
                               C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, c_i32p, c_f32p, c_i32p, c_i32p,
                               c_i32p, C.c_void_p]),
    "lemon_split_cast": (C.c_int, [C.c_void_p, c_f32p, C.c_void_p, c_f32p, c_f32p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "lemon_knn_exact": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i32p, c_i32p, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_int, C.c_int, C.c_int, c_f32p, c_i32p, C.c_void_p]),
    "lemon_score": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_f32p, c_i32p,
                              c_i64p, c_i32p, c_i32p, c_f32p, c_i32p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.POINTER(C.c_double), c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p,
                              C.c_void_p, C.c_int, C.c_int, c_f64p, c_f64p, c_f64p, C.c_void_p]),
    "lemon_hash_rows": (C.c_int, [C.c_void_p, c_f32p, C.c_int64, C.c_int, c_i64p, C.c_void_p]),
    "lemon_dedup_count_workspace_bytes": (C.c_int64, [C.c_int64]),
    "lemon_dedup_count": (C.c_int, [C.c_void_p, c_f32p, C.c_int64, C.c_int, C.c_void_p, c_i32p, C.c_void_p]),
    "lemon_dedup_workspace_bytes": (C.c_int64, [C.c_int64]),
    "lemon_dedup_build": (C.c_int, [C.c_void_p, c_f32p, C.c_int64, C.c_int, C.c_void_p, c_i32p, c_i32p, c_i64p, c_i32p,
                                    C.c_void_p]),
    "lemon_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, c_i32p, c_i32p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "lemon_keep_lowest_workspace_bytes": (C.c_int64, [C.c_int64]),
    "lemon_keep_lowest": (C.c_int, [C.c_void_p, c_f64p, C.c_int64, C.c_int64, C.c_void_p, c_i64p, c_f64p, C.c_void_p]),
    "lemon_expand_groups": (C.c_int, [C.c_void_p, c_f32p, c_i32p, c_i64p, c_i32p, C.c_int64, C.c_int, C.c_int, c_f32p,
                                      c_i32p, C.c_void_p]),
    "lemon_f1_grid": (C.c_int, [C.c_void_p, c_f64p, c_f64p, c_f64p, C.c_void_p, C.c_int64, c_f64p, c_f64p, c_i32p,
                                C.c_int64, C.c_double, C.c_int, c_f64p, c_f64p, c_f64p, C.c_int64, C.c_void_p]),
    "lemon_discrepancy": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_i32p, c_i32p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    "lemon_combine_scores": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f64p, C.c_int64,
                                       C.c_int, C.POINTER(C.c_double), c_f64p, c_f64p, c_f64p, C.c_void_p]),
}

_lib = None


class LemonError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library and bind every symbol the header declares.  Does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LemonError(
            f"{LIB_PATH} is missing: build it with `python -m lemon_b200.build` "
            "(there is no CPU fallback for the LEMoN scoring path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """One lemon_ctx per device."""

    def __init__(self, device: int):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.lemon_ctx_create(int(device), C.byref(h))
        if rc != 0:
            why = {-2: "CUDA error / no such device", -3: "device is not sm_100 (B200 required)"}.get(rc, "error")
            raise LemonError(f"lemon_ctx_create(device={device}) failed: {why} (rc={rc}); no CPU fallback exists")
        self.handle = h
        self.device = int(device)

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self.lib.lemon_last_error(self.handle)
            raise LemonError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")

    def launch_count(self) -> int:
        return int(self.lib.lemon_launch_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.lemon_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts: dict[int, Context] = {}


def get_context(device: int) -> Context:
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx
