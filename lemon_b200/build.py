"""Build liblemon_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m lemon_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblemon_b200.so")
SOURCES = ["capi.cu", "k0_normalize.cu", "k1_knn_exact.cu", "k1_knn_tc.cu", "k2_rerank.cu", "k2_score.cu", "k3_dedup.cu", "k4_hparam.cu", "k5_discrepancy.cu"]
EXTRA_FLAGS = {"k4_hparam.cu": ["-fmad=false"]}   # float64 operation order must follow the CPU code
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--use_fast_math=false"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lemon_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False, out_path: str | None = None) -> str:
    """out_path: alternative output path (objects go to a sibling directory), for experimental builds."""
    if out_path is None and not force and not needs_build():
        return OUT
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("LEMON_BUILD_DEFS", "").split()      # e.g. -DLEMON_TC_PROFILE (in-kernel clock counters of K1)
    objs = []
    procs = []
    objdir = os.path.join(HERE, "build") if out_path is None else out_path + ".objs"
    os.makedirs(objdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} (exit {p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    target = OUT if out_path is None else out_path
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target, *objs]
    subprocess.run(cmd, check=True)
    return target


if __name__ == "__main__":
    out_arg = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out_path=out_arg))
