"""The drop-in boundary is a C ABI: a plain C program (no Python, no torch) compiles against include/lemon_b200.h, links
liblemon_b200.so and runs.  Without a GPU the library must refuse loudly (no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc unavailable")
    from lemon_b200 import LIB_PATH
    assert os.path.exists(LIB_PATH), "build the library first (python -m lemon_b200.build)"
    exe = str(tmp_path / "c_abi_host")
    libdir = os.path.dirname(LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_host.c"),
           "-L" + libdir, "-llemon_b200", "-Wl,-rpath," + libdir, "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_c_host_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "lemon_version 100" in p.stdout
    if not torch.cuda.is_available():
        assert p.returncode == 3 and "no CPU fallback exists" in p.stdout


@pytest.mark.gpu
def test_c_host_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "context created" in p.stdout and "normalize_cast(NULL input) -> -1" in p.stdout
