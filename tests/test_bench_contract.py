"""bench.py's reference arm and JSON contract, on the CPU (the GPU arm is exercised by the driver and by gpurun calls)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_config_match_the_gpu_arm():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0")       # torchrun exports OMP_NUM_THREADS=1: the arm must undo it
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                        "--warmup", "0", "--ref-queries", "96", "--gpus", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == json.loads(json.dumps(bench.bench_config(bench.WORKLOADS["tiny"], 1)))    # same object in both arms
    assert "sample" not in line["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_traffic_table_is_keyed_by_workload_and_gpu_count():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.k1_traffic("c4", 8)
    assert t and t["dram_bytes_per_launch"] > 1e11 and "412500" in t["note"]
    assert bench.k1_traffic("c3", 1)["dram_bytes_per_launch"] < 2e10
    assert bench.k1_traffic("c3", 4) is None                    # no capture of that launch shape: the bench line says null
