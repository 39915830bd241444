#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_seams.py -m gpu -q -x > gpurun_out/r2_pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu10.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r2_bench_c3_f.json 2> gpurun_out/r2_bench_c3_f.err; echo "c3 n1 rc=$?"; tail -3 gpurun_out/r2_bench_c3_f.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-parity > gpurun_out/r2_bench_c3_n2_c.json 2> gpurun_out/r2_bench_c3_n2_c.err; echo "c3 n2 rc=$?"; tail -3 gpurun_out/r2_bench_c3_n2_c.err
$TR bench.py --gpus 2 --workload c1 --steps 10 --warmup 3 --no-parity > gpurun_out/r2_bench_c1_n2.json 2> gpurun_out/r2_bench_c1_n2.err; echo "c1 n2 rc=$?"; tail -3 gpurun_out/r2_bench_c1_n2.err
$TR bench.py --gpus 2 --workload c2 --steps 10 --warmup 3 --no-parity > gpurun_out/r2_bench_c2_n2.json 2> gpurun_out/r2_bench_c2_n2.err; echo "c2 n2 rc=$?"; tail -3 gpurun_out/r2_bench_c2_n2.err
python - <<'P'
import json
for w in ("c3_f", "c3_n2_c", "c1_n2", "c2_n2"):
    try:
        b = json.loads(open(f"gpurun_out/r2_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["roofline"]["frac"], b["clocks"]["sm_mhz"])
    except Exception as e:
        print(w, "failed", e)
P
