#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu16.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_h.json 2> gpurun_out/r2_bench_c3_h.err; echo "c3 rc=$?"; tail -3 gpurun_out/r2_bench_c3_h.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_h.json 2> gpurun_out/r2_bench_ref_h.err; echo "ref rc=$?"
python - <<'P'
import json
b = json.loads(open("gpurun_out/r2_bench_c3_h.json").read().strip().splitlines()[-1])
print("c3_h", b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["roofline"]["k1_launches_per_step"], b["run_info"], b["clocks"])
print(json.dumps(b["secondary"]))
r = json.loads(open("gpurun_out/r2_bench_ref_h.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_baseline"])
P
python __graft_entry__.py --smoke 2>&1 | tail -1
