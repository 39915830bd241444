#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu15.log
for shape in "118000 118000 512" "370000 370000 512" "75776 1000000 768" "46250 370000 512"; do python tools/k1_launch.py $shape 40 3 | tail -1; done 2>&1 | tee gpurun_out/r2_k1_rawkeys.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_c3_g.json 2> gpurun_out/r2_bench_c3_g.err; echo "c3 rc=$?"; tail -3 gpurun_out/r2_bench_c3_g.err
python - <<'P'
import json
b = json.loads(open("gpurun_out/r2_bench_c3_g.json").read().strip().splitlines()[-1])
print("c3_g", b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["clocks"])
P
