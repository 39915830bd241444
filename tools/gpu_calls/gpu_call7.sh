#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu7.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu7.log
tail -3 gpurun_out/r2_pytest_gpu7.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_c3_d.json 2> gpurun_out/r2_bench_c3_d.err; echo "c3 rc=$?"; tail -3 gpurun_out/r2_bench_c3_d.err
python - <<'P'
import json
b = json.loads(open("gpurun_out/r2_bench_c3_d.json").read().strip().splitlines()[-1])
print("c3_d", b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["roofline"]["traffic"], b["clocks"])
P
STEP="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
$STEP > gpurun_out/r2_plain_step7.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_final.csv $STEP > gpurun_out/r2_ncu_step7.log 2>&1
echo "launch list rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -2
