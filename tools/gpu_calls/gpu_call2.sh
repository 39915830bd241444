#!/bin/bash
# round-2 GPU call 2: full GPU test suite, C2 / C1 bench lines, launch list of a C3 step, ncu captures of K1 (C3 and d=768 shapes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu2.log
tail -15 gpurun_out/r2_pytest_gpu2.log
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "c2 rc=$?"
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; echo "c1 rc=$?"
python - <<'P'
import json
for w in ("c2", "c1"):
    try:
        b = json.loads(open(f"gpurun_out/r2_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"], b["roofline"]["frac"], b["run_info"], b["clocks"])
    except Exception as e:
        print(w, "failed", e); print(open(f"gpurun_out/r2_bench_{w}.err").read()[-2000:])
P
STEP="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
$STEP > gpurun_out/r2_plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c3.csv $STEP > gpurun_out/r2_ncu_step.log 2>&1
echo "launch list rc=$?"
K1A="python tools/k1_launch.py 75776 370000 512"
$K1A > gpurun_out/r2_k1a_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c3shape $K1A > gpurun_out/r2_k1a_ncu.log 2>&1
echo "k1 c3 ncu rc=$?"; cat gpurun_out/r2_k1a_plain.log | tail -1
K1B="python tools/k1_launch.py 37888 1000000 768"
$K1B > gpurun_out/r2_k1b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_d768 $K1B > gpurun_out/r2_k1b_ncu.log 2>&1
echo "k1 d768 ncu rc=$?"; cat gpurun_out/r2_k1b_plain.log | tail -1
