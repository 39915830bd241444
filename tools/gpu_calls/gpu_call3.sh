#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu3.log
tail -6 gpurun_out/r2_pytest_gpu3.log; grep -n "uncertified\|eps first" gpurun_out/r2_pytest_gpu3.log | head
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; echo "c1 rc=$?"; tail -3 gpurun_out/r2_bench_c1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_b.json 2> gpurun_out/r2_bench_c3_b.err; echo "c3 rc=$?"; tail -3 gpurun_out/r2_bench_c3_b.err
python - <<'P'
import json
for w in ("c1", "c3_b"):
    try:
        b = json.loads(open(f"gpurun_out/r2_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["run_info"], b["clocks"], b["gpu_launches"])
    except Exception as e:
        print(w, "failed", e)
P
python tools/uncert_rate.py 3300000 768 16384 > gpurun_out/r2_uncert_rate.jsonl 2> gpurun_out/r2_uncert_rate.err; echo "uncert rc=$?"; cat gpurun_out/r2_uncert_rate.jsonl; tail -3 gpurun_out/r2_uncert_rate.err
K1A="python tools/k1_launch.py 370000 370000 512"
$K1A > gpurun_out/r2_k1c3full_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c3full $K1A > gpurun_out/r2_k1c3full_ncu.log 2>&1
echo "k1 c3 full ncu rc=$?"; tail -1 gpurun_out/r2_k1c3full_plain.log
K1B="python tools/k1_launch.py 412500 3300000 768"
$K1B > gpurun_out/r2_k1c4rank_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c4rank $K1B > gpurun_out/r2_k1c4rank_ncu.log 2>&1
echo "k1 c4 rank ncu rc=$?"; tail -1 gpurun_out/r2_k1c4rank_plain.log
K2="python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
$K2 > gpurun_out/r2_k2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rerank_kernel|score_kernel" -s 6 -c 3 -o gpurun_out/r2_k2_c2 $K2 > gpurun_out/r2_k2_ncu.log 2>&1
echo "k2 ncu rc=$?"
