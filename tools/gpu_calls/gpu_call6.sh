#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu6.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu6.log
tail -3 gpurun_out/r2_pytest_gpu6.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_c.json 2> gpurun_out/r2_bench_c3_c.err; echo "c3 rc=$?"; tail -3 gpurun_out/r2_bench_c3_c.err
python bench.py --workload c2 --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_bench_c2_c.json 2> gpurun_out/r2_bench_c2_c.err; echo "c2 rc=$?"
python - <<'P'
import json
for w in ("c3_c", "c2_c"):
    try:
        b = json.loads(open(f"gpurun_out/r2_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["roofline"]["k1_share_of_step"], b["clocks"], b["gpu_launches"])
    except Exception as e:
        print(w, "failed", e)
P
K1A="python tools/k1_launch.py 370000 370000 512"
$K1A > gpurun_out/r2_k1c3full_paced_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c3full_paced $K1A > gpurun_out/r2_k1c3full_paced_ncu.log 2>&1
echo "k1 c3 full ncu rc=$?"; tail -1 gpurun_out/r2_k1c3full_paced_plain.log
K1B="python tools/k1_launch.py 412500 3300000 768"
$K1B > gpurun_out/r2_k1c4rank_paced_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c4rank_paced $K1B > gpurun_out/r2_k1c4rank_paced_ncu.log 2>&1
echo "k1 c4 rank ncu rc=$?"; tail -1 gpurun_out/r2_k1c4rank_paced_plain.log
K2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary --no-parity"
$K2 > gpurun_out/r2_k2c3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rerank_kernel|score_kernel" -s 3 -c 3 -o gpurun_out/r2_k2_c3 $K2 > gpurun_out/r2_k2c3_ncu.log 2>&1
echo "k2 ncu rc=$?"
