#!/bin/bash
# 8 GPUs: C3 strong scaling point, the north-star configuration C4 (3.3 M pairs, 768-d), bitwise check, config 5 hand-off
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n8.json 2> gpurun_out/r2_bench_c3_n8.err; echo "c3 n8 rc=$?"; tail -2 gpurun_out/r2_bench_c3_n8.err
$TR bench.py --gpus 8 --workload c4 --steps 3 --warmup 3 --parity-rows 256 > gpurun_out/r2_bench_c4_n8.json 2> gpurun_out/r2_bench_c4_n8.err; echo "c4 n8 rc=$?"; tail -3 gpurun_out/r2_bench_c4_n8.err
python - <<'P'
import json
for w in ("c3_n8", "c4_n8"):
    try:
        b = json.loads(open(f"gpurun_out/r2_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"], b["roofline"]["frac"], b["roofline"]["k1_share_of_step"], b["run_info"], b["clocks"])
    except Exception as e:
        print(w, "failed", e)
P
$TR tools/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n8.log 2>&1; echo "check rc=$?"; grep MULTI_GPU gpurun_out/r2_multi_gpu_check_n8.log
$TR tools/e2e_clip.py --encoder clip --pairs 118000 > gpurun_out/r2_e2e_clip_n8.json 2> gpurun_out/r2_e2e_clip_n8.err; echo "e2e clip rc=$?"; cat gpurun_out/r2_e2e_clip_n8.json
$TR tools/e2e_clip.py --encoder projection --pairs 118000 > gpurun_out/r2_e2e_projection_n8.json 2> gpurun_out/r2_e2e_projection_n8.err; echo "e2e proj rc=$?"; cat gpurun_out/r2_e2e_projection_n8.json
