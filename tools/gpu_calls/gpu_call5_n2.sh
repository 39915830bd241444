#!/bin/bash
# 2 GPUs: bitwise check across world sizes, C3 bench at N=2, reference arm under torchrun, embedding hand-off
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n2.log 2>&1; echo "check rc=$?"; grep MULTI_GPU gpurun_out/r2_multi_gpu_check_n2.log
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n2.json 2> gpurun_out/r2_bench_c3_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2_bench_c3_n2.err
$TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<'P'
import json
b = json.loads(open("gpurun_out/r2_bench_c3_n2.json").read().strip().splitlines()[-1])
print("n2", b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["clocks"])
r = json.loads(open("gpurun_out/r2_bench_ref_n2.json").read().strip().splitlines()[-1])
print("ref n2", r["value"], r["cpu_baseline"]["cores"])
P
$TR tools/e2e_clip.py --encoder projection --pairs 118000 > gpurun_out/r2_e2e_projection_n2.json 2> gpurun_out/r2_e2e_projection_n2.err; echo "e2e proj rc=$?"; cat gpurun_out/r2_e2e_projection_n2.json; tail -2 gpurun_out/r2_e2e_projection_n2.err
$TR tools/e2e_clip.py --encoder clip --pairs 24000 > gpurun_out/r2_e2e_clip_n2.json 2> gpurun_out/r2_e2e_clip_n2.err; echo "e2e clip rc=$?"; cat gpurun_out/r2_e2e_clip_n2.json; tail -2 gpurun_out/r2_e2e_clip_n2.err
