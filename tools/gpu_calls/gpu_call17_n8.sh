#!/bin/bash
# final multi-GPU check of the driver's SCALE command lines (C3 at N = 4 and 8, both arms) with the final code
mkdir -p gpurun_out
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N"
  $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n${N}_final.json 2> gpurun_out/r2_bench_c3_n${N}_final.err; echo "c3 n$N rc=$?"; tail -2 gpurun_out/r2_bench_c3_n${N}_final.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
$TR bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err; echo "ref n8 rc=$?"
$TR tools/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_n8_final.log 2>&1; grep MULTI_GPU gpurun_out/r2_multi_gpu_check_n8_final.log
python - <<'P'
import json
for w in ("c3_n8_final", "c3_n4_final"):
    try:
        b = json.loads(open(f"gpurun_out/r2_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["parity"]["wrong"], b["roofline"]["frac"], b["roofline"]["k1_share_of_step"], b["run_info"], b["clocks"]["sm_mhz"])
    except Exception as e:
        print(w, "failed", e)
r = json.loads(open("gpurun_out/r2_bench_ref_n8.json").read().strip().splitlines()[-1])
print("ref n8", r["value"], r["cpu_baseline"]["cores"])
P
