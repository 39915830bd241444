#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu4.log
tail -4 gpurun_out/r2_pytest_gpu4.log
export LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so
python tools/k1_variants.py 118000 118000 512 PACE=0 PACE=24 PACE=0 PACE=24 2>&1 | tee gpurun_out/r2_k1_pace.log
python tools/k1_variants.py 370000 370000 512 PACE=0 PACE=24 PACE=48 PACE=12 PACE=0 PACE=24 2>&1 | tee -a gpurun_out/r2_k1_pace.log
K1_REPS=2 python tools/k1_variants.py 412500 3300000 768 PACE=0 PACE=24 PACE=64 PACE=0 PACE=24 2>&1 | tee -a gpurun_out/r2_k1_pace.log
unset LEMON_B200_LIB
python tools/uncert_rate.py 3300000 768 4096 0.95 0.98 0.99 > gpurun_out/r2_uncert_rate2.jsonl 2> gpurun_out/r2_uncert_rate2.err; echo "uncert rc=$?"; cat gpurun_out/r2_uncert_rate2.jsonl; tail -3 gpurun_out/r2_uncert_rate2.err
