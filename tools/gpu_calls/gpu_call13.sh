#!/bin/bash
export LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so
python tools/k1_variants.py 151552 370000 512 PACE=24 PACE=24,DEBUG=2 PACE=24,DEBUG=1 PACE=24,KRES=7 PACE=24,KRES=6 PACE=24,CERT=34 PACE=24,BOOT=0 PACE=24,BOOT=16 2>&1 | tee gpurun_out/r2_k1_floor.log
python tools/k1_variants.py 75776 1000000 768 PACE=24 PACE=24,DEBUG=2 PACE=24,DEBUG=1 PACE=24,KRES=9 PACE=24,KRES=11 PACE=24,KRES=12 2>&1 | tee -a gpurun_out/r2_k1_floor.log
