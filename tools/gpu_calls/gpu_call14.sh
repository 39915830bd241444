#!/bin/bash
export LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so
python tools/k1_variants.py 118000 118000 512 BOOT=8 BOOT=16 BOOT=8 BOOT=16 2>&1 | tee gpurun_out/r2_k1_boot.log
python tools/k1_variants.py 370000 370000 512 BOOT=8 BOOT=16 BOOT=8 BOOT=16 2>&1 | tee -a gpurun_out/r2_k1_boot.log
python tools/k1_variants.py 75776 1000000 768 BOOT=8 BOOT=16 BOOT=8 BOOT=16 2>&1 | tee -a gpurun_out/r2_k1_boot.log
python tools/k1_variants.py 46250 370000 512 BOOT=8 BOOT=16 2>&1 | tee -a gpurun_out/r2_k1_boot.log
