#!/bin/bash
# final ncu captures of the shipped K1 at the dominant launch shapes of the C3 (N=1) and C4 (N=8) bench lines
mkdir -p gpurun_out
K1A="python tools/k1_launch.py 359936 370000 512"
$K1A > gpurun_out/r2_k1c3main_final_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c3main_final $K1A > gpurun_out/r2_k1c3main_final_ncu.log 2>&1
echo "k1 c3 main ncu rc=$?"; tail -1 gpurun_out/r2_k1c3main_final_plain.log
K1B="python tools/k1_launch.py 412500 3300000 768"
$K1B > gpurun_out/r2_k1c4rank_final_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 1 -c 1 -o gpurun_out/r2_k1_c4rank_final $K1B > gpurun_out/r2_k1c4rank_final_ncu.log 2>&1
echo "k1 c4 rank ncu rc=$?"; tail -1 gpurun_out/r2_k1c4rank_final_plain.log
