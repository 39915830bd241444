#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "second_pass or split_precision" > gpurun_out/r2_pytest_gpu19a.log 2>&1; echo "subset rc=$?"; grep -E "uncertified|eps first|passed|failed" gpurun_out/r2_pytest_gpu19a.log | tail -8
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu19.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/r2_smoke_ncu.log 2>&1; echo "smoke ncu rc=$?"; grep -c "lemon::" gpurun_out/r2_smoke_launches.csv; grep -v "lemon::" gpurun_out/r2_smoke_launches.csv | grep -c "at::\|native"
