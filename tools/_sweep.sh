timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baselines.py -x -q 2>&1 | tail -2
LEMON_K2_BULK=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "score_pairs or property" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_kernel" -c 10 --csv --log-file gpurun_out/k2_times_r8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/k2_times_r8.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[-2:]:
    print(r[ix["Kernel Name"]][:40], r[ix["Grid Size"]], r[ix["Block Size"]], float(r[ix["Metric Value"]])/1e6)
PY
