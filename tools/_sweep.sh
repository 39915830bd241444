python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_kernel|rerank_kernel" -c 40 --csv --log-file gpurun_out/k2_times_r7.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/k2_times_r7.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
for r in rows[-5:]:
    print(r[ix["Kernel Name"]][:40], r[ix["Grid Size"]], float(r[ix["Metric Value"]])/1e6)
PY
python bench.py --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['achieved'], j['clocks'])"
