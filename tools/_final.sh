set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:knn_tc -c 1 -o gpurun_out/k1_c2_final -f python tools/tc_debug.py 2 118000 118000 512 1 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/bench_c2_final.json; cat gpurun_out/bench_ref_final.json
