python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/tc_debug.py 2 4000 20000 512 1 2>&1 | tail -6
python tools/tc_debug.py 2 118000 118000 512 1 --time 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/bench_ladder.json 2> gpurun_out/bench_ladder.err; tail -2 gpurun_out/bench_ladder.err; cat gpurun_out/bench_ladder.json
