#!/bin/bash
# round-2 GPU call 1: tests, default bench, reference arm, K1 variant timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench_c3.json; tail -5 gpurun_out/r2_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r2_bench_ref.json
export LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so
for shape in "118000 118000 512" "75776 370000 512" "37888 400000 768"; do
  python tools/k1_variants.py $shape 2>&1 | tee -a gpurun_out/r2_k1_variants.log
done
