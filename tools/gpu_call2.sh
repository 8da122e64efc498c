#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -rf -x 2>&1 | tail -25 | tee gpurun_out/r2b_gpu_tests.log
echo "== bench dblgauss"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_dblgauss.json 2> gpurun_out/r2b_bench_dblgauss.err; tail -c 4000 gpurun_out/r2b_bench_dblgauss.json; tail -5 gpurun_out/r2b_bench_dblgauss.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; tail -c 2500 gpurun_out/r2b_bench_ref.json; tail -5 gpurun_out/r2b_bench_ref.err
for m in rc evenasph cellphone zoom52; do
  echo "== bench $m"; timeout 600 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_$m.json 2> gpurun_out/r2b_bench_$m.err; head -c 600 gpurun_out/r2b_bench_$m.json; echo; tail -3 gpurun_out/r2b_bench_$m.err
done
