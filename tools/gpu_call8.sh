#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 | tee gpurun_out/r2h_gpu_tests.log
echo "== gpu tests, dynamic scheduling"; B200RT_DYNAMIC=1 timeout 900 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 | tee gpurun_out/r2h_gpu_tests_dynamic.log
for dyn in 0 1; do
for m in dblgauss rc cellphone zoom52; do
  n=""; [ $m = zoom52 ] && n="--num 256"
  if [ $dyn = 1 ]; then export B200RT_DYNAMIC=1; else unset B200RT_DYNAMIC; fi
  timeout 600 python bench.py --model $m $n --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dynamic=$dyn', '$m', round(d['value']/1e9,3), 'Grays/s', round(d['roofline']['kernel_ms'],4), 'ms', 'frac', round(d['roofline']['frac'],4))"
done
done
