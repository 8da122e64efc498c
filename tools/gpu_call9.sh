#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests, dynamic scheduling"; timeout 900 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8 | tee gpurun_out/r2j_gpu_tests_dynamic.log
echo "== gpu tests, static"; B200RT_STATIC=1 timeout 900 python -m pytest tests -m gpu -q -rf -x 2>&1 | tail -4 | tee gpurun_out/r2j_gpu_tests_static.log
for dyn in 0 1; do
for m in dblgauss rc cellphone evenasph zoom52; do
  n=""; [ $m = zoom52 ] && n="--num 256"
  if [ $dyn = 0 ]; then export B200RT_STATIC=1; else unset B200RT_STATIC; fi
  timeout 600 python bench.py --model $m $n --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dynamic=$dyn', '$m', round(d['value']/1e9,3), 'Grays/s', round(d['roofline']['kernel_ms'],4), 'ms', 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,3))"
done
done
