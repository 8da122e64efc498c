#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -rf -x 2>&1 | tail -8 | tee gpurun_out/r2g_gpu_tests.log
for m in dblgauss evenasph cellphone rc zoom52; do
  timeout 600 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_$m.json 2> gpurun_out/r2g_bench_$m.err
  python - $m <<'PY'
import json,sys
m=sys.argv[1]
d=json.load(open(f'gpurun_out/r2g_bench_{m}.json'))
print(m,'value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e9,'frac',d['roofline']['frac'], 'parity', d['parity_vs_reference'].get('bit_identical_p_d_op'))
PY
  tail -2 gpurun_out/r2g_bench_$m.err
done
echo "== e2e breakdown"
timeout 300 python tools/e2e_breakdown.py 2>&1 | head -12
