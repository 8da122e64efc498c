#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -rf -x 2>&1 | tail -8 | tee gpurun_out/r2c_gpu_tests.log
echo "== bench dblgauss"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_dblgauss.json 2> gpurun_out/r2c_bench_dblgauss.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_dblgauss.json'))
print('value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e9,'frac',d['roofline']['frac'])
PY
tail -3 gpurun_out/r2c_bench_dblgauss.err
for m in evenasph cellphone rc; do
  timeout 600 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_$m.json 2> gpurun_out/r2c_bench_$m.err
  python - $m <<'PY'
import json,sys
m=sys.argv[1]
d=json.load(open(f'gpurun_out/r2c_bench_{m}.json'))
print(m,'value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e9,'frac',d['roofline']['frac'], 'parity', d['parity_vs_reference'].get('bit_identical_p_d_op'))
PY
  tail -2 gpurun_out/r2c_bench_$m.err
done
echo "== ncu cellphone poly kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_grid_lean -c 1 -f -o gpurun_out/prof_r2c_cellphone python bench.py --model cellphone --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r2c_cell.log; tail -2 gpurun_out/ncu_r2c_cell.log
echo "== ncu dblgauss kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_grid_lean -c 1 -f -o gpurun_out/prof_r2c_dblgauss python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r2c_dbl.log; tail -2 gpurun_out/ncu_r2c_dbl.log
ls -la gpurun_out/*.ncu-rep | tail -3
