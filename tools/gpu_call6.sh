#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -rf -x 2>&1 | tail -8 | tee gpurun_out/r2f_gpu_tests.log
for m in dblgauss evenasph cellphone; do
  timeout 600 python bench.py --model $m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_$m.json 2> gpurun_out/r2f_bench_$m.err
  python - $m <<'PY'
import json,sys
m=sys.argv[1]
d=json.load(open(f'gpurun_out/r2f_bench_{m}.json'))
print(m,'value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e9,'frac',d['roofline']['frac'], 'parity', d['parity_vs_reference'].get('bit_identical_p_d_op'))
PY
  tail -2 gpurun_out/r2f_bench_$m.err
done
echo "== e2e breakdown"
timeout 300 python tools/e2e_breakdown.py 2>&1 | head -12
echo "== compute-sanitizer (memcheck + racecheck) on a small grid with summary"
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bench import load_model
from rayoptics_b200 import table as T, engine as E, analyses as A
for name in ('dblgauss', 'cellphone'):
    opm = load_model(name); tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
    grid = E.grid_for_model(opm, tab, 40)
    r = E.trace_grid(tab, grid)
    sd = A.spot_diagram(opm, 40, table=tab)
    torch.cuda.synchronize()
    print(name, 'ok', float(r.summary[:, 0].sum()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/r2f_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/r2f_sanitizer_racecheck.log
