#!/bin/bash
# ncu captures, summarised ON THE BOX (gpurun_out is limited to 64 MiB): keep md + raw csv, drop most reports
set -u
mkdir -p gpurun_out
cap() {   # name, env, command...
  name=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_grid --launch-skip 3 -c 1 -f \
      -o /tmp/prof_$name "$@" > /dev/null 2> gpurun_out/ncu_$name.log
  tail -1 gpurun_out/ncu_$name.log
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep gpurun_out/r02e_ncu_$name.md > /dev/null 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/r02e_raw_$name.csv 2>/dev/null
  head -12 gpurun_out/r02e_ncu_$name.md | tail -6
}
cap dblgauss python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
cp /tmp/prof_dblgauss.ncu-rep gpurun_out/prof_r2e_dblgauss.ncu-rep
cap cellphone python bench.py --model cellphone --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
cp /tmp/prof_cellphone.ncu-rep gpurun_out/prof_r2e_cellphone.ncu-rep
cap evenasph python bench.py --model evenasph --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
B200RT_NO_LEAN=1 cap general python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
cat > /tmp/fullray.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bench import load_model
from rayoptics_b200 import table as T, engine as E
opm = load_model('dblgauss'); tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
grid = E.grid_for_model(opm, tab, 512)
res = E.BundleResult(grid.n_rays, tab.n_ifc, torch.device('cuda', 0), ('status', 'n_seg', 'full'))
for _ in range(6):
    E.trace_grid(tab, grid, res=res, summary=False)
torch.cuda.synchronize()
PY
cap fullray python /tmp/fullray.py
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02e_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -5 gpurun_out/r02e_launches_bench.csv | cut -c1-200
echo "== e2e breakdown"
timeout 300 python tools/e2e_breakdown.py 2>&1 | head -45
ls -la gpurun_out | tail -20
