#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== fp64 latency / peak"
python - <<'PY'
from rayoptics_b200 import engine as E
print('dependent DFMA latency (cycles):', E.measure_fp64_latency(0), ' DFMA peak TF:', E.measure_fp64_peak(0))
PY
echo "== variants"
for v in "" _b256_c4 _b128_c6 _b128_c7 _b128_c8 _b192_c4 _b192_c5; do
  lib=$PWD/rayoptics_b200/csrc/libb200rt$v.so
  echo "-- variant '$v'"
  B200RT_LIB=$lib python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  for m in dblgauss zoom52; do
    n=512; [ $m = zoom52 ] && n=256
    B200RT_LIB=$lib python bench.py --model $m --num $n --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', round(d['value']/1e9,3), 'Grays/s', round(d['roofline']['kernel_ms'],4), 'ms')"
  done
done
NCU="ncu --set full --clock-control none --import-source on -k regex:k_trace_grid --launch-skip 3 -c 1 -f"
echo "== ncu cellphone poly kernel"
timeout 600 $NCU -o gpurun_out/prof_r2d_cellphone python bench.py --model cellphone --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r2d_cell.log; tail -1 gpurun_out/ncu_r2d_cell.log
echo "== ncu evenasph poly kernel"
timeout 600 $NCU -o gpurun_out/prof_r2d_evenasph python bench.py --model evenasph --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r2d_even.log; tail -1 gpurun_out/ncu_r2d_even.log
echo "== ncu dblgauss kernel"
timeout 600 $NCU -o gpurun_out/prof_r2d_dblgauss python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r2d_dbl.log; tail -1 gpurun_out/ncu_r2d_dbl.log
echo "== ncu general kernel (dblgauss forced through the general path)"
B200RT_NO_LEAN=1 timeout 600 $NCU -o gpurun_out/prof_r2d_general python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r2d_gen.log; tail -1 gpurun_out/ncu_r2d_gen.log
B200RT_NO_LEAN=1 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('general kernel dblgauss', round(d['value']/1e9,3), 'Grays/s', round(d['roofline']['kernel_ms'],4), 'ms')"
echo "== ncu full-ray kernel"
cat > /tmp/fullray.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bench import load_model
from rayoptics_b200 import table as T, engine as E
opm = load_model('dblgauss'); tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
grid = E.grid_for_model(opm, tab, 512)
res = E.BundleResult(grid.n_rays, tab.n_ifc, torch.device('cuda', 0), ('status', 'n_seg', 'full'))
for _ in range(5):
    E.trace_grid(tab, grid, res=res, summary=False)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_grid --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2d_fullray python /tmp/fullray.py > /dev/null 2> gpurun_out/ncu_r2d_full.log; tail -1 gpurun_out/ncu_r2d_full.log
ls -la gpurun_out/prof_r2d* 
