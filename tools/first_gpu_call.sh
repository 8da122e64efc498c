#!/bin/bash
# First GPU call of a round (run under gpurun, one B200): everything that was only validated on
# CPU, then the headline, then the profile.  Outputs under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
echo "== OPD grid timing (WAVE kernels)"
timeout 120 python - <<'PY' 2>&1 | tail -3
import sys, time, torch
sys.path.insert(0, 'tests')
from conftest import load_model
from rayoptics_b200 import analyses as A
opm = load_model('dblgauss')
A.RayGrid(opm, f=1, num_rays=512)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    A.RayGrid(opm, f=1, num_rays=512)
torch.cuda.synchronize()
print('RayGrid 512x512 (one tile, OPD epilogue):', (time.perf_counter() - t0)/5*1e3, 'ms per call')
PY
echo "== launch list + full capture of the hot kernel"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trace_grid_lean -c 1 \
    -o gpurun_out/prof_hot python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
