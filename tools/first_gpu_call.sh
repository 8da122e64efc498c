#!/bin/bash
# First GPU call of a round (one B200): validation, headline, profile -- every step under its OWN
# short timeout, ncu reports summarised on the box (gpurun_out is limited to 64 MiB: keep .md + raw
# csv, at most two .ncu-rep).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 300 python -m pytest tests -m gpu -q -rf 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.json
echo "== reference arm"; timeout 240 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
echo "== launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
echo "== full capture of the hot kernel (summarised here, report kept)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace_grid --launch-skip 3 -c 1 -f \
    -o gpurun_out/prof_hot python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_hot.ncu-rep gpurun_out/ncu_hot.md > /dev/null 2>&1; head -20 gpurun_out/ncu_hot.md
ls -la gpurun_out | tail -8
