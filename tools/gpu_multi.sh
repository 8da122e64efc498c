#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh N "<model:mode> ..."
#   e.g. bash tools/gpu_multi.sh 4 "dblgauss:replica dblgauss:shard evenasph:shard cellphone:shard"
set -u
N=$1; shift
JOBS=${1:-"dblgauss:replica dblgauss:shard"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== mgpu_check world=$N"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -3
port=29520
for jm in $JOBS; do
  m=${jm%%:*}; mode=${jm##*:}
  port=$((port+1))
  out=gpurun_out/r2_multi_${m}_${mode}_n$N
  timeout 900 $TR --master-port $port bench.py --gpus $N --model $m --mode $mode --steps 20 --warmup 5 > $out.json 2> $out.err
  python - $out.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'scaling',d['scaling'], 'imb', d.get('rank_imbalance'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  grep -v "^W\|^\*\*\*\|^$" $out.err | tail -3
done
