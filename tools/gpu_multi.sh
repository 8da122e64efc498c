#!/bin/bash
# EVERY job carries its own short timeout: a hung collective must cost 150 s x N GPUs, not the round's
# budget (round 2 lost 112 GPU-minutes to one hang under an outer limit of 843 s).
# usage (under gpurun --gpus NMAX): bash tools/gpu_multi.sh "<model:mode:N[:extra flags]> ..."
#   e.g. bash tools/gpu_multi.sh "dblgauss:replica:2:--no-e2e zoom52:shard:8:--steps=10"
set -u
JOBS=$1
mkdir -p gpurun_out
NMAX=$(nvidia-smi -L | wc -l)
echo "== mgpu_check world=$NMAX"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2
port=29520
for jm in $JOBS; do
  IFS=: read m mode N extra <<< "$jm"
  extra=${extra:-}
  port=$((port+1))
  out=gpurun_out/r2_multi_${m}_${mode}_n$N
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --model $m --mode $mode --steps 20 --warmup 5 ${extra//,/ } > $out.json 2> $out.err
  python - $out.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d['e2e']['value'] if d.get('e2e') else float('nan')
    print(sys.argv[1].split('/')[-1], 'value %.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4g'%e, d['scaling'], d['step_submission'][:14], 'imb', d.get('rank_imbalance'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" $out.err | tail -3
done
