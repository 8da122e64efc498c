#!/bin/bash
# usage (on the GPU box): tools/variant_bench.sh <suffix> [<suffix> ...]
# runs the oracle smoke check and the kernel-only bench with each libb200rt<suffix>.so
for v in "$@"; do
  lib=$PWD/rayoptics_b200/csrc/libb200rt$v.so
  echo "== variant '$v'"
  B200RT_LIB=$lib python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  for m in dblgauss zoom52; do
    n=512; [ $m = zoom52 ] && n=256
    B200RT_LIB=$lib python bench.py --model $m --num $n --steps 100 --warmup 5 --no-e2e --no-cpu-baseline \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', round(d['value']/1e9,3), 'Grays/s', round(d['roofline']['kernel_ms'],4), 'ms')"
  done
done
