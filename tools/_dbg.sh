#!/bin/bash
F="--steps 30 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e --no-optimizer --no-extra-configs"
for r in 32768 1; do
  for c in code2-pna; do
  echo "=== rows>$r $c"; GT_WGRAD_SIDE_MAX_ROWS=$r timeout 100 python bench.py $F --config $c 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"
  done
done
