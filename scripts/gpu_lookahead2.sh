#!/bin/bash
F="--no-render --no-cpu --no-gpu-ref --no-configs --no-large --steps 300 --warmup 20"
run() { echo "== lookahead=$1 prio=$2 at=$3 adam_blocks_per_sm=$4"
  LNRF_LOOKAHEAD=$1 LNRF_LOOKAHEAD_PRIO=$2 LNRF_LOOKAHEAD_AT=$3 LNRF_ADAM_BLOCKS_PER_SM=$4 timeout 300 python bench.py $F 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('e2e',{}).get('value'))"; }
for b in 1 2 3; do run 0 0 x $b; done
for b in 1 2 3; do for p in 0 -1; do run 1 $p adam $b; done; done
for b in 2 3; do run 1 0 nerf_bwd $b; done
