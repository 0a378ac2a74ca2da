#!/bin/bash
# lookahead placements of the next batch's march: none / beside the exchange+Adam / beside the hash-grid backward (two stream priorities)
F="--no-render --no-cpu --no-gpu-ref --no-configs --no-large --steps 300 --warmup 20"
for mode in "0 0 x" "1 0 enc_bwd" "1 -1 enc_bwd" "1 0 start" "1 -1 start" "1 0 nerf_bwd" "1 -1 nerf_bwd" "1 0 adam" "1 -1 adam"; do
  set -- $mode
  echo "== LNRF_LOOKAHEAD=$1 prio=$2 at=$3"
  LNRF_LOOKAHEAD=$1 LNRF_LOOKAHEAD_PRIO=$2 LNRF_LOOKAHEAD_AT=$3 timeout 300 python bench.py $F 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('e2e',{}).get('value'), d['config'].get('step_mode'))"
done
