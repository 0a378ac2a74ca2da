#!/bin/bash
F="--no-render --no-cpu --no-gpu-ref --no-configs --no-large --steps 300 --warmup 20"
run() { echo "== lookahead=$1 fused_tail=$2 early_amp=$3"
  LNRF_LOOKAHEAD=$1 LNRF_LOOKAHEAD_AT=nerf_bwd LNRF_FUSED_TAIL=$2 LNRF_EARLY_AMP_UPDATE=$3 timeout 300 python bench.py $F 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('e2e',{}).get('value'))"; }
run 0 0 0; run 0 1 1; run 1 1 0; run 1 1 1
