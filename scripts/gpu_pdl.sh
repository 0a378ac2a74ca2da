#!/bin/bash
cd "$(dirname "$0")/.."
F="--no-render --no-cpu --no-gpu-ref --no-configs --no-large --steps 300 --warmup 20"
for la in 0 1; do for pdl in 0 1; do
  echo "== LNRF_LOOKAHEAD=$la LNRF_PDL=$pdl"
  LNRF_LOOKAHEAD=$la LNRF_PDL=$pdl timeout 300 python bench.py $F 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['eager']['ms_per_step'] if 'eager' in d else '')"
done; done
