#!/bin/bash
cd "$(dirname "$0")/.."
F="--no-render --no-cpu --no-gpu-ref --no-configs --no-large --steps 300 --warmup 20"
for cap in "16,16" "16,8" "16,4" "16,32" "8,16" "32,16" "16,16"; do
  echo "== LNRF_GRID_CAP=$cap"
  LNRF_GRID_CAP=$cap timeout 300 python bench.py $F 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
