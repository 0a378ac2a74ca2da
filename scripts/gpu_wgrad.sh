#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
F="--no-render --no-cpu --no-gpu-ref --no-configs --no-large --steps 300 --warmup 20"
for w in 0 1 0 1; do
  echo "== LNRF_WGRAD_SIDE=$w"
  LNRF_WGRAD_SIDE=$w timeout 300 python bench.py $F 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
