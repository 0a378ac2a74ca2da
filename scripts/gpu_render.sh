#!/bin/bash
cd "$(dirname "$0")/.."
for g in 4 8; do echo "== compact march lanes per ray $g"
LNRF_COMPACT_G=$g timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "schedule or economies or edge_cases" 2>&1 | tail -1
LNRF_COMPACT_G=$g python scripts/diag_render_shards.py 2>&1 | grep "auto" | grep "world 1 \|world 8 " | cut -c1-75; done
