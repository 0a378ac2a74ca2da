#!/bin/bash
cd "$(dirname "$0")/.."
for g in 8 4 16 1; do echo "== composite group $g"
LNRF_COMPOSITE_INFER_G=$g timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "schedule or device_driven or refstack" 2>&1 | tail -1
LNRF_COMPOSITE_INFER_G=$g python scripts/diag_render_shards.py 2>&1 | grep "auto" | head -6 | cut -c1-70; done
