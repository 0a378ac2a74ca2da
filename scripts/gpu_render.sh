#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "schedule or device_driven or refstack or render" 2>&1 | tail -2
for c in 1 0; do echo "== clip $c"; LNRF_RENDER_CLIP=$c python scripts/diag_render_shards.py 2>&1 | grep "auto\|reference" | head -12; done
