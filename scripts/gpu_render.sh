#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "schedule or economies or edge_cases or refstack" 2>&1 | tail -1
for c in 1 0; do echo "== LNRF_FIXUP_CAPS=$c"; LNRF_FIXUP_CAPS=$c python scripts/diag_fixup.py 2>&1 | grep "rays: fast" | cut -c1-150; done
