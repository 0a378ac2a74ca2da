#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== render tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "render or distill or schedule" > gpurun_out/pytest_render.log 2>&1; echo "rc=$?"; grep "^E  \|^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_render.log | head -20
for v in 8 16 32 64; do echo "== render shares fast spr=$v"; LNRF_RENDER_SPR=$v bash scripts/gpu_render_sizes.sh 2>&1 | grep "share\|Error\|error"; done
