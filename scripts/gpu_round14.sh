#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== march tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "march or matches or render or distill or schedule or jump" > gpurun_out/pytest_march.log 2>&1; echo "rc=$?"; grep "^E  \|^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_march.log | head -20
for v in 1; do echo "== scenes LNRF_MARCH_FF=$v"; LNRF_MARCH_FF=$v timeout 800 python scripts/bench_scenes.py > gpurun_out/scenes_ff$v.json 2> gpurun_out/scenes_ff$v.err; echo "rc=$?"; tail -2 gpurun_out/scenes_ff$v.err; done
