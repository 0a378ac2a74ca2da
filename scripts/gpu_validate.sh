#!/bin/bash
# GPU-box validation pass: all -m gpu tests, smoke(), bench.py (default) and the reference arm.  gpurun -- bash scripts/gpu_validate.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -30
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; wc -l gpurun_out/bench.json; head -c 300 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; head -c 200 gpurun_out/bench_reference.json; echo
