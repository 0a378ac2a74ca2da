#!/bin/bash
# round-2 validation pass: all -m gpu tests (no -x: every failure is wanted), bench.py default + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -40
echo "== bench (default)"; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; wc -c gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; head -c 300 gpurun_out/bench_reference.json; echo
