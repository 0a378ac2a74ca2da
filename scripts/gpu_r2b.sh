#!/bin/bash
# gpu tests + default bench (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -40
echo "== bench (default)"; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; wc -c gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json
