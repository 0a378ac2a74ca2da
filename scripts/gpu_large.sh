#!/bin/bash
cd "$(dirname "$0")/.."
LNRF_BENCH_DEBUG=1 timeout 300 python bench.py --no-render --no-cpu --no-gpu-ref --no-configs --steps 100 --warmup 10 2>gpurun_out/large.err | tail -1 | python scripts/show_large.py
grep -v "^\s*$" gpurun_out/large.err | tail -25
