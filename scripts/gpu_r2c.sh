#!/bin/bash
# round-2 pass: gpu tests, quick bench (headline only), ncu launch list of the training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -40
echo "== bench (headline only)"; timeout 900 python bench.py --no-configs --no-cpu --no-large ${BENCH_EXTRA} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"; tail -3 gpurun_out/bench_quick.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_launches.log
python scripts/summarize_ncu.py r2c 2 > /dev/null 2>&1; cp profiles/r2c_launches.txt gpurun_out/ 2>/dev/null; cat profiles/r2c_launches.txt | cut -c1-150
