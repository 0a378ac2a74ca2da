#!/bin/bash
# quick GPU pass: MLP/fused correctness, bench, ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fused + parity(ffmlp) tests"; timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?"; grep "^E  \|^FAILED\|passed\|failed" gpurun_out/pytest_fused.log | head -40
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; head -c 300 gpurun_out/bench.json; echo; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 ${PROFILE_ARGS} > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_launches.log
