#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for grp in "utils or march or composite or infer or distill or overflow or empty or zero_fill or compact" "grid or ffmlp or sh or inference_equals"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-24)
  echo "== parity [$grp]"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$grp" > "gpurun_out/pytest_parity_$name.log" 2>&1; echo "rc=$?"; grep "^E  .*Error\|^FAILED\|passed\|failed" "gpurun_out/pytest_parity_$name.log" | head -30
done
echo "== modules"; timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/pytest_modules.log 2>&1; echo "rc=$?"; grep "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/pytest_modules.log | head -30
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_launches.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_ffmlp_bwd|k_grid_bwd_tile|k_grid_fwd_tile|k_march_train|k_ffmlp_fwd|k_composite" -c 10 -o gpurun_out/prof_r1b -f python scripts/profile_step.py --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
