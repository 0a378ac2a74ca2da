#!/bin/bash
# One gpurun call: golden vectors from the reference extensions, GPU parity tests (grouped per process so that a
# faulting kernel cannot poison the other groups), a short bench, and the ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== golden" ; timeout 900 python oracle/gen_golden.py > gpurun_out/gen_golden.log 2>&1 ; echo "rc=$?"; tail -15 gpurun_out/gen_golden.log
mkdir -p tests/golden; cp gpurun_out/golden/ref_*.npz tests/golden/ 2>/dev/null
for grp in "utils or march or composite or infer or distill or overflow or empty or zero_fill or compact" "grid" "ffmlp_sigma" "ffmlp_color or sh" "inference_equals"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-24)
  echo "== parity [$grp]"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$grp" > "gpurun_out/pytest_parity_$name.log" 2>&1; echo "rc=$?"; tail -25 "gpurun_out/pytest_parity_$name.log"
done
echo "== modules"; timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/pytest_modules.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_modules.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_launches.log
