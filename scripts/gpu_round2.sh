#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== diag ffmlp"; timeout 600 python scripts/diag_ffmlp.py 2>&1 | tail -12
echo "== dump ours"; timeout 600 python scripts/dump_ours.py > gpurun_out/dump_ours.log 2>&1; tail -13 gpurun_out/dump_ours.log
for grp in "utils or march or composite or infer or distill or overflow or empty or zero_fill or compact" "grid or ffmlp or sh or inference_equals"; do
  name=$(echo "$grp" | tr ' ' '_' | cut -c1-24)
  echo "== parity [$grp]"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$grp" > "gpurun_out/pytest_parity_$name.log" 2>&1; echo "rc=$?"; grep "^E  .*Error\|^FAILED\|passed\|failed" "gpurun_out/pytest_parity_$name.log" | head -30
done
echo "== modules"; timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/pytest_modules.log 2>&1; echo "rc=$?"; grep "^E  .*Error\|^FAILED\|passed\|failed" gpurun_out/pytest_modules.log | head -30
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
