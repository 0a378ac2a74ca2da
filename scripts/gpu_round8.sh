#!/bin/bash
# round-1j GPU pass: fixes (shared smem mark, jump resolve only for 32-lane groups), direct-sum slot assignment, Adam streaming hints
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -30
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; head -c 300 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
for v in "LNRF_ADAM_STREAM=0" "LNRF_ADAM_BLOCKS_PER_SM=4" "LNRF_ADAM_BLOCKS_PER_SM=16" "LNRF_ADAM_BLOCKS_PER_SM=41" "LNRF_GRID_PAIR=2"; do
  echo "== bench $v"; env $v timeout 600 python bench.py --no-cpu --no-render > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "rc=$?"
  python - "$v" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), {k.replace("lnrf_", ""): round(v["mean_ms"] * 1e3, 1) for k, v in d["kernels"].items()})
PY
done
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_launches.log
