#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "8,8" "6,6" "4,4" "12,12" "16,16" "3,3"; do
  LNRF_GRID_CAP=$v timeout 600 python bench.py --no-cpu --no-render > gpurun_out/bench_cap.json 2> gpurun_out/bench_cap.err; echo "cap=$v rc=$?"
  python - "$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_cap.json"))
k = d["kernels"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "enc fwd", round(k["lnrf_grid_encode_forward_world"]["mean_ms"] * 1e3, 1), "enc bwd", round(k["lnrf_grid_encode_backward_world"]["mean_ms"] * 1e3, 1), "samples", d["config"]["samples_per_step_padded"])
PY
done
