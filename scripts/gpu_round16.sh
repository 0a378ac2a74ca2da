#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fused/mlp tests"; timeout 900 python -m pytest tests -m gpu -q -k "fused or ffmlp or train_step or graphed or whole or style" > gpurun_out/pytest_f.log 2>&1; echo "rc=$?"; grep "^E  \|^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_f.log | head -30
echo "== bench"; timeout 900 python bench.py --no-cpu --no-render > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches"] / d["steps"], {k.replace("lnrf_", ""): round(v["mean_ms"] * 1e3, 1) for k, v in d["kernels"].items()})
PY
