#!/bin/bash
# round-1k GPU pass: software-pipelined marcher windows (A/B), render share timing, full tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -30
for v in "LNRF_MARCH_PREFETCH=3" "LNRF_MARCH_PREFETCH=0" "LNRF_MARCH_PREFETCH=1"; do
  echo "== bench $v"; env $v timeout 600 python bench.py --no-cpu > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "rc=$?"
  python - "$v" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "march", round(d["kernels"]["lnrf_march_rays_train"]["mean_ms"] * 1e3, 1), "render ms/frame", round(d["render"]["ms_per_frame"], 2), d["render"].get("rounds"))
PY
done
for v in "LNRF_MARCH_PREFETCH=3" "LNRF_MARCH_PREFETCH=1"; do echo "== render shares $v"; env $v bash scripts/gpu_render_sizes.sh 2>&1 | grep "share"; done
