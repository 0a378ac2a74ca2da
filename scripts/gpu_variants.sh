#!/bin/bash
# quick headline bench under several env variants (A/B switches), one line each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-configs --no-cpu --no-large --no-gpu-ref --no-render > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json")); print("$tag", "ms/step %.4f" % d["ms_per_step"], "e2e %.3fM" % (d["e2e"]["value"]/1e6), "occ %.4f" % d["value_with_occupancy_update"]["ms_per_step"], d["config"]["step_mode"][:60])
except Exception as e: print("$tag", "FAILED", e)
PY
}
run base A=1
run lookahead LNRF_LOOKAHEAD=1
run lookahead_adam2 LNRF_LOOKAHEAD=1 LNRF_ADAM_BLOCKS_PER_SM=2
run lookahead_adam1 LNRF_LOOKAHEAD=1 LNRF_ADAM_BLOCKS_PER_SM=1
run adam2 LNRF_ADAM_BLOCKS_PER_SM=2
run three_launch LNRF_ADAM_ONE_LAUNCH=0
run no_count LNRF_DEVICE_COUNT=0
