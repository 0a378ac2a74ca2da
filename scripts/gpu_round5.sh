#!/bin/bash
# round-1c GPU pass: fused rows f-1 / f-4.  tests -> bench -> reference-extension step timing -> ncu launch list -> ncu full
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?"; grep "^E  \|^FAILED\|passed\|failed" gpurun_out/pytest_fused.log | head -40
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|passed\|failed" gpurun_out/pytest_gpu.log | head -30
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; head -c 700 gpurun_out/bench.json; echo; tail -5 gpurun_out/bench.err
echo "== gpu reference step"; timeout 600 python tests/bench_gpu_reference.py > gpurun_out/gpu_reference.json 2> gpurun_out/gpu_reference.err; echo "rc=$?"; head -c 1500 gpurun_out/gpu_reference.json; echo; tail -5 gpurun_out/gpu_reference.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_launches.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_ffmlp_bwd|k_grid_bwd_tile|k_grid_fwd_tile|k_march_train|k_nerf_fwd|k_composite|k_adam_step|k_grad_nonfinite" -c 12 -o gpurun_out/prof_r1c -f python scripts/profile_step.py --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
