#!/bin/bash
# evidence pass: bench, ncu launch list of the training step, ncu --set full of the hot kernels (scripts/summarize_ncu.py turns them into profiles/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; head -c 200 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_launches.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_mlp_bwd2|k_grid_bwd_tile|k_grid_fwd_tile|k_march_train|k_nerf_fwd|k_composite|k_adam_step|k_grad_nonfinite" -c 12 -o gpurun_out/prof_r1k -f python scripts/profile_step.py --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_r1k.ncu-rep
