#!/bin/bash
# round-1i GPU pass: fused composite+loss tail (row f-5), x-pair merged encoder gathers/reductions, warp-uniform corner indices.
# correctness first, then bench (A/B on LNRF_GRID_PAIR), reference arm, ncu launch list + full capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== new tests first"; timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x -k "composite_loss or fused_loss or grid or encode or train_step or jump or paired or march" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; grep "^E  \|^FAILED\|passed\|failed" gpurun_out/pytest_new.log | head -30
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -30
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; head -c 400 gpurun_out/bench.json; echo; tail -5 gpurun_out/bench.err
for pm in 0 1 2; do
  echo "== bench LNRF_GRID_PAIR=$pm"; LNRF_GRID_PAIR=$pm timeout 600 python bench.py --no-cpu --no-render > gpurun_out/bench_pair$pm.json 2> gpurun_out/bench_pair$pm.err; echo "rc=$?"; head -c 200 gpurun_out/bench_pair$pm.json; echo
done
echo "== bench LNRF_MARCH_JUMP=0"; LNRF_MARCH_JUMP=0 timeout 600 python bench.py --no-cpu > gpurun_out/bench_jump0.json 2> gpurun_out/bench_jump0.err; echo "rc=$?"; head -c 200 gpurun_out/bench_jump0.json; echo
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; head -c 300 gpurun_out/bench_reference.json; echo
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_launches.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_mlp_bwd2|k_grid_bwd_tile|k_grid_fwd_tile|k_march_train|k_nerf_fwd|k_composite|k_adam_step|k_grad_nonfinite" -c 12 -o gpurun_out/prof_r1i -f python scripts/profile_step.py --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
