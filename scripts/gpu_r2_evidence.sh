#!/bin/bash
# round-2 evidence pass (1 GPU): gpu tests, bench (default) + reference arm, ncu launch list, ncu --set full of the hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r2}
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep "^FAILED\|^ERROR\|passed\|failed" gpurun_out/pytest_gpu.log | head -40
echo "== bench (default)"; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; wc -c gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"
echo "== ncu launches (train)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_launches.log
echo "== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_nerf_bwd|k_nerf_fwd|k_grid_bwd_tile|k_grid_fwd_tile|k_march_train|k_composite|k_adam_step|k_grad_nonfinite|k_wgrad_reduce" -c 14 -o gpurun_out/prof_${TAG} -f python scripts/profile_step.py --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_${TAG}.ncu-rep
echo "== ncu launches (render, auto schedule)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/render_launches.csv python scripts/profile_step.py --steps 0 --render-rays 640000 --render-schedule auto > gpurun_out/ncu_render.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_render.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
