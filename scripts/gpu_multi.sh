#!/bin/bash
# N-GPU check of the ray-sharded step and the tile-sharded render (tight timeout: a hang must not burn the budget).  gpurun --gpus N -- bash scripts/gpu_multi.sh N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N rc=$?"; tail -c 600 gpurun_out/bench_n$N.json | head -c 300; echo; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$N.err | tail -6
