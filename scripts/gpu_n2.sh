#!/bin/bash
# 2-GPU check of the ray-sharded step (tight timeouts: a hang must not burn the budget)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p2p in 1; do
  LNRF_P2P=$p2p timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$p2p bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_n2_p2p$p2p.json 2> gpurun_out/bench_n2_p2p$p2p.err; echo "p2p=$p2p rc=$?"; tail -c 900 gpurun_out/bench_n2_p2p$p2p.json | head -c 400; echo; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2_p2p$p2p.err | tail -4
done
