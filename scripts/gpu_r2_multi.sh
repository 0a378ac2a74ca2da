#!/bin/bash
# N-GPU pass: the 2-GPU sharded-vs-single test, then bench.py under torchrun on all visible GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus: $N"
echo "== 2-GPU sharded-vs-single test"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_multi.log
echo "== bench --gpus $N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_EXTRA} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; wc -c gpurun_out/bench_n$N.json; tail -4 gpurun_out/bench_n$N.err
