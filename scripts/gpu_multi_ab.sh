#!/bin/bash
# pipelined (two gradient buffers) vs one-buffer exchange under torchrun on all visible GPUs
cd "$(dirname "$0")/.."
N=$(python -c "import torch; print(torch.cuda.device_count())")
for pe in 0 1 0 1; do
  echo "== LNRF_PIPELINED_EXCHANGE=$pe (N=$N)"
  LNRF_PIPELINED_EXCHANGE=$pe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$pe bench.py --gpus $N --steps 200 --warmup 20 --no-render --no-cpu --no-gpu-ref --no-configs --no-large 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config'].get('replicas_in_sync'))"
done
