#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): usage gpurun_retry.sh <timeout> <command...>
t=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
