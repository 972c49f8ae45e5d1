#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "busy" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pod busy"; exit 3
