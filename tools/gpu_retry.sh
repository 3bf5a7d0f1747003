#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> <ngpus> '<command>'   -- retries while the pod answers busy (exit 3)
T=$1; N=$2; shift 2
for i in $(seq 1 40); do
  if [ "$N" = "1" ]; then gpurun --timeout $T -- "$@"; else gpurun --gpus $N --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
