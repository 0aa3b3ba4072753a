#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout> '<command>'   — retries while the pod answers busy (exit 3)
N=$1; T=$2; shift; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
