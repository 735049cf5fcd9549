#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 45; continue; fi
  break
done
