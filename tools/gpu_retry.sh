#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers "transient"/busy
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|exit code 3\|rc=3"; then sleep 120; continue; fi
  break
done
