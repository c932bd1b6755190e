#!/bin/bash
# gpurun with retry on "transient" (pod busy, nothing charged): tools/gpu.sh <timeout_s> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -${TAILN:-40}
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  break
done
