#!/bin/bash
# usage: gpurun_retry.sh [--gpus N] TIMEOUT "command"   -- retries while the pod is busy
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up: pod busy"
