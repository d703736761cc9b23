#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] --timeout S -- 'command'   -- retries while the pod answers "busy" (nothing charged)
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 75; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 30 busy answers"; exit 3
