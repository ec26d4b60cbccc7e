#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   — retries while the pod answers "busy/transient" (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
