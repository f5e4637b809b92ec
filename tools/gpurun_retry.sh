#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>' [extra gpurun flags]; retries while the pod answers "transient" (nothing charged)
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" --timeout $T -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$OUT" | tail -60
  exit 0
done
echo "gave up after 40 transient answers"
