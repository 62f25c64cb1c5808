#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers busy (exit 3) / transient
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy\|no box"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up"; exit 3
