#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [gpurun args...] -- retries while the pod answers "transient" (busy slots), nothing is charged for those
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 100; continue; fi
  break
done
