#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   (retries while the pod answers "busy")
log=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 90; continue; fi
  break
done
tail -25 $log
