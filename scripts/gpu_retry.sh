#!/bin/bash
# usage: scripts/gpu_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers "transient"/busy
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy\|no box\|rc=3" $log && ! grep -q "status=ok" $log; then sleep 90; continue; fi
  break
done
