#!/bin/bash
# usage: tools/gpurun_retry.sh LOG TIMEOUT [--gpus N] -- cmd     (retries while the pod answers busy/transient)
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
