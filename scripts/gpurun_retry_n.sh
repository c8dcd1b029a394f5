#!/bin/bash
# like gpurun_retry.sh with --gpus N: scripts/gpurun_retry_n.sh <N> <log> <timeout> <cmd...>
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
tail -40 $log | cut -c1-1800
