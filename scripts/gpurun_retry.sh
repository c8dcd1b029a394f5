#!/bin/bash
# gpurun with retries while the pod answers "transient" (exit code 3) or busy: scripts/gpurun_retry.sh <log> <timeout> <cmd...>
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
tail -60 $log | cut -c1-3500
