#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing is charged for those): tools/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for k in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_try.log 2>&1
  if grep -q "status=transient" /tmp/gpurun_try.log; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_try.log
