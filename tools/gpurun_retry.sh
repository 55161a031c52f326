#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy / draining): bash tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  if grep -q "status=transient\|status=busy" /tmp/gpurun_last.log; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.log
