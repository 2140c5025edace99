#!/bin/bash
# gpurun_retry.sh TIMEOUT 'command': retries while the pod answers "busy / draining" (exit 3, nothing charged)
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then cat /tmp/gpurun_last.log; exit $rc; fi
  sleep 45
done
cat /tmp/gpurun_last.log
exit 3
