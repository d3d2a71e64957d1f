#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
#   scripts/gpu.sh <timeout_s> <log file> '<command>'
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
