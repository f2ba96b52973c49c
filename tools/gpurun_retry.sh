#!/bin/bash
# Developer script: retry a gpurun call while the pod answers "transient" (busy, nothing charged).
# usage: tools/gpurun_retry.sh <logfile> <timeout-seconds> '<command>'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|refused" "$log"; then exit 0; fi
  sleep 150
done
exit 1
