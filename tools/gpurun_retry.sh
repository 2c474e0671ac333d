#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> [gpurun args...] -- 'command'
# retries while gpurun answers "busy / transient" (exit 3), up to 12 times
log=$1; shift
for attempt in $(seq 1 12); do
  timeout 3500 /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "EXIT $rc" >> "$log"; exit $rc; fi
  sleep 90
done
echo "EXIT 3 (gave up)" >> "$log"
