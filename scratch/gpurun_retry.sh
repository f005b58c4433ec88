#!/bin/bash
# usage: scratch/gpurun_retry.sh <timeout-seconds> <log> [--gpus N] -- <command string>
# retries gpurun while the pod answers "busy" (exit 3), up to ~40 min
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
