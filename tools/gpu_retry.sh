#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3) or another call of this repo is still in flight
# (exit code 2, retried a few times only: 2 also means "budget spent"): tools/gpu_retry.sh [--gpus N] <timeout_s> '<command>'
GP=()
if [ "$1" = "--gpus" ]; then GP=(--gpus "$2"); shift 2; fi
T=$1; shift
refused=0
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "${GP[@]}" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -eq 2 ] && [ $refused -lt 6 ]; then refused=$((refused+1)); sleep 45; continue; fi
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
