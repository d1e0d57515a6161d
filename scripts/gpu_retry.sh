#!/bin/bash
# usage: scripts/gpu_retry.sh <logfile> <timeout_s> [--gpus N] -- '<command>'
# Retries gpurun while the pod answers "busy" (exit code 3), every 3 minutes, up to 15 tries.
log=$1; shift
tmo=$1; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[gpu_retry] rc=$rc try=$i" >> "$log"; exit $rc; fi
  sleep 180
done
echo "[gpu_retry] gave up" >> "$log"; exit 3
