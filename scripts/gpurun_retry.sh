#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3) or another call is still running (exit 2).
#   [GPUS=2] scripts/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
