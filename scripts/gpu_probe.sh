#!/bin/bash
# Development probe: independent pytest processes (a failure in one area does not hide the others) + short benches.
#   TAG=<name> scripts/gpu_probe.sh
TAG=${TAG:-probe}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt; rm -f $S
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a $S
  SECONDS=0
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider -rA > gpurun_out/${TAG}_$name.log 2>&1
  echo "exit $? after ${SECONDS}s" | tee -a $S
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_$name.log | cut -c1-260 | tail -n 40 | tee -a $S
  grep -E "relerr|differ|agreement" gpurun_out/${TAG}_$name.log | cut -c1-220 | head -n 60 >> $S
}
bench() {  # name, args...
  local name=$1; shift
  echo "=== bench $name: $*" | tee -a $S
  timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err; echo "exit $?" | tee -a $S
  cut -c1-330 gpurun_out/${TAG}_bench_$name.json | tee -a $S
  tail -n 3 gpurun_out/${TAG}_bench_$name.err | cut -c1-300 >> $S
}
