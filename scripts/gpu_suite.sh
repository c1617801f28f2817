#!/bin/bash
# Full GPU validation in the driver's order: pytest -m gpu (one process, -x), smoke, default bench.
#   TAG=<name> scripts/gpu_suite.sh [extra bench args]
TAG=${TAG:-suite}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $S
echo "=== full GPU suite, one process" | tee -a $S
SECONDS=0
timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -rA --durations=15 > gpurun_out/${TAG}_full.log 2>&1; echo "exit $? after ${SECONDS}s" | tee -a $S
grep -E "passed|failed|error" gpurun_out/${TAG}_full.log | tail -n 3 | cut -c1-300 | tee -a $S
grep -E "relerr|mismatch|bit-exact|ids" gpurun_out/${TAG}_full.log | cut -c1-220 | head -n 80 >> $S
echo "=== smoke" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $S
echo "=== bench (default)" | tee -a $S
timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "exit $?" | tee -a $S; cut -c1-400 gpurun_out/${TAG}_bench.json | tee -a $S
