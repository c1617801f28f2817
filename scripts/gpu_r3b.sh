#!/bin/bash
# final validation: the driver's two GPU commands + the default bench
source scripts/gpu_probe.sh
run full 2400 tests/ -x
echo "=== smoke" | tee -a $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $S
bench default
bench reference --impl reference --steps 3 --warmup 1
