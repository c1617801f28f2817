#!/bin/bash
# validation after the GroupNorm-statistics fusion: full suite, smoke, default bench, ART-V + train benches
source scripts/gpu_probe.sh
run full 2400 tests/ -x
echo "=== smoke" | tee -a $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $S
bench default
bench artv --no-cpu-baseline --workload artv --steps 2 --warmup 1
bench train --no-cpu-baseline --workload train --precision tf32 --steps 3 --warmup 3
