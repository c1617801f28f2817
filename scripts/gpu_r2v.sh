#!/bin/bash
# speculative exponentials (no row max in the common case)
source scripts/gpu_probe.sh
run att 600 tests/test_gpu_3_kernels.py -k attention
echo "=== att_bench" | tee -a $S
timeout 900 python scripts/att_bench.py fp16 bf16 tf32 2>&1 | grep "^ATT" | tee -a $S
run models 900 tests/test_gpu_0_models.py -x
bench default --no-cpu-baseline
