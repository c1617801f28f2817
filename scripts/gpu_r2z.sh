#!/bin/bash
# pair-GEMM epilogue: bias by shuffle, fetched one tile ahead (16-bit results)
source scripts/gpu_probe.sh
run kernels 900 tests/test_gpu_3_kernels.py -x
echo "=== gemm trace" | tee -a $S
timeout 300 python scripts/gemm_trace2.py fp16 2>&1 | grep "^==\|tile [0-5]:" | tee -a $S
run models 900 tests/test_gpu_0_models.py -x
bench default --no-cpu-baseline
