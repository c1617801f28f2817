#!/bin/bash
source scripts/gpu_probe.sh
run gemm 600 tests/test_gpu_3_kernels.py -k "linear or pair or qkv or 16bit"
run models 900 tests/test_gpu_0_models.py -k "transformer or bert_forward"
echo "=== gemm trace fp16" | tee -a $S
timeout 300 python scripts/gemm_trace2.py fp16 2>&1 | grep -E "==|tile [0-2]:" | tee -a $S
echo "=== gemm trace tf32" | tee -a $S
timeout 300 python scripts/gemm_trace2.py tf32 2>&1 | grep -E "==|tile [0-2]:" | tee -a $S
bench default --no-cpu-baseline
bench tf32 --no-cpu-baseline --precision tf32
