#!/bin/bash
source scripts/gpu_probe.sh
echo "=== gemm trace fp16" | tee -a $S
timeout 300 python scripts/gemm_trace2.py fp16 2>&1 | tail -36 | tee -a $S
echo "=== gemm trace tf32" | tee -a $S
timeout 300 python scripts/gemm_trace2.py tf32 2>&1 | tail -36 | tee -a $S
echo "=== library yardstick" | tee -a $S
timeout 300 python scripts/library_yardstick.py 2>&1 | tail -30 | tee -a $S
