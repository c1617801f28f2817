#!/bin/bash
source scripts/gpu_probe.sh
echo "=== microbench" | tee -a $S
timeout 200 python scripts/att_trace3.py fp16 2>&1 | grep "^MUFU\|^PIPE" | tee -a $S
echo "=== gemm trace" | tee -a $S
timeout 300 python scripts/gemm_trace2.py fp16 2>&1 | grep "^==\|tile [0-5]:" | head -16 | tee -a $S
