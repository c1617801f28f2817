#!/bin/bash
source scripts/gpu_probe.sh
echo "=== decode trace" | tee -a $S
timeout 300 python scripts/decode_trace.py 1300 2>&1 | tail -16 | tee -a $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py --impl stream 2>&1 | tail -4 | tee -a $S
run artv 900 tests/test_gpu_0_models.py tests/test_gpu_1_fullsize.py -k "artv"
bench artv_fp16 --no-cpu-baseline --workload artv --steps 2 --warmup 1
echo "=== GELU tanh parity" | tee -a $S
MMVID_GELU_TANH=1 timeout 600 python -m pytest tests/test_gpu_0_models.py -q -m gpu -p no:cacheprovider -rA -k "bert_forward and fp16" 2>&1 | grep -E "relerr|passed|failed" | tee -a $S
MMVID_GELU_TANH=1 bench gelu_tanh --no-cpu-baseline
bench default --no-cpu-baseline
