#!/bin/bash
source scripts/gpu_probe.sh
echo "=== decode trace" | tee -a $S
timeout 300 python scripts/decode_trace.py 1300 2>&1 | tail -18 | tee -a $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py --impl stream 2>&1 | tail -4 | tee -a $S
run artv 900 tests/test_gpu_0_models.py tests/test_gpu_1_fullsize.py -k "artv"
bench artv_fp16 --no-cpu-baseline --workload artv --steps 2 --warmup 1
bench default
