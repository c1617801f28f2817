#!/bin/bash
source scripts/gpu_probe.sh
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py --impl stream 2>&1 | tail -4 | tee -a $S
run full 2400 tests/ -x
bench artv_fp16 --no-cpu-baseline --workload artv --steps 2 --warmup 1
bench default
