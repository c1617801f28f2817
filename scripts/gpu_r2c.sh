#!/bin/bash
source scripts/gpu_probe.sh
run att 600 tests/test_gpu_3_kernels.py -k "attention or qkv"
run stream_tiny 600 tests/test_gpu_0_models.py -k "streaming or artv"
run stream_A 900 tests/test_gpu_1_fullsize.py -k "streaming"
run full 1800 tests/ -x
bench artv_fp16 --no-cpu-baseline --workload artv --precision fp16 --steps 1 --warmup 1
bench artv_fp32 --no-cpu-baseline --workload artv --precision fp32 --steps 1 --warmup 1
