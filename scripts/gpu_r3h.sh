#!/bin/bash
# ART-V token graphs captured without emptying the caching allocator: tests + the ART-V bench twice on one box
source scripts/gpu_probe.sh
run artv 900 tests/test_gpu_0_models.py tests/test_gpu_1_fullsize.py -k "artv"
bench artv_a --no-cpu-baseline --workload artv --steps 3 --warmup 1
bench artv_b --no-cpu-baseline --workload artv --steps 3 --warmup 1
