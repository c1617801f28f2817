#!/bin/bash
source scripts/gpu_probe.sh
run base 1500 tests/ -x -k "not fp16 and not 16bit"
run k16 600 tests/test_gpu_3_kernels.py -k "fp16 or 16bit or bf16"
run m16 900 tests/test_gpu_0_models.py -k "fp16 or bf16"
bench tf32 --no-cpu-baseline --precision tf32
bench fp16 --no-cpu-baseline --precision fp16
bench bf16 --no-cpu-baseline --precision bf16
