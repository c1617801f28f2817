#!/bin/bash
# GroupNorm partials with one shuffle butterfly (shifted sums)
source scripts/gpu_probe.sh
run gn 600 tests/test_gpu_3_kernels.py -k "groupnorm or conv"
run vae 600 tests/test_gpu_0_models.py -k vae
bench default --no-cpu-baseline
