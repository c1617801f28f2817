#!/bin/bash
# GroupNorm statistics fused into the decoder convs' epilogues
source scripts/gpu_probe.sh
run gn 600 tests/test_gpu_3_kernels.py -k "groupnorm or conv"
run vae 600 tests/test_gpu_0_models.py -k vae
MMVID_GN_FUSE=0 bench gn0 --no-cpu-baseline
bench gn1 --no-cpu-baseline
