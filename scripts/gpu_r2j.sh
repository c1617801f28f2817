#!/bin/bash
source scripts/gpu_probe.sh
run samp 600 tests/test_gpu_3_kernels.py -k "mp_sample or mp_keep"
run mp 900 tests/test_gpu_0_models.py tests/test_gpu_1_fullsize.py -k "mask_predict or batched or generation or generate"
bench default --no-cpu-baseline
MMVID_SAMPLER=torch bench torch_sampler --no-cpu-baseline
