#!/bin/bash
source scripts/gpu_probe.sh
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py --impl stream,fused 2>&1 | tail -8 | tee -a $S
run conv16 600 tests/test_gpu_3_kernels.py -k "conv or groupnorm or upsample"
run vae16 600 tests/test_gpu_0_models.py -k "vae"
bench fp16_vae16 --no-cpu-baseline --precision fp16 --vae-precision fp16
bench fp16_vaetf32 --no-cpu-baseline --precision fp16 --vae-precision tf32
MMVID_CONV_SWAP=1 bench fp16_vaetf32_swap --no-cpu-baseline --precision fp16 --vae-precision tf32
