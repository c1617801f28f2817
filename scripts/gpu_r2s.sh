#!/bin/bash
# v7 attention: 16 softmax warps (two per query row)
source scripts/gpu_probe.sh
MMVID_ATT_IMPL=7 run att7 600 tests/test_gpu_3_kernels.py -k attention
echo "=== att_bench" | tee -a $S
timeout 900 python scripts/att_bench.py fp16 bf16 2>&1 | grep "^ATT" | tee -a $S
echo "=== trace v7 fp16" | tee -a $S
MMVID_ATT_IMPL=7 MMVID_ATT_POLY=2 timeout 200 python scripts/att_trace3.py fp16 2>&1 | grep "^tile\|^n=" | tee -a $S
MMVID_ATT_IMPL=7 bench v7 --no-cpu-baseline
bench train --no-cpu-baseline --workload train --precision tf32 --steps 3 --warmup 3
