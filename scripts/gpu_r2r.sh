#!/bin/bash
# trace-free production kernels (TRACE template), v5 vs v6, extended v6 timeline, MUFU / N=64 MMA microbenchmarks
source scripts/gpu_probe.sh
run att 600 tests/test_gpu_3_kernels.py -k attention
echo "=== att_bench" | tee -a $S
timeout 900 python scripts/att_bench.py fp16 tf32 2>&1 | grep "^ATT" | tee -a $S
echo "=== microbench + trace v5 fp16" | tee -a $S
MMVID_ATT_IMPL=5 MMVID_ATT_POLY=2 timeout 200 python scripts/att_trace3.py fp16 2>&1 | grep "^MMA\|^MUFU\|^tile" | tee -a $S
echo "=== trace6 fp16" | tee -a $S
timeout 200 python scripts/att_trace6.py fp16 2>&1 | tee -a $S
echo "=== trace6 fp16 spin" | tee -a $S
MMVID_ATT_SPIN=1 timeout 200 python scripts/att_trace6.py fp16 2>&1 | tail -30 | tee -a $S
run train 900 tests/test_gpu_2_training.py -x
MMVID_ATT_IMPL=5 bench v5 --no-cpu-baseline
bench v6 --no-cpu-baseline
