#!/bin/bash
# v6 attention (persistent, decoupled tile streams): correctness of the attention tests, variant timing, timeline, bench
source scripts/gpu_probe.sh
run att 600 tests/test_gpu_3_kernels.py -k attention
echo "=== att_bench" | tee -a $S
timeout 900 python scripts/att_bench.py fp16 bf16 2>&1 | grep "^ATT" | tee -a $S
echo "=== trace v6 fp16" | tee -a $S
timeout 200 python scripts/att_trace3.py fp16 2>&1 | grep -v "^MMA" | tee -a $S
echo "=== trace v6 fp16 poly 4" | tee -a $S
MMVID_ATT_POLY=4 timeout 200 python scripts/att_trace3.py fp16 2>&1 | grep -v "^MMA" | tail -12 | tee -a $S
run models 900 tests/test_gpu_0_models.py -x
bench default --no-cpu-baseline
bench artv_fp16 --no-cpu-baseline --workload artv --steps 2 --warmup 1
