#!/bin/bash
# P conversions off the XU pipe (MMVID_ATT_CVT = how many of every 4)
source scripts/gpu_probe.sh
echo "=== att_bench cvt" | tee -a $S
for prec in fp16 bf16; do for c in 0 1 2 3; do
  echo "cvt4=$c" | tee -a $S
  MMVID_ATT_CVT=$c timeout 120 python scripts/att_bench.py one $prec 5 2 0 1 0 2>&1 | grep "^ATT" | tee -a $S
done; done
MMVID_ATT_CVT=2 run att 600 tests/test_gpu_3_kernels.py -k attention
