#!/bin/bash
source scripts/gpu_probe.sh
for poly in 2 4; do
echo "=== attention trace fp16 poly $poly" | tee -a $S
MMVID_ATT_POLY=$poly timeout 300 python scripts/att_trace3.py fp16 2>&1 | tail -24 | tee -a $S
done
echo "=== attention bench" | tee -a $S
timeout 600 python scripts/att_bench.py 2>&1 | tail -20 | cut -c1-300 | tee -a $S
