#!/bin/bash
source scripts/gpu_probe.sh
echo "=== attention trace fp16 poly 2" | tee -a $S
MMVID_ATT_POLY=2 timeout 300 python scripts/att_trace3.py fp16 2>&1 | tail -12 | tee -a $S
