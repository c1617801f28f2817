#!/bin/bash
source scripts/gpu_probe.sh
echo "=== pipe microbench" | tee -a $S
MMVID_ATT_IMPL=5 MMVID_ATT_POLY=2 timeout 200 python scripts/att_trace3.py fp16 2>&1 | grep "^MMA\|^MUFU\|^PIPE" | tee -a $S
