#!/bin/bash
source scripts/gpu_probe.sh
echo "=== att_bench" | tee -a $S
timeout 900 python scripts/att_bench.py fp16 2>&1 | grep "^ATT" | cut -c1-330 | tee -a $S
