#!/bin/bash
# experiment: bf16 P by PRMT truncation (no XU conversion) - does XU relief pay at equal instruction count?
source scripts/gpu_probe.sh
for t in 0 1 0 1; do
  echo "trunc=$t" | tee -a $S
  MMVID_ATT_TRUNC=$t timeout 120 python scripts/att_bench.py one bf16 5 2 0 1 1 2>&1 | grep "^ATT" | cut -c1-60,230-330 | tee -a $S
done
