#!/bin/bash
# compute-sanitizer memcheck over the kernels that changed this round (small shapes)
source scripts/gpu_probe.sh
for sel in "conv_epilogue_groupnorm_partials" "attention_speculative" "attention_masks and fp16" "mp_sample or mp_keep" "linear and fp16"; do
  echo "=== memcheck: $sel" | tee -a $S
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 0 python -m pytest tests/test_gpu_3_kernels.py -x -q -m gpu -p no:cacheprovider -k "$sel" > gpurun_out/${TAG}_mc.log 2>&1
  echo "exit $?" | tee -a $S
  grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Error" gpurun_out/${TAG}_mc.log | sort | uniq -c | head -8 | tee -a $S
done
