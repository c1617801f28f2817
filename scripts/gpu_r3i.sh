#!/bin/bash
source scripts/gpu_probe.sh
run new 900 tests/test_gpu_1_fullsize.py -k "programmatic or pinned"
grep "PDL off" gpurun_out/${TAG}_new.log | tee -a $S
