#!/bin/bash
source scripts/gpu_probe.sh
echo "=== ncu full attention (fp16)" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 2 -c 1 -o gpurun_out/${TAG}_prof_attention python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/${TAG}_ncu_att.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full gemm_tc2 (fp16)" | tee -a $S
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc2 -s 8 -c 4 -o gpurun_out/${TAG}_prof_gemm2 python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "exit $?" | tee -a $S
ls -la gpurun_out/${TAG}_prof_* | tee -a $S
