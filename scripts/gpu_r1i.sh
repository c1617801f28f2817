#!/bin/bash
# fused PDL decode + attention pipeline trace
mkdir -p gpurun_out
S=gpurun_out/summary_r1i.txt; rm -f $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py 2>&1 | tail -10 | tee -a $S
echo "=== artv tests" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "artv or decode" > gpurun_out/r1i_a.log 2>&1; echo "exit $?" | tee -a $S; tail -n 6 gpurun_out/r1i_a.log | cut -c1-300 | tee -a $S
echo "=== bench artv fused" | tee -a $S
timeout 600 python bench.py --workload artv --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r1i_artv_fused.json 2> gpurun_out/r1i_artv_fused.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1i_artv_fused.json | tee -a $S
echo "=== attention trace tf32" | tee -a $S
timeout 300 python scripts/att_trace.py tf32 2>&1 | tail -22 | tee -a $S
echo "=== attention trace bf16" | tee -a $S
timeout 300 python scripts/att_trace.py bf16 2>&1 | tail -22 | tee -a $S
