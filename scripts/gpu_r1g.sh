#!/bin/bash
# decode GEMV rewrite + full-size property tests + fused optimiser test
mkdir -p gpurun_out
S=gpurun_out/summary_r1g.txt; rm -f $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py 2>&1 | tail -8 | tee -a $S
echo "=== decode / artv / optimiser tests" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_training.py -m gpu -q -x -p no:cacheprovider -k "decode or artv or fused or optimizer" > gpurun_out/r1g_a.log 2>&1; echo "exit $?" | tee -a $S; tail -n 6 gpurun_out/r1g_a.log | cut -c1-300 | tee -a $S
echo "=== full-size property tests" | tee -a $S
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r1g_b.log 2>&1; echo "exit $?" | tee -a $S; tail -n 30 gpurun_out/r1g_b.log | cut -c1-300 | tee -a $S
echo "=== bench artv native / persistent" | tee -a $S
timeout 600 python bench.py --workload artv --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r1g_artv_native.json 2> gpurun_out/r1g_artv_native.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1g_artv_native.json | tee -a $S
MMVID_ARTV_DECODE=persistent timeout 600 python bench.py --workload artv --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r1g_artv_pers.json 2> gpurun_out/r1g_artv_pers.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1g_artv_pers.json | tee -a $S
