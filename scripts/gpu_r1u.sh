#!/bin/bash
# r1u: TMA-store epilogue in the CTA-pair GEMM: correctness, sweep, traces
mkdir -p gpurun_out
S=gpurun_out/summary_r1u.txt; rm -f $S
for v in 128 256; do
  echo "=== kernel tests 2CTA=$v" | tee -a $S
  MMVID_GEMM_2CTA=$v timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "linear" 2>&1 | tail -3 | cut -c1-300 | tee -a $S
done
echo "=== sweep 2cta" | tee -a $S
timeout 400 python scripts/gemm_sweep.py --2cta 2>&1 | grep tf32 | tee -a $S
MMVID_GEMM_2CTA=256 timeout 100 python scripts/gemm_trace.py tf32 > gpurun_out/r1u_trace_2cta256.txt 2>&1
grep -E "^==|tile [23]:" gpurun_out/r1u_trace_2cta256.txt | cut -c1-330 | tee -a $S
echo "=== bench tf32 2CTA=256 forced" | tee -a $S
MMVID_GEMM_2CTA=256 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1u_bench_2cta256.json 2> gpurun_out/r1u_bench_2cta256.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1u_bench_2cta256.json | tee -a $S
