#!/bin/bash
# r1r: TMA-store GEMM epilogue: correctness (kernel + model tests), sweep BN 128/256, traces
mkdir -p gpurun_out
S=gpurun_out/summary_r1r.txt; rm -f $S
export MMVID_GEMM_TMA_STORE=1
for bn in 0 256; do
  echo "=== kernel tests TMA store BN=$bn" | tee -a $S
  MMVID_GEMM_BN=$bn timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "linear or qkv or conv or groupnorm" 2>&1 | tail -3 | cut -c1-300 | tee -a $S
done
echo "=== model tests TMA store (default BN)" | tee -a $S
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_training.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300 | tee -a $S
echo "=== sweep (1-CTA only)" | tee -a $S
for bn in 128 256; do for p in tf32; do echo "TMA BN=$bn $p" | tee -a $S; MMVID_GEMM_2CTA=1 MMVID_GEMM_BN=$bn timeout 100 python scripts/gemm_sweep.py --child $p | tail -1 | tee -a $S; done; done
echo "old epilogue BN=128 tf32" | tee -a $S; MMVID_GEMM_TMA_STORE=0 MMVID_GEMM_2CTA=1 MMVID_GEMM_BN=128 timeout 100 python scripts/gemm_sweep.py --child tf32 | tail -1 | tee -a $S
MMVID_GEMM_BN=256 timeout 100 python scripts/gemm_trace.py tf32 > gpurun_out/r1r_trace_tma_bn256.txt 2>&1
MMVID_GEMM_BN=128 timeout 100 python scripts/gemm_trace.py tf32 > gpurun_out/r1r_trace_tma_bn128.txt 2>&1
grep -E "^==|tile [23]:" gpurun_out/r1r_trace_tma_bn256.txt | cut -c1-330 | tee -a $S
grep -E "^==|tile [23]:" gpurun_out/r1r_trace_tma_bn128.txt | cut -c1-330 | tee -a $S
for bn in 0 256; do
echo "=== bench tf32 TMA store BN=$bn" | tee -a $S
MMVID_GEMM_BN=$bn timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1r_bench_bn$bn.json 2> gpurun_out/r1r_bench_bn$bn.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1r_bench_bn$bn.json | tee -a $S
done
