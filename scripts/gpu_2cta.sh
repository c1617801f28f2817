#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary_2cta.txt; rm -f $S
for bn in 128 256; do
  echo "=== 2cta BN=$bn linear tests" | tee -a $S
  MMVID_GEMM_2CTA=$bn timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "linear and (tf32 or bf16)" > gpurun_out/t2cta_$bn.log 2>&1; echo "exit $?" | tee -a $S; tail -n 6 gpurun_out/t2cta_$bn.log | cut -c1-300 | tee -a $S
done
echo "=== sweep 2cta" | tee -a $S
timeout 600 python scripts/gemm_sweep.py --2cta 2>&1 | cut -c1-600 | tee -a $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py 2>&1 | tail -8 | tee -a $S
