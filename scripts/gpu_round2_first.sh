#!/bin/bash
# First GPU call of the next round: validate the code written after round 1's GPU budget ran out, then A/B it.
#   1. opt-in tests (MMVID_TEST_EXPERIMENTAL=1): transposed conv tile (MMVID_CONV_SWAP=1)
#   2. VQGAN decode A/B: bench.py with / without MMVID_CONV_SWAP (the Cout = 128 convs are 10.7 % of the step)
#   3. the regular suite + default bench as the baseline of the round
mkdir -p gpurun_out
S=gpurun_out/summary_round2_first.txt; rm -f $S
echo "=== experimental tests" | tee -a $S
MMVID_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "experimental" 2>&1 | tail -4 | cut -c1-300 | tee -a $S
echo "=== model tests with MMVID_CONV_SWAP=1" | tee -a $S
MMVID_CONV_SWAP=1 timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider -k "vae or tensor_core or generate or batched" 2>&1 | tail -3 | cut -c1-300 | tee -a $S
for sw in 0 1; do
  echo "=== bench tf32 MMVID_CONV_SWAP=$sw" | tee -a $S
  MMVID_CONV_SWAP=$sw timeout 600 python bench.py --no-cpu-baseline > gpurun_out/round2_first_bench_swap$sw.json 2> gpurun_out/round2_first_bench_swap$sw.err; echo "exit $?" | tee -a $S
  cut -c1-260 gpurun_out/round2_first_bench_swap$sw.json | tee -a $S
done
echo "=== full GPU suite, one process" | tee -a $S
SECONDS=0
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/round2_first_full.log 2>&1; echo "exit $? after ${SECONDS}s" | tee -a $S; tail -n 3 gpurun_out/round2_first_full.log | cut -c1-300 | tee -a $S
