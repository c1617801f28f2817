#!/bin/bash
# r1w: new GEMM dispatch (CTA-pair 192/256 tiles + TMA-store epilogue): full suite, sweeps, benches
mkdir -p gpurun_out
S=gpurun_out/summary_r1w.txt; rm -f $S
echo "=== full GPU suite, one process" | tee -a $S
SECONDS=0
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r1w_full.log 2>&1; echo "exit $? after ${SECONDS}s" | tee -a $S; tail -n 4 gpurun_out/r1w_full.log | cut -c1-300 | tee -a $S
echo "=== sweep default dispatch tf32 / bf16, and bf16 forced pair tiles" | tee -a $S
for p in tf32 bf16; do timeout 100 python scripts/gemm_sweep.py --child $p | tail -1 | tee -a $S; done
for v in 1 128 192 256; do echo "bf16 2CTA=$v" | tee -a $S; MMVID_GEMM_2CTA=$v timeout 100 python scripts/gemm_sweep.py --child bf16 | tail -1 | tee -a $S; done
echo "=== bench tf32 (default)" | tee -a $S
timeout 600 python bench.py > gpurun_out/r1w_bench_tf32.json 2> gpurun_out/r1w_bench_tf32.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1w_bench_tf32.json | tee -a $S
echo "=== bench bf16" | tee -a $S
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/r1w_bench_bf16.json 2> gpurun_out/r1w_bench_bf16.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1w_bench_bf16.json | tee -a $S
echo "=== bench train tf32" | tee -a $S
timeout 600 python bench.py --workload train --no-cpu-baseline --batch 2 > gpurun_out/r1w_bench_train.json 2> gpurun_out/r1w_bench_train.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1w_bench_train.json | tee -a $S; tail -2 gpurun_out/r1w_bench_train.err | tee -a $S
