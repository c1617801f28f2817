#!/bin/bash
# r1s: new GEMM defaults (TMA-store epilogue, 256-wide tiles where the cost model picks them): full suite, bench, artv bench
mkdir -p gpurun_out
S=gpurun_out/summary_r1s.txt; rm -f $S
echo "=== full GPU suite, one process" | tee -a $S
SECONDS=0
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r1s_full.log 2>&1; echo "exit $? after ${SECONDS}s" | tee -a $S; tail -n 4 gpurun_out/r1s_full.log | cut -c1-300 | tee -a $S
echo "=== sweep default dispatch tf32 / bf16" | tee -a $S
for p in tf32 bf16; do timeout 100 python scripts/gemm_sweep.py --child $p | tail -1 | tee -a $S; done
echo "=== bench tf32 (default)" | tee -a $S
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1s_bench_tf32.json 2> gpurun_out/r1s_bench_tf32.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1s_bench_tf32.json | tee -a $S
echo "=== bench bf16" | tee -a $S
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/r1s_bench_bf16.json 2> gpurun_out/r1s_bench_bf16.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r1s_bench_bf16.json | tee -a $S
echo "=== bench artv tf32" | tee -a $S
timeout 900 python bench.py --workload artv --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1s_bench_artv.json 2> gpurun_out/r1s_bench_artv.err; echo "exit $?" | tee -a $S; cut -c1-400 gpurun_out/r1s_bench_artv.json | tee -a $S; tail -2 gpurun_out/r1s_bench_artv.err | tee -a $S
echo "=== bench --impl reference" | tee -a $S
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1s_bench_ref.json 2> gpurun_out/r1s_bench_ref.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1s_bench_ref.json | tee -a $S
