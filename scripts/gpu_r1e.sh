#!/bin/bash
# validation after 2-CTA default + persistent decode rework
mkdir -p gpurun_out
S=gpurun_out/summary_r1e.txt; rm -f $S
echo "=== pytest -m gpu (all)" | tee -a $S
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r1e_pytest.log 2>&1; echo "exit $?" | tee -a $S; tail -n 8 gpurun_out/r1e_pytest.log | cut -c1-300 | tee -a $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py 2>&1 | tail -8 | tee -a $S
echo "=== bench tf32" | tee -a $S
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1e_bench_tf32.json 2> gpurun_out/r1e_bench_tf32.err; echo "exit $?" | tee -a $S; cut -c1-900 gpurun_out/r1e_bench_tf32.json | tee -a $S
echo "=== bench tf32, 2cta off" | tee -a $S
MMVID_GEMM_2CTA=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1e_bench_tf32_1cta.json 2> gpurun_out/r1e_bench_tf32_1cta.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1e_bench_tf32_1cta.json | tee -a $S
echo "=== bench artv native / persistent" | tee -a $S
timeout 600 python bench.py --workload artv --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r1e_artv_native.json 2> gpurun_out/r1e_artv_native.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1e_artv_native.json | tee -a $S
MMVID_ARTV_DECODE=persistent timeout 600 python bench.py --workload artv --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r1e_artv_pers.json 2> gpurun_out/r1e_artv_pers.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1e_artv_pers.json | tee -a $S
