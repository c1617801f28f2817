#!/bin/bash
# r1p: attention v5 as default: full GPU suite in ONE process, bench, GEMM pipeline trace, ncu launch list + full captures
mkdir -p gpurun_out
S=gpurun_out/summary_r1p.txt; rm -f $S
echo "=== gemm trace tf32" | tee -a $S
timeout 200 python scripts/gemm_trace.py tf32 > gpurun_out/r1p_gemm_trace_tf32.txt 2>&1; echo "exit $?" | tee -a $S
echo "=== full GPU suite, one process (as the driver runs it)" | tee -a $S
SECONDS=0
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r1p_full.log 2>&1; echo "exit $? after ${SECONDS}s" | tee -a $S; tail -n 4 gpurun_out/r1p_full.log | cut -c1-300 | tee -a $S
echo "=== bench tf32 (default)" | tee -a $S
timeout 600 python bench.py > gpurun_out/r1p_bench_tf32.json 2> gpurun_out/r1p_bench_tf32.err; echo "exit $?" | tee -a $S; cut -c1-330 gpurun_out/r1p_bench_tf32.json | tee -a $S
echo "=== bench bf16" | tee -a $S
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/r1p_bench_bf16.json 2> gpurun_out/r1p_bench_bf16.err; echo "exit $?" | tee -a $S; cut -c1-330 gpurun_out/r1p_bench_bf16.json | tee -a $S
echo "=== ncu launch list" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1p_launches.csv python bench.py --profile --no-cpu-baseline > gpurun_out/r1p_ncu_launch.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full attention + gemm" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o gpurun_out/r1p_prof_attention python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/r1p_ncu_att.log 2>&1; echo "exit $?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 4 -o gpurun_out/r1p_prof_gemm python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/r1p_ncu_gemm.log 2>&1; echo "exit $?" | tee -a $S
ls -la gpurun_out | tail -n 12 | tee -a $S
