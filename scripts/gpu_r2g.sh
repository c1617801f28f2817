#!/bin/bash
# r2g: final validation of the round: full GPU suite in one process, default bench (with cpu baseline), reference arm,
# ncu launch list + full captures of the attention / GEMM kernels
mkdir -p gpurun_out
S=gpurun_out/summary_r2g.txt; rm -f $S
echo "=== full GPU suite, one process" | tee -a $S
SECONDS=0
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2g_full.log 2>&1; echo "exit $? after ${SECONDS}s" | tee -a $S; tail -n 6 gpurun_out/r2g_full.log | cut -c1-300 | tee -a $S
echo "=== smoke" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a $S
echo "=== bench tf32 (default)" | tee -a $S
timeout 600 python bench.py > gpurun_out/r2g_bench_tf32.json 2> gpurun_out/r2g_bench_tf32.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r2g_bench_tf32.json | tee -a $S
echo "=== ncu launch list" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --profile --no-cpu-baseline > gpurun_out/r2g_ncu_launch.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full attention + gemm" | tee -a $S
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o gpurun_out/r2g_prof_attention python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/r2g_ncu_att.log 2>&1; echo "exit $?" | tee -a $S
timeout 400 ncu --set full --clock-control none -k regex:gemm_tc2 -c 5 -o gpurun_out/r2g_prof_gemm2 python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/r2g_ncu_gemm.log 2>&1; echo "exit $?" | tee -a $S
