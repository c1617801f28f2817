#!/bin/bash
# full GPU validation (what the driver runs) + training tests + launch list + ncu captures
mkdir -p gpurun_out
S=gpurun_out/summary_full.txt; rm -f $S
echo "=== pytest -m gpu (all)" | tee -a $S
timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/full_pytest.log 2>&1; echo "exit $?" | tee -a $S; tail -n 15 gpurun_out/full_pytest.log | tee -a $S
echo "=== smoke" | tee -a $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $S
echo "=== ncu launch list" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --profile --no-cpu-baseline > gpurun_out/ncu_launch_c.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -c 2 -o gpurun_out/prof_attention_r1c python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/ncu_att_c.log 2>&1; echo "exit $?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 5 -o gpurun_out/prof_gemm_r1c python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/ncu_gemm_c.log 2>&1; echo "exit $?" | tee -a $S
