#!/bin/bash
# end-of-round evidence: ncu launch list + --set full captures of the final kernels, all bench arms, smoke()
source scripts/gpu_probe.sh
echo "=== smoke" | tee -a $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee -a $S
echo "=== ncu launch list (fp16 default, one step)" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "exit $?" | tee -a $S
python scripts/summarise_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1; head -n 45 gpurun_out/${TAG}_launches.md | tee -a $S
echo "=== ncu full attention (fp16)" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 2 -c 1 -o gpurun_out/${TAG}_prof_attention python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/${TAG}_ncu_att.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full gemm_tc2 (fp16)" | tee -a $S
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc2 -s 8 -c 4 -o gpurun_out/${TAG}_prof_gemm2 python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "exit $?" | tee -a $S
ls -la gpurun_out/${TAG}_prof_* | tee -a $S
bench artv --no-cpu-baseline --workload artv --steps 2 --warmup 1
bench train --no-cpu-baseline --workload train --precision tf32 --steps 3 --warmup 3
bench eager --impl eager --steps 1 --warmup 1
bench reference --impl reference --steps 2 --warmup 1
bench default
