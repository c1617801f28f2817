#!/bin/bash
# final launch list (after the GroupNorm-statistics fusion)
source scripts/gpu_probe.sh
echo "=== ncu launch list (fp16 default, one step)" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "exit $?" | tee -a $S
python scripts/summarise_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1; head -n 40 gpurun_out/${TAG}_launches.md | tee -a $S
