#!/bin/bash
# re-run failed tests, profile the decode paths
mkdir -p gpurun_out
S=gpurun_out/summary_r1h.txt; rm -f $S
echo "=== full-size property tests + optimiser" | tee -a $S
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider > gpurun_out/r1h_b.log 2>&1; echo "exit $?" | tee -a $S; tail -n 14 gpurun_out/r1h_b.log | cut -c1-300 | tee -a $S
echo "=== ncu launch list, native decode (pos 1300)" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_decode_native.csv python scripts/decode_microbench.py --impl native --pos 1300 --reps 2 --warm 1 > gpurun_out/ncu_dn.log 2>&1; echo "exit $?" | tee -a $S
echo "=== ncu full, persistent decode (pos 1300)" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:artv_decode_persistent -s 1 -c 1 -o gpurun_out/prof_decode_persistent python scripts/decode_microbench.py --impl persistent --pos 1300 --reps 2 --warm 1 > gpurun_out/ncu_dp.log 2>&1; echo "exit $?" | tee -a $S
ls -la gpurun_out/*.ncu-rep | tee -a $S
