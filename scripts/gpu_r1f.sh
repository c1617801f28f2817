#!/bin/bash
# fused optimiser + train / visual-control bench workloads
mkdir -p gpurun_out
S=gpurun_out/summary_r1f.txt; rm -f $S
echo "=== training tests" | tee -a $S
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r1f_train.log 2>&1; echo "exit $?" | tee -a $S; tail -n 8 gpurun_out/r1f_train.log | cut -c1-300 | tee -a $S
echo "=== bench train shape A b4" | tee -a $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/r1f_train_A.json 2> gpurun_out/r1f_train_A.err; echo "exit $?" | tee -a $S; cut -c1-1500 gpurun_out/r1f_train_A.json | tee -a $S; tail -n 5 gpurun_out/r1f_train_A.err | cut -c1-300 | tee -a $S
echo "=== bench bert visuals=1 (config 4 per-GPU slice)" | tee -a $S
timeout 600 python bench.py --visuals 1 --no-cpu-baseline > gpurun_out/r1f_bert_v1.json 2> gpurun_out/r1f_bert_v1.err; echo "exit $?" | tee -a $S; cut -c1-1200 gpurun_out/r1f_bert_v1.json | tee -a $S; tail -n 5 gpurun_out/r1f_bert_v1.err | cut -c1-300 | tee -a $S
