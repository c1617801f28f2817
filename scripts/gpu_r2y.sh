#!/bin/bash
# two GPUs of one box: NCCL sharded generate_images test, weak-scaling bench at N = 2 (own arm and reference arm)
source scripts/gpu_probe.sh
nvidia-smi --query-gpu=index,name --format=csv | tee -a $S
run dist 900 tests/test_gpu_4_dist.py
echo "=== bench --gpus 2" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "exit $?" | tee -a $S
grep '"metric"' gpurun_out/${TAG}_bench_n2.json | cut -c1-600 | tee -a $S
tail -n 3 gpurun_out/${TAG}_bench_n2.err | cut -c1-300 >> $S
echo "=== bench --gpus 2 train" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --workload train --precision tf32 > gpurun_out/${TAG}_bench_n2_train.json 2> gpurun_out/${TAG}_bench_n2_train.err; echo "exit $?" | tee -a $S
grep '"metric"' gpurun_out/${TAG}_bench_n2_train.json | cut -c1-400 | tee -a $S
