#!/bin/bash
# four GPUs of one box: weak-scaling bench (own arm) and the reference arm under torchrun (rank 0 works, the others exit 0)
source scripts/gpu_probe.sh
nvidia-smi --query-gpu=index,name --format=csv | tee -a $S
echo "=== bench --gpus 4" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n4.json 2> gpurun_out/${TAG}_bench_n4.err; echo "exit $?" | tee -a $S
grep '"metric"' gpurun_out/${TAG}_bench_n4.json | cut -c1-700 | tee -a $S
tail -n 3 gpurun_out/${TAG}_bench_n4.err | cut -c1-300 >> $S
echo "=== bench --impl reference --gpus 4" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref4.json 2> gpurun_out/${TAG}_bench_ref4.err; echo "exit $?" | tee -a $S
grep '"impl"' gpurun_out/${TAG}_bench_ref4.json | cut -c1-300 | tee -a $S
