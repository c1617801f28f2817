#!/bin/bash
# Staged GPU validation: each stage is its own process so a trapped kernel cannot poison later stages.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest "$@" -m gpu -q -s -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run k_fp32 900 tests/test_gpu_kernels.py -k "not tf32 and not bf16 and not tensor_core"
run k_lin_tf32 300 tests/test_gpu_kernels.py -k "linear and tf32"
run k_conv_tf32 300 tests/test_gpu_kernels.py -k "conv3x3_tensor_core"
run k_lin_bf16 300 tests/test_gpu_kernels.py -k "linear and bf16 or bf16_output"
run k_att_tf32 300 tests/test_gpu_kernels.py -k "attention and tf32"
run k_att_bf16 300 tests/test_gpu_kernels.py -k "attention and bf16"
run m_fp32 1500 tests/test_gpu_models.py -k "(fp32 or vae or generate or kv_cache) and not tensor_core"
run m_tc 1200 tests/test_gpu_models.py -k "tf32 or bf16 or batched or tensor_core"
