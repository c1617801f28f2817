#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary2.txt
  timeout $to python -m pytest "$@" -m gpu -q -s -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary2.txt; tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary2.txt; }
rm -f gpurun_out/summary2.txt
run k_conv_tf32 300 tests/test_gpu_kernels.py -k "conv3x3_tensor_core or upsample2x"
run m_vae_tc 600 tests/test_gpu_models.py -k "tensor_core"
echo "=== bench tf32" | tee -a gpurun_out/summary2.txt
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "exit $?" | tee -a gpurun_out/summary2.txt
tail -c 3000 gpurun_out/bench_tf32.json | tee -a gpurun_out/summary2.txt
echo "=== bench bf16" | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "exit $?" | tee -a gpurun_out/summary2.txt
tail -c 2500 gpurun_out/bench_bf16.json | tee -a gpurun_out/summary2.txt
echo "=== bench tf32, fp32 vae" | tee -a gpurun_out/summary2.txt
timeout 600 python bench.py --steps 2 --warmup 3 --vae-precision fp32 --no-cpu-baseline > gpurun_out/bench_tf32_vaefp32.json 2> gpurun_out/bench_tf32_vaefp32.err; echo "exit $?" | tee -a gpurun_out/summary2.txt
tail -c 1500 gpurun_out/bench_tf32_vaefp32.json | tee -a gpurun_out/summary2.txt
echo "=== ncu launch list" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary2.txt
echo "=== ncu full attention + gemm" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 2 -o gpurun_out/prof_attention_r1 python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/ncu_att.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 4 -o gpurun_out/prof_gemm_r1 python bench.py --profile --no-cpu-baseline --mp-steps 1 > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary2.txt
ls -la gpurun_out | tee -a gpurun_out/summary2.txt
