#!/bin/bash
# fused PDL decode (fixed smem opt-in) + attention with polling waits
mkdir -p gpurun_out
S=gpurun_out/summary_r1j.txt; rm -f $S
echo "=== decode microbench" | tee -a $S
timeout 300 python scripts/decode_microbench.py 2>&1 | tail -10 | tee -a $S
echo "=== attention trace tf32" | tee -a $S
timeout 300 python scripts/att_trace.py tf32 2>&1 | tail -8 | tee -a $S
echo "=== attention + artv tests" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider -k "attention or artv or decode" > gpurun_out/r1j_a.log 2>&1; echo "exit $?" | tee -a $S; tail -n 6 gpurun_out/r1j_a.log | cut -c1-300 | tee -a $S
echo "=== bench artv fused" | tee -a $S
timeout 600 python bench.py --workload artv --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r1j_artv_fused.json 2> gpurun_out/r1j_artv_fused.err; echo "exit $?" | tee -a $S; cut -c1-300 gpurun_out/r1j_artv_fused.json | tee -a $S; tail -3 gpurun_out/r1j_artv_fused.err | tee -a $S
echo "=== bench tf32" | tee -a $S
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1j_bench_tf32.json 2> gpurun_out/r1j_bench_tf32.err; echo "exit $?" | tee -a $S; cut -c1-330 gpurun_out/r1j_bench_tf32.json | tee -a $S
python - <<'PY' | tee -a $S
import json
d=json.load(open('gpurun_out/r1j_bench_tf32.json')); print(d['roofline']['kernels'])
PY
