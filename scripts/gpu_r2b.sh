#!/bin/bash
# r2b: CUDA-graphed forward in batched mask-predict: model tests, bench with / without graph
mkdir -p gpurun_out
S=gpurun_out/summary_r2b.txt; rm -f $S
echo "=== model + fullsize tests (graph on)" | tee -a $S
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300 | tee -a $S
for gsw in 0 1; do
echo "=== bench tf32 graph=$gsw" | tee -a $S
MMVID_CUDA_GRAPH=$gsw timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_g$gsw.json 2> gpurun_out/r2b_bench_g$gsw.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r2b_bench_g$gsw.json | tee -a $S; tail -2 gpurun_out/r2b_bench_g$gsw.err | tee -a $S
done
echo "=== bench bf16 graph=1" | tee -a $S
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/r2b_bench_bf16.json 2> gpurun_out/r2b_bench_bf16.err; echo "exit $?" | tee -a $S; cut -c1-260 gpurun_out/r2b_bench_bf16.json | tee -a $S
python - <<'PY' | tee -a $S
import json
for f in ("r2b_bench_g0","r2b_bench_g1","r2b_bench_bf16"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["clocks"])
    except Exception as e: print(f, "parse failed", e)
PY
