#!/bin/bash
# quick iteration: tensor-core kernel tests, model parity (tc), short bench
mkdir -p gpurun_out
S=gpurun_out/summary_iter.txt; rm -f $S
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a $S
  timeout $to python -m pytest "$@" -m gpu -q -s -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a $S; tail -n 4 gpurun_out/$name.log | tee -a $S; }
run it_k_lin 300 tests/test_gpu_kernels.py -k "linear or conv3x3_tensor_core or conv_out_fused or groupnorm or fused_qkv"
run it_k_att 300 tests/test_gpu_kernels.py -k "attention and (tf32 or bf16)"
run it_models 1200 tests/test_gpu_models.py -k "tf32 or bf16 or batched or tensor_core or vae or kv_cache"
run it_train 900 tests/test_gpu_training.py
for prec in tf32 bf16; do
  echo "=== bench $prec" | tee -a $S
  timeout 600 python bench.py --steps 3 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/it_bench_$prec.json 2> gpurun_out/it_bench_$prec.err; echo "exit $?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d=json.load(open("gpurun_out/it_bench_$prec.json"))
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    print({k:(round(v["ms"],4), round(v["tflops"],1)) for k,v in d["roofline"]["kernels"].items()})
except Exception as ex:
    print("bench parse failed", ex); print(open("gpurun_out/it_bench_$prec.err").read()[-1500:])
PY
done
echo "=== bench artv tf32" | tee -a $S
timeout 900 python bench.py --workload artv --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/it_bench_artv.json 2> gpurun_out/it_bench_artv.err; echo "exit $?" | tee -a $S
python - <<PY | tee -a $S
import json
try:
    d=json.load(open("gpurun_out/it_bench_artv.json"))
    print("artv value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "launches", d["gpu_launches"])
except Exception as ex:
    print("artv bench parse failed", ex); print(open("gpurun_out/it_bench_artv.err").read()[-1500:])
PY
