#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary_quick.txt; rm -f $S
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a $S
  timeout $to python -m pytest "$@" -m gpu -q -s -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a $S; tail -n 3 gpurun_out/$name.log | tee -a $S; }
run q_att 300 tests/test_gpu_kernels.py -k "attention"
run q_artv 600 tests/test_gpu_models.py -k "artv or transformer"
for w in "bert tf32" "bert bf16" "artv tf32"; do
  set -- $w
  echo "=== bench $1 $2" | tee -a $S
  timeout 900 python bench.py --workload $1 --steps 2 --warmup 3 --precision $2 --no-cpu-baseline > gpurun_out/q_bench_$1_$2.json 2> gpurun_out/q_bench_$1_$2.err; echo "exit $?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d=json.load(open("gpurun_out/q_bench_$1_$2.json"))
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "launches", d["gpu_launches"], {k:(round(v["ms"],4), round(v["tflops"],1)) for k,v in d["roofline"]["kernels"].items()})
except Exception as ex:
    print("parse failed", ex); print(open("gpurun_out/q_bench_$1_$2.err").read()[-1200:])
PY
done
