#!/bin/bash
# r1k: rotating-score-buffer attention (impl 3) variants, GEMM polling waits A/B, full GPU suite in ONE process, bench
mkdir -p gpurun_out
S=gpurun_out/summary_r1k.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv | tee -a $S
echo "=== attention variants" | tee -a $S
timeout 600 python scripts/att_bench.py 2>&1 | grep "^ATT" | cut -c1-700 | tee -a $S
echo "=== gemm spin A/B" | tee -a $S
for prec in tf32 bf16; do for sp in 0 1; do
  echo "--- $prec spin=$sp" | tee -a $S
  MMVID_GEMM_SPIN=$sp timeout 200 python scripts/gemm_sweep.py --child $prec 2>&1 | tail -1 | cut -c1-600 | tee -a $S
done; done
echo "=== attention tests with impl 3 (poly 2)" | tee -a $S
MMVID_ATT_IMPL=3 MMVID_ATT_POLY=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "attention" > gpurun_out/r1k_att3.log 2>&1; echo "exit $?" | tee -a $S; tail -n 5 gpurun_out/r1k_att3.log | cut -c1-300 | tee -a $S
echo "=== model tests with impl 3 (poly 2)" | tee -a $S
MMVID_ATT_IMPL=3 MMVID_ATT_POLY=2 timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider -k "tf32 or bf16 or batched" > gpurun_out/r1k_models3.log 2>&1; echo "exit $?" | tee -a $S; tail -n 5 gpurun_out/r1k_models3.log | cut -c1-300 | tee -a $S
echo "=== bench tf32 impl 2 (baseline)" | tee -a $S
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench_impl2.json 2> gpurun_out/r1k_bench_impl2.err; echo "exit $?" | tee -a $S; cut -c1-330 gpurun_out/r1k_bench_impl2.json | tee -a $S
echo "=== bench tf32 impl 3 poly 2" | tee -a $S
MMVID_ATT_IMPL=3 MMVID_ATT_POLY=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench_impl3.json 2> gpurun_out/r1k_bench_impl3.err; echo "exit $?" | tee -a $S; cut -c1-330 gpurun_out/r1k_bench_impl3.json | tee -a $S
echo "=== bench tf32 impl 3 poly 2 + gemm spin" | tee -a $S
MMVID_GEMM_SPIN=1 MMVID_ATT_IMPL=3 MMVID_ATT_POLY=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench_impl3s.json 2> gpurun_out/r1k_bench_impl3s.err; echo "exit $?" | tee -a $S; cut -c1-330 gpurun_out/r1k_bench_impl3s.json | tee -a $S
echo "=== full GPU suite, one process (as the driver runs it)" | tee -a $S
/usr/bin/time -v timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r1k_full.log 2>&1; echo "exit $?" | tee -a $S; grep -E "passed|failed|error|Elapsed" gpurun_out/r1k_full.log | tail -5 | cut -c1-300 | tee -a $S
echo "=== smoke" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee -a $S
