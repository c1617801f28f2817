#!/bin/bash
# programmatic dependent launch across the forward (MMVID_PDL), GEMM pipeline cadence, full suite
source scripts/gpu_probe.sh
echo "=== gemm trace" | tee -a $S
timeout 300 python scripts/gemm_trace2.py fp16 2>&1 | tee -a $S
run full 2400 tests/ -x
MMVID_PDL=0 bench pdl0 --no-cpu-baseline
bench pdl1 --no-cpu-baseline
bench artv --no-cpu-baseline --workload artv --steps 2 --warmup 1
