#!/bin/bash
source scripts/gpu_probe.sh
run ids 600 tests/test_gpu_0_models.py -k "ids_bit_exact_vs_oracle"
