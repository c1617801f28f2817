#!/bin/bash
source scripts/gpu_probe.sh
run clip 600 tests/test_gpu_5_clip.py
