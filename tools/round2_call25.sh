#!/bin/bash
# 2-GPU pass: NCCL tests, driver command with the deferred activation all-gathers; GPU 0 alone: even-g unit SpMM variants
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_units_even.py -q -x > gpurun_out/r2n_tests_even.log 2>&1; echo "even tests rc=$?"; tail -2 gpurun_out/r2n_tests_even.log | cut -c1-200
timeout 300 python tools/units_lab.py 6 16 12 > gpurun_out/r2n_units_lab.log 2>&1; cat gpurun_out/r2n_units_lab.log | cut -c1-230
bash tools/round2_multi_gpu.sh 2 short
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2n_n2_tests.log 2>&1; echo "multi-GPU tests rc=$?"; tail -2 gpurun_out/r2n_n2_tests.log | cut -c1-200
