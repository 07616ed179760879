#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_mask or fused" > gpurun_out/t_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -15 gpurun_out/t_gemm.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t1.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b1_fused.log 2>gpurun_out/b1_fused.err; echo "fused rc=$?"; tail -c 500 gpurun_out/b1_fused.log; tail -3 gpurun_out/b1_fused.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fused-gemm > gpurun_out/b1_nofused.log 2>gpurun_out/b1_nofused.err; echo "nofused rc=$?"; tail -c 500 gpurun_out/b1_nofused.log
