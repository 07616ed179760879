#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t1.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b1.log 2>gpurun_out/b1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/b1.log; tail -2 gpurun_out/b1.err
