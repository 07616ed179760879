#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t1.log | cut -c1-300
timeout 900 python tools/rmat_sweep.py 2>&1 | tee gpurun_out/rmat_sweep.log
timeout 600 python tools/spmm_lab.py --d 256,960,2256 --impl ldg 2>&1 | tee gpurun_out/spmm_lab2.log
