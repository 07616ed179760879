#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_syrk_tcgen05.py -m gpu -q -x > gpurun_out/t_syrk.log 2>&1; echo "syrk tests rc=$?"; tail -8 gpurun_out/t_syrk.log | cut -c1-250
timeout 300 python tools/syrk_lab.py --n 256,48,100,128 --impl tcgen05 --k 4000000 2>&1 | tee gpurun_out/syrk_lab.log | tail -10
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spmm" > gpurun_out/t_spmm.log 2>&1; echo "spmm tests rc=$?"; tail -3 gpurun_out/t_spmm.log | cut -c1-250
timeout 600 python tools/spmm_lab.py --d 3840,4096,3072,2560 --impl ldg,bulk 2>&1 | tee gpurun_out/spmm_lab3.log | tail -10
