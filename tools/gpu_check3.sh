#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed" > gpurun_out/t_pack.log 2>&1; echo "packed tests rc=$?"; tail -12 gpurun_out/t_pack.log | cut -c1-250
timeout 300 python tools/pack_lab.py 2>&1 | tee gpurun_out/pack_lab.log | tail -5
