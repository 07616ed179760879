#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log | cut -c1-250
FULL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$FULL > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"spmm_bulk|spmm_vec" --csv --log-file gpurun_out/traffic.csv $FULL > gpurun_out/ncu0.log 2>&1; echo "ncu traffic rc=$?"
tail -c 900 gpurun_out/plain_full.log
