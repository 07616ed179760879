#!/bin/bash
# Round 2 profiling pass (one B200): ncu launch list of the bench command (kernel shares), one --set full capture of the
# dominant kernel at the bench shape, CPU port at 1/4 scale, R-MAT 2^24.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-parity > gpurun_out/r2m_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmm_units_staged -s 6 -c 1 -f -o gpurun_out/r2m_spmm_units \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-parity > gpurun_out/r2m_ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/r2m_spmm_units.ncu-rep
timeout 900 python tools/rmat_sweep.py --scales 24 --degrees 16 > gpurun_out/r2m_rmat24.log 2>gpurun_out/r2m_rmat24.err; echo "rmat 2^24 rc=$?"
cat gpurun_out/r2m_rmat24.log | cut -c1-600; tail -2 gpurun_out/r2m_rmat24.err | cut -c1-300
timeout 900 python bench.py --impl reference --cpu-sample-div 4 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
timeout 900 python bench.py --impl reference --cpu-sample-div 16 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
