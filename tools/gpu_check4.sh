#!/bin/bash
# Single-GPU pass after the unit-compacted slabs: tests, smoke, default bench + dense-slab bench, ncu launch
# list, DRAM traffic of the full-size unit SpMM launches, full captures of the hot kernels.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_products.log 2>gpurun_out/bench_products.err; echo "products default rc=$?"
head -c 300 gpurun_out/bench_products.log; echo
FULL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm_units_staged --csv --log-file gpurun_out/traffic_units.csv $FULL > gpurun_out/ncu0.log 2>&1; echo "ncu traffic rc=$?"
CMD="python bench.py --workload products --scale 0.125 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_units_staged_kernel|unit_pack_kernel|spmm_vec_kernel" -s 6 -c 8 -o gpurun_out/prof_r1h $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
