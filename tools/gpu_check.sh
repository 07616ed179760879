#!/bin/bash
# Final single-GPU pass: tests, smoke, default bench + reference arm, ncu launch list, DRAM traffic of the
# full-size SpMM launches, full captures of the three hot kernels.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
FULL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$FULL > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm_bulk --csv --log-file gpurun_out/traffic.csv $FULL > gpurun_out/ncu0.log 2>&1; echo "ncu traffic rc=$?"
CMD="python bench.py --workload products --scale 0.125 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_bulk_kernel|syrk_tcgen05_kernel|gemm_mask_kernel" -s 3 -c 6 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
python bench.py > gpurun_out/bench_products.log 2>gpurun_out/bench_products.err; echo "products default rc=$?"
tail -c 2500 gpurun_out/bench_products.log
