#!/bin/bash
# One GPU-box pass: parity tests, kernel labs, bench lines, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log
timeout 600 python tools/spmm_lab.py > gpurun_out/spmm_lab.log 2>&1; echo "spmm_lab rc=$?"; cat gpurun_out/spmm_lab.log | tail -12
timeout 600 python tools/syrk_lab.py > gpurun_out/syrk_lab.log 2>&1; echo "syrk_lab rc=$?"
LGNN_SYRK_SEG_STEPS=64 timeout 300 python tools/syrk_lab.py --n 256 --impl tcgen05 >> gpurun_out/syrk_lab.log 2>&1
LGNN_SYRK_SEG_STEPS=1000000 timeout 300 python tools/syrk_lab.py --n 256 --impl tcgen05 >> gpurun_out/syrk_lab.log 2>&1
cat gpurun_out/syrk_lab.log | tail -24
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_products.log 2>gpurun_out/bench_products.err; echo "products rc=$?"
python bench.py --workload arxiv --steps 3 --warmup 2 > gpurun_out/bench_arxiv.log 2>gpurun_out/bench_arxiv.err; echo "arxiv rc=$?"
CMD="python bench.py --workload products --scale 0.125 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_bulk|syrk_tcgen05_kernel" -s 6 -c 4 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
tail -c 700 gpurun_out/bench_products.log
