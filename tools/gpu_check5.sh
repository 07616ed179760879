#!/bin/bash
# Final single-GPU pass of the round: tests, smoke, default bench, full-size ncu launch list and full captures
# of the unit-compacted SpMM / unit_pack / output-layer SpMM launches.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_products.log 2>gpurun_out/bench_products.err; echo "products default rc=$?"
head -c 300 gpurun_out/bench_products.log; echo
timeout 300 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --rhs-tile-gb 82 > gpurun_out/bench_products_g16.log 2>&1; echo "g16 rc=$?"; head -c 260 gpurun_out/bench_products_g16.log; echo
FULL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv $FULL > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_units_staged_kernel|unit_pack_kernel|spmm_vec_kernel<5" -s 4 -c 3 -o gpurun_out/prof_r1i $FULL > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
for w in arxiv pubmed cora; do timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>gpurun_out/bench_$w.err; echo "$w rc=$?"; head -c 200 gpurun_out/bench_$w.log; echo; done
