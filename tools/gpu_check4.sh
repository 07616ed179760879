#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t1.log | cut -c1-200
S=$(date +%s); python bench.py > gpurun_out/bench_products.log 2>gpurun_out/bench_products.err; echo "products default rc=$? in $(( $(date +%s) - S )) s"
S=$(date +%s); python bench.py --impl reference > gpurun_out/bench_reference.log 2>gpurun_out/bench_reference.err; echo "reference rc=$? in $(( $(date +%s) - S )) s"
tail -c 1800 gpurun_out/bench_products.log; echo; tail -c 600 gpurun_out/bench_reference.log
