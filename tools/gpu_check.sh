#!/bin/bash
# One GPU-box pass: parity tests, the default bench line, ncu launch list + full captures of the three hot kernels.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
/usr/bin/time -v python bench.py > gpurun_out/bench_products.log 2>gpurun_out/bench_products.err; echo "products default rc=$?"; grep -E "Elapsed" gpurun_out/bench_products.err
/usr/bin/time -v python bench.py --impl reference > gpurun_out/bench_reference.log 2>gpurun_out/bench_reference.err; echo "reference rc=$?"; grep -E "Elapsed" gpurun_out/bench_reference.err
python bench.py --workload arxiv --steps 5 --warmup 3 > gpurun_out/bench_arxiv.log 2>gpurun_out/bench_arxiv.err; echo "arxiv rc=$?"
python bench.py --workload pubmed --steps 5 --warmup 3 > gpurun_out/bench_pubmed.log 2>gpurun_out/bench_pubmed.err; echo "pubmed rc=$?"
python bench.py --workload cora --steps 5 --warmup 3 > gpurun_out/bench_cora.log 2>gpurun_out/bench_cora.err; echo "cora rc=$?"
CMD="python bench.py --workload products --scale 0.125 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_bulk|syrk_tcgen05_kernel|gemm_mask_kernel" -s 8 -c 5 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
tail -c 1500 gpurun_out/bench_products.log
