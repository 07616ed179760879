#!/bin/bash
# Round 2, second single-GPU call: the two-blocks-per-warp unit SpMM (g = 6), the on-the-fly output-layer SpMM
# (lab tests + lab + bench), and one `ncu --set full` each of the fused GEMM and the tcgen05 SYRK (K = n = 256).
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2b_tests_even 400 python -m pytest tests/test_gpu_units_even.py -q -x
tail -3 gpurun_out/r2b_tests_even.log | cut -c1-200
LGNN_LAB=1 run r2b_tests_lab 400 python -m pytest tests/test_gpu_lab.py -q
tail -8 gpurun_out/r2b_tests_lab.log | cut -c1-220
run r2b_units_lab 300 python tools/units_lab.py 6
cat gpurun_out/r2b_units_lab.log | cut -c1-260
run r2b_hess_spmm_lab 200 python tools/hess_spmm_lab.py 16 8
cat gpurun_out/r2b_hess_spmm_lab.log | cut -c1-300; tail -3 gpurun_out/r2b_hess_spmm_lab.err | cut -c1-300
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-parity"
run r2b_bench_hess 300 $B --fused-hess-spmm
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench_hess.log").read().strip().splitlines()[-1])
    print("bench_hess", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
except Exception as e:
    print("bench_hess: no bench line:", e)
PY
tail -3 gpurun_out/r2b_bench_hess.err | cut -c1-300
# ncu captures (the third launch of the product kernel in each lab = first shape, K = n = 256)
run r2b_ncu_gemm 400 ncu --set full --import-source on --clock-control none -k regex:^gemm_mask_kernel -s 2 -c 1 -f -o gpurun_out/r2b_gemm_mask python tools/gemm_lab.py
run r2b_ncu_syrk 400 ncu --set full --import-source on --clock-control none -k regex:^syrk_tcgen05_kernel -s 1 -c 1 -f -o gpurun_out/r2b_syrk python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
ls -la gpurun_out/*.ncu-rep
