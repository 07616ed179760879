#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2h_tests_syrk 600 python -m pytest tests/test_gpu_syrk_tcgen05.py tests/test_gpu_parity.py -m gpu -q -x
tail -4 gpurun_out/r2h_tests_syrk.log | cut -c1-250
run r2h_syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
tail -9 gpurun_out/r2h_syrk_lab.log | cut -c1-200
run r2h_syrk_lab47 200 python tools/syrk_lab.py --n 47 --dist randn --impl tcgen05
head -2 gpurun_out/r2h_syrk_lab47.log | cut -c1-200
run r2h_gemm_lab 200 python tools/gemm_lab.py
head -5 gpurun_out/r2h_gemm_lab.log | cut -c1-200
LGNN_LIB_PATH=$PWD/laplace_gnn_b200/liblgnn_epi.so run r2h_gemm_lab_epi 200 python tools/gemm_lab.py
head -5 gpurun_out/r2h_gemm_lab_epi.log | cut -c1-200
LGNN_LIB_PATH=$PWD/laplace_gnn_b200/liblgnn_epi.so run r2h_tests_epi 400 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_kernel_shape_golden.py -m gpu -q -x -k "gemm or kernel_shapes or golden"
tail -3 gpurun_out/r2h_tests_epi.log | cut -c1-250
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-parity"
run r2h_bench 300 $B
LGNN_LIB_PATH=$PWD/laplace_gnn_b200/liblgnn_epi.so run r2h_bench_epi 300 $B
for f in r2h_bench r2h_bench_epi; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1)) for t in d["roofline"]["tensor_kernels"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
done
run r2h_rmat 900 python tools/rmat_sweep.py --scales 20,22 --degrees 16,64
cat gpurun_out/r2h_rmat.log | cut -c1-700; tail -3 gpurun_out/r2h_rmat.err | cut -c1-300
