#!/bin/bash
# Round 2, third single-GPU call: warp-collective elected MMA / TMA issue in the two tcgen05 kernels.
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2c_tests 700 python -m pytest tests -m gpu -q -x
tail -3 gpurun_out/r2c_tests.log | cut -c1-200
run r2c_gemm_lab 200 python tools/gemm_lab.py
tail -12 gpurun_out/r2c_gemm_lab.log | cut -c1-200
run r2c_syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
tail -12 gpurun_out/r2c_syrk_lab.log | cut -c1-200
run r2c_syrk_lab48 200 python tools/syrk_lab.py --n 47 --dist randn --impl tcgen05
tail -3 gpurun_out/r2c_syrk_lab48.log | cut -c1-200
run r2c_bench 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-parity
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench.log").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1)) for t in d["roofline"]["tensor_kernels"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
tail -3 gpurun_out/r2c_bench.err | cut -c1-300
run r2c_ncu_gemm 400 ncu --set full --import-source on --clock-control none -k regex:^gemm_mask_kernel -s 2 -c 1 -f -o gpurun_out/r2c_gemm_mask python tools/gemm_lab.py
run r2c_ncu_syrk 400 ncu --set full --import-source on --clock-control none -k regex:^syrk_tcgen05_kernel -s 1 -c 1 -f -o gpurun_out/r2c_syrk python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
