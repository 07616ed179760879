#!/bin/bash
# First GPU call of round 2 (one B200): everything that was built after round 1's GPU budget ended
# gets its first run — lab tests, kernel labs, then the bench with each switch — so that the defaults can be flipped
# on evidence.  Each step is bounded by its own timeout; logs land in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run tf32_peak 60 python tools/tf32_peak.py
cat gpurun_out/tf32_peak.log
run r2_tests_default 600 python -m pytest tests -m gpu -q -x
tail -2 gpurun_out/r2_tests_default.log | cut -c1-200
LGNN_LAB=1 run r2_tests_lab 600 python -m pytest tests/test_gpu_lab.py -q
tail -15 gpurun_out/r2_tests_lab.log | cut -c1-220
run r2_units_lab 300 python tools/units_lab.py 6 8 10 12
cat gpurun_out/r2_units_lab.log | cut -c1-260
run r2_gemm_lab 200 python tools/gemm_lab.py
tail -9 gpurun_out/r2_gemm_lab.log | cut -c1-200
run r2_syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
tail -9 gpurun_out/r2_syrk_lab.log | cut -c1-200
run r2_hess_spmm_lab 200 python tools/hess_spmm_lab.py 16 8 6
cat gpurun_out/r2_hess_spmm_lab.log | cut -c1-260
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run r2_bench_base 300 $B
run r2_bench_hess 300 $B --fused-hess-spmm
run r2_bench_stack 300 $B --syrk-stack-narrow
run r2_bench_overlap 300 $B --overlap-groups
for f in r2_bench_base r2_bench_hess r2_bench_stack r2_bench_overlap; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
except Exception as e:
    print(sys.argv[1], "no bench line:", e)
PY
done
run r2_rmat_hub 500 python tools/rmat_sweep.py --scales 20,22 --degrees 16,64 --hub-split
cat gpurun_out/r2_rmat_hub.log | cut -c1-400
