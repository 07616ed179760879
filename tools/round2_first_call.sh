#!/bin/bash
# First GPU call of round 2 (one B200): the driver's own bench command (leak fix, parity block), the default and the
# LGNN_LAB test suites, the kernel labs, then the bench with each lab switch, then the DRAM-traffic capture of the
# dominant kernel at the bench shape.  Each step is bounded by its own timeout; logs land in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run tf32_peak 60 python tools/tf32_peak.py
cat gpurun_out/tf32_peak.log
run r2a_bench_driver 400 python bench.py --steps 20 --warmup 5
tail -c 1500 gpurun_out/r2a_bench_driver.err
run r2a_tests_default 700 python -m pytest tests -m gpu -q -x
tail -2 gpurun_out/r2a_tests_default.log | cut -c1-200
LGNN_LAB=1 run r2a_tests_lab 600 python -m pytest tests/test_gpu_lab.py -q
tail -15 gpurun_out/r2a_tests_lab.log | cut -c1-220
run r2a_units_lab 300 python tools/units_lab.py 6 8 10 12 16
cat gpurun_out/r2a_units_lab.log | cut -c1-260
run r2a_gemm_lab 200 python tools/gemm_lab.py
tail -12 gpurun_out/r2a_gemm_lab.log | cut -c1-200
run r2a_syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
tail -12 gpurun_out/r2a_syrk_lab.log | cut -c1-200
run r2a_hess_spmm_lab 200 python tools/hess_spmm_lab.py 16 8 6
cat gpurun_out/r2a_hess_spmm_lab.log | cut -c1-260
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-parity"
run r2a_bench_base 300 $B
run r2a_bench_hess 300 $B --fused-hess-spmm
run r2a_bench_stack 300 $B --syrk-stack-narrow
run r2a_bench_overlap 300 $B --overlap-groups
run r2a_bench_hess_stack 300 $B --fused-hess-spmm --syrk-stack-narrow
for f in r2a_bench_driver r2a_bench_base r2a_bench_hess r2a_bench_stack r2a_bench_overlap r2a_bench_hess_stack; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"],
          "parity", (d.get("parity") or {}).get("ok"), "alloc", d.get("allocator_in_timed_region"), "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "no bench line:", e)
PY
done
# DRAM traffic of the dominant kernel at the bench shape (ncu replays only the filtered kernel)
run r2a_traffic 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:spmm_units --csv --log-file gpurun_out/r2a_traffic_spmm_units.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-parity
tail -3 gpurun_out/r2a_traffic.err | cut -c1-300
