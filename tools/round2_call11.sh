#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
for cfg in "256 64 2000000 20" "64 64 19999992 8" "64 64 2000000 20" "256 128 2000000 20" "47 256 19999992 6" "256 256 19999992 6"; do
  timeout 200 python tools/gemm_repro.py $cfg 2>&1 | grep -E "ok, rel|Error" | head -1 | cut -c1-160
  echo "   [$cfg] rc=${PIPESTATUS[0]}"
done
run r2g_tests 900 python -m pytest tests -m gpu -q
tail -4 gpurun_out/r2g_tests.log | cut -c1-250
run r2g_gemm_lab 200 python tools/gemm_lab.py
tail -13 gpurun_out/r2g_gemm_lab.log | cut -c1-200
run r2g_bench 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-parity
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2g_bench.log").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1)) for t in d["roofline"]["tensor_kernels"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
tail -3 gpurun_out/r2g_bench.err | cut -c1-300
