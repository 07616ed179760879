#!/bin/bash
# Round 2, single-GPU call: whole default GPU suite (reference-driven drop-in on the device, epoch-loop and kNN goldens),
# the tcgen05 kernels with 32-bit incremental loop state / dynamic tiles per cluster, bench.
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2e_tests 900 python -m pytest tests -m gpu -q
tail -6 gpurun_out/r2e_tests.log | cut -c1-250
run r2e_gemm_lab 200 python tools/gemm_lab.py
tail -12 gpurun_out/r2e_gemm_lab.log | cut -c1-200
for tpc in 32 256 4096; do
  LGNN_GEMM_TILES_PER_CLUSTER=$tpc timeout 100 python tools/gemm_lab.py 2>/dev/null | head -3 | sed "s/^/tpc=$tpc  /" | cut -c1-150
done
run r2e_syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
tail -9 gpurun_out/r2e_syrk_lab.log | cut -c1-200
run r2e_bench 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-parity
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2e_bench.log").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1)) for t in d["roofline"]["tensor_kernels"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
tail -3 gpurun_out/r2e_bench.err | cut -c1-300
