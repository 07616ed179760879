#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2i_tests 900 python -m pytest tests -m gpu -q
tail -4 gpurun_out/r2i_tests.log | cut -c1-250
run r2i_syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
tail -9 gpurun_out/r2i_syrk_lab.log | cut -c1-200
for n in 47 100 128 200; do timeout 100 python tools/syrk_lab.py --n $n --dist randn --impl tcgen05 2>/dev/null | head -1 | cut -c1-200; done
run r2i_bench 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-parity
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2i_bench.log").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1)) for t in d["roofline"]["tensor_kernels"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
run r2i_ncu_syrk 400 ncu --set full --import-source on --clock-control none -k regex:^syrk_tcgen05_kernel -s 1 -c 1 -f -o gpurun_out/r2i_syrk python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
