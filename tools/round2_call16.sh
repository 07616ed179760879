#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2j_tests 900 python -m pytest tests -m gpu -q
tail -6 gpurun_out/r2j_tests.log | cut -c1-250
run r2j_bench 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-parity
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2j_bench.log").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1), round(t["issued_frac"],2)) for t in d["roofline"]["tensor_kernels"]])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["frac"],2)) for t in d["roofline"]["spmm_groups"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
tail -3 gpurun_out/r2j_bench.err | cut -c1-300
for w in arxiv pubmed cora; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r2j_bench_$w.log 2>gpurun_out/r2j_bench_$w.err
  python - $w <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2j_bench_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), "nodes/s", round(d["ms_per_step"], 2), "ms  marglik", d["marglik"], "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("vs_oracle"), "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "no bench line:", e)
PY
  tail -2 gpurun_out/r2j_bench_$w.err | cut -c1-200
done
for w in cora pubmed; do
  timeout 300 python bench.py --impl reference --workload $w --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-700
done
