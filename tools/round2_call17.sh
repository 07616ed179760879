#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
run r2k_tests 900 python -m pytest tests -m gpu -q
tail -6 gpurun_out/r2k_tests.log | cut -c1-250
run r2k_bench 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-parity
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2k_bench.log").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["achieved"],1), round(t["issued_frac"],2)) for t in d["roofline"]["tensor_kernels"]])
    print([ (t["kernel"], round(t["avg_launch_ms"],2), round(t["frac"],2)) for t in d["roofline"]["spmm_groups"]])
except Exception as e:
    print("bench: no bench line:", e)
PY
tail -3 gpurun_out/r2k_bench.err | cut -c1-300
timeout 200 python bench.py --workload arxiv --steps 5 --warmup 3 --no-e2e > gpurun_out/r2k_bench_arxiv.log 2>gpurun_out/r2k_bench_arxiv.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2k_bench_arxiv.log").read().strip().splitlines()[-1])
    print("arxiv", round(d["value"]), "nodes/s", round(d["ms_per_step"], 2), "ms  parity", d["parity"]["ok"], d["parity"]["vs_oracle"])
except Exception as e:
    print("arxiv: no bench line:", e)
PY
timeout 200 python bench.py --workload arxiv --steps 2 --warmup 1 --no-e2e --no-fused-hess-spmm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv materialised rhs:', d['parity']['vs_oracle'])"
timeout 200 python bench.py --workload arxiv --steps 2 --warmup 1 --no-e2e --syrk simt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv simt syrk:', d['parity']['vs_oracle'])"
timeout 200 python bench.py --workload arxiv --steps 2 --warmup 1 --no-e2e --no-fused-gemm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv cublas gemm:', d['parity']['vs_oracle'])"
timeout 200 python bench.py --workload arxiv --steps 2 --warmup 1 --no-e2e --dense-slabs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv dense slabs:', d['parity']['vs_oracle'])"
