#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2l_smoke.txt | cut -c1-300
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_products.log 2>gpurun_out/r2l_bench_products.err; echo "products rc=$?"
for w in arxiv pubmed cora; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2l_bench_$w.log 2>gpurun_out/r2l_bench_$w.err; echo "$w rc=$?"
done
for w in products arxiv pubmed cora; do python - $w <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2l_bench_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print(sys.argv[1], round(d["value"]), "nodes/s", round(d["ms_per_step"], 2), "ms  marglik", d["marglik"], "parity", p.get("ok"), {k: v for k, v in (p.get("vs_oracle") or {}).items() if k != "per_block_rel"},
          "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("seconds"))
    if sys.argv[1] == "products":
        print("  ", d["roofline"]["ms_per_step_by_kind"], "frac", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"])
        print("  ", [(t["kernel"], round(t["avg_launch_ms"], 2), round(t["achieved"], 1), round(t["issued_frac"], 2)) for t in d["roofline"]["tensor_kernels"]])
        print("  ", [(t["kernel"], round(t["avg_launch_ms"], 2), round(t["frac"], 2)) for t in d["roofline"]["spmm_groups"]])
except Exception as e:
    print(sys.argv[1], "no bench line:", e)
PY
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
