#!/bin/bash
# One B200: everything the round-2 numbers in DESIGN.md come from, in one command (each step bounded by its own timeout;
# logs land in gpurun_out/, summaries are copied to profiles/ by hand):
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/round2_single_gpu.sh [tests|bench|labs|profile|rmat|all]'
set -u
WHAT=${1:-all}
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; local S=$(date +%s); timeout "$t" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$? in $(( $(date +%s) - S )) s"; }
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.log").read().strip().splitlines()[-1])
    p = d.get("parity") or {}
    print("  ", sys.argv[1], round(d["value"]), "nodes/s", round(d["ms_per_step"], 2), "ms  marglik", d["marglik"], "parity", p.get("ok"),
          "e2e", (d.get("e2e") or {}).get("value"), d["roofline"]["ms_per_step_by_kind"])
    print("  ", [(t["kernel"], round(t["avg_launch_ms"], 2), round(t["achieved"], 1), round(t["issued_frac"], 2)) for t in d["roofline"]["tensor_kernels"]],
          [(t["kernel"], round(t["avg_launch_ms"], 2), round(t["frac"], 2)) for t in d["roofline"]["spmm_groups"]])
except Exception as e:
    print("  ", sys.argv[1], "no bench line:", e)
PY
}
if [ $WHAT = tests ] || [ $WHAT = all ]; then
  run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"; tail -1 gpurun_out/smoke.log | cut -c1-300
  run tests_gpu 900 python -m pytest tests -m gpu -q; tail -3 gpurun_out/tests_gpu.log | cut -c1-250
fi
if [ $WHAT = bench ] || [ $WHAT = all ]; then
  run bench_products 500 python bench.py --steps 20 --warmup 5; line bench_products
  for w in arxiv pubmed cora; do run bench_$w 300 python bench.py --workload $w --steps 20 --warmup 5; line bench_$w; done
  run bench_reference 400 python bench.py --impl reference --steps 2 --warmup 1; tail -1 gpurun_out/bench_reference.log | cut -c1-400
  for w in cora pubmed; do run bench_reference_$w 300 python bench.py --impl reference --workload $w --steps 2 --warmup 1; tail -1 gpurun_out/bench_reference_$w.log | cut -c1-300; done
fi
if [ $WHAT = labs ] || [ $WHAT = all ]; then
  run tf32_peak 60 python tools/tf32_peak.py; cat gpurun_out/tf32_peak.log
  run units_lab 300 python tools/units_lab.py 6 12 16; cat gpurun_out/units_lab.log | cut -c1-230
  run hess_spmm_lab 200 python tools/hess_spmm_lab.py 16 8; cat gpurun_out/hess_spmm_lab.log | cut -c1-300
  run gemm_lab 200 python tools/gemm_lab.py; tail -13 gpurun_out/gemm_lab.log | cut -c1-200
  run syrk_lab 200 python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05; tail -9 gpurun_out/syrk_lab.log | cut -c1-200
  run gemm_fit_diag 200 python tools/gemm_fit_diag.py; grep gemm gpurun_out/gemm_fit_diag.log | head -6 | cut -c1-300
fi
if [ $WHAT = profile ] || [ $WHAT = all ]; then
  B="python bench.py --steps 1 --warmup 1 --no-e2e --no-parity"
  run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B
  run ncu_traffic 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm_units --csv --log-file gpurun_out/traffic_spmm_units.csv $B
  run ncu_full_units 600 ncu --set full --import-source on --clock-control none -k regex:spmm_units_staged -s 6 -c 1 -f -o gpurun_out/spmm_units $B
  run ncu_full_gemm 400 ncu --set full --import-source on --clock-control none -k regex:^gemm_mask_kernel -s 2 -c 1 -f -o gpurun_out/gemm_mask python tools/gemm_lab.py
  run ncu_full_syrk 400 ncu --set full --import-source on --clock-control none -k regex:^syrk_tcgen05_kernel -s 1 -c 1 -f -o gpurun_out/syrk python tools/syrk_lab.py --n 256 --dist randn --impl tcgen05
  ls -la gpurun_out/*.ncu-rep
fi
if [ $WHAT = rmat ] || [ $WHAT = all ]; then
  run rmat 1500 python tools/rmat_sweep.py --scales 20,22 --degrees 16,64; cat gpurun_out/rmat.log | cut -c1-600
  run rmat24 900 python tools/rmat_sweep.py --scales 24 --degrees 16; cat gpurun_out/rmat24.log | cut -c1-600
fi
