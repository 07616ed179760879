#!/bin/bash
# Final multi-GPU pass of round 2: the driver's command at N ranks with the final code, then the deferred-gather A/B.
set -u
N=${1:-8}
mkdir -p gpurun_out
P=29740
bench() {  # tag, steps, warmup, extra flags...
  local tag=$1 steps=$2 warm=$3; shift 3
  local S=$(date +%s)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps $steps --warmup $warm "$@" > gpurun_out/r2f_n${N}_$tag.log 2>gpurun_out/r2f_n${N}_$tag.err
  echo "N=$N $tag [$*] rc=$? in $(( $(date +%s) - S )) s"
  python - "gpurun_out/r2f_n${N}_$tag.log" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"],
          "parity", (d.get("parity") or {}).get("ok"), "alloc", d.get("allocator_in_timed_region"), "e2e", e.get("value"), e.get("ms_per_step"))
except Exception as e:
    print("    no bench line:", e)
PY
  grep -vE "^\s*$|OMP_NUM|\*\*\*\*|Warning|sparse_csr|warn" "gpurun_out/r2f_n${N}_$tag.err" | tail -4 | cut -c1-300
  P=$((P + 1))
}
bench driver 20 5
bench no_defer 10 3 --no-e2e --no-parity --no-defer-gathers
bench defer 10 3 --no-e2e --no-parity
