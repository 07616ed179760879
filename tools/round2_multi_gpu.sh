#!/bin/bash
# Multi-GPU pass of round 2 (gpurun --gpus N): NCCL parity tests, then the DRIVER's bench command at N ranks
# (--steps 20 --warmup 5: the run that ran out of memory at N = 2 / 4 in round 1), then single switches.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_multi_gpu.sh 8 full'
set -u
N=${1:-8}
MODE=${2:-full}
mkdir -p gpurun_out
P=29540
if [ "$MODE" = full ]; then
  S=$(date +%s)
  timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_n${N}_tests.log 2>&1; echo "multi-GPU tests rc=$? in $(( $(date +%s) - S )) s"
  tail -3 gpurun_out/r2_n${N}_tests.log | cut -c1-200
fi
bench() {  # tag, steps, warmup, extra flags...
  local tag=$1 steps=$2 warm=$3; shift 3
  local S=$(date +%s)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps $steps --warmup $warm "$@" > gpurun_out/r2_n${N}_$tag.log 2>gpurun_out/r2_n${N}_$tag.err
  echo "N=$N $tag [$*] rc=$? in $(( $(date +%s) - S )) s"
  python - "gpurun_out/r2_n${N}_$tag.log" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"],
          "parity", (d.get("parity") or {}).get("ok"), "alloc", d.get("allocator_in_timed_region"), "e2e", e.get("value"), e.get("ms_per_step"))
except Exception as e:
    print("    no bench line:", e)
PY
  grep -vE "^\s*$|OMP_NUM|\*\*\*\*|Warning|sparse_csr|warn" "gpurun_out/r2_n${N}_$tag.err" | tail -4 | cut -c1-300
  P=$((P + 1))
}
bench driver 20 5
if [ "$MODE" = full ]; then
  bench pad4 5 3 --no-e2e --no-parity --no-unit-even-groups
  bench replicated_eigh 5 3 --no-e2e --no-parity --no-shard-eigh
fi
