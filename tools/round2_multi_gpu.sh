#!/bin/bash
# Multi-GPU pass of round 2 (gpurun --gpus N): NCCL parity tests (incl. the lab switches), then the DRIVER's bench
# command at N ranks (--steps 20 --warmup 5: the run that ran out of memory at N = 2 / 4 in round 1), then switches.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_multi_gpu.sh 2'
set -u
N=${1:-8}
mkdir -p gpurun_out
P=29540
S=$(date +%s)
LGNN_LAB=1 timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_n${N}_tests.log 2>&1; echo "multi-GPU tests (incl. lab switches over NCCL) rc=$? in $(( $(date +%s) - S )) s"
tail -3 gpurun_out/r2_n${N}_tests.log | cut -c1-200
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
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"],
          "parity", d.get("parity"), "alloc", d.get("allocator_in_timed_region"), "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("    no bench line:", e)
PY
  tail -4 "gpurun_out/r2_n${N}_$tag.err" | cut -c1-300
  P=$((P + 1))
}
bench driver 20 5
bench shard_eigh 5 3 --no-e2e --no-parity --shard-eigh
bench pad4 5 3 --no-e2e --no-parity --no-unit-even-groups
bench rows 5 3 --no-e2e --no-parity --backward-parallel rows
# single-GPU extras riding on this call: the unicast variant of the fused GEMM against the multicast one
if [ "${EXTRA_GEMM_LAB:-0}" = "1" ]; then
  timeout 200 python tools/gemm_lab.py > gpurun_out/r2_gemm_lab_mc.log 2>&1; head -4 gpurun_out/r2_gemm_lab_mc.log | cut -c1-200
  LGNN_GEMM_UNICAST=1 timeout 200 python tools/gemm_lab.py > gpurun_out/r2_gemm_lab_uc.log 2>&1; head -4 gpurun_out/r2_gemm_lab_uc.log | cut -c1-200
  LGNN_GEMM_UNICAST=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm" > gpurun_out/r2_gemm_uc_tests.log 2>&1; tail -2 gpurun_out/r2_gemm_uc_tests.log | cut -c1-200
fi
