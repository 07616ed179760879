#!/bin/bash
# Multi-GPU follow-up of round 2 (gpurun --gpus N): the default command, then the switches of DESIGN.md §6c.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_multi_gpu.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out
P=29540
S=$(date +%s)
LGNN_LAB=1 timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_n${N}_tests.log 2>&1; echo "multi-GPU tests (incl. lab switches over NCCL) rc=$? in $(( $(date +%s) - S )) s"
tail -3 gpurun_out/r2_n${N}_tests.log | cut -c1-200
for flags in "" "--shard-eigh" "--unit-even-groups --shard-eigh" "--unit-even-groups --shard-eigh --fused-hess-spmm"; do
  tag=$(echo "base $flags" | tr -d '-' | tr ' ' '_')
  S=$(date +%s)
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-parity $flags > gpurun_out/r2_n${N}_$tag.log 2>gpurun_out/r2_n${N}_$tag.err
  echo "N=$N [$flags] rc=$? in $(( $(date +%s) - S )) s"
  python - "gpurun_out/r2_n${N}_$tag.log" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
except Exception as e:
    print("    no bench line:", e)
PY
  P=$((P + 1))
done
