#!/bin/bash
# 2-GPU diagnostic: the row-partitioned backward at the products shape (launch failure in round2_multi_gpu.sh),
# the NCCL parity tests, and the driver's command with the sharded ingest.
set -u
N=${1:-2}
mkdir -p gpurun_out
P=29640
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2d_n${N}_tests.log 2>&1; echo "multi-GPU tests rc=$?"
tail -3 gpurun_out/r2d_n${N}_tests.log | cut -c1-200
for tag in rows_noov rows_ov; do
  fl="--backward-parallel rows"; [ $tag = rows_noov ] && fl="$fl --no-overlap"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-parity $fl > gpurun_out/r2d_n${N}_$tag.log 2>gpurun_out/r2d_n${N}_$tag.err
  echo "$tag rc=$?"; P=$((P + 1))
  grep -E "LgnnError|Error|File \"/.*laplace_gnn_b200|File \"/.*bench.py" gpurun_out/r2d_n${N}_$tag.err | head -12 | cut -c1-260
  python - "gpurun_out/r2d_n${N}_$tag.log" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], d["roofline"]["ms_per_step_by_kind"])
except Exception as e:
    print("    no bench line:", e)
PY
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2d_n${N}_driver.log 2>gpurun_out/r2d_n${N}_driver.err
echo "driver rc=$?"
python - "gpurun_out/r2d_n${N}_driver.log" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], "parity", (d.get("parity") or {}).get("ok"), "e2e", d.get("e2e"))
except Exception as e:
    print("    no bench line:", e)
PY
tail -5 gpurun_out/r2d_n${N}_driver.err | cut -c1-300
