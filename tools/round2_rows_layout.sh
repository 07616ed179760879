#!/bin/bash
# N GPUs: the row-partitioned backward ("rows", the north star's own partitioning) at the products shape — ragged
# unit-compacted rows travelling (default) against dense slabs (--no-unit-rows), two column groups in flight against
# one, the output layer from all-gathered softmax statistics (--rows-hess-stats, opt-in) — after the NCCL parity tests and the single-GPU kernel tests of the ragged layout.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/round2_rows_layout.sh 2 [tests|bench|all]'
set -u
N=${1:-2}
WHAT=${2:-all}
mkdir -p gpurun_out
P=29640
if [ $WHAT = tests ] || [ $WHAT = all ]; then
  timeout 300 python -m pytest tests/test_gpu_units_ragged.py -x -q > gpurun_out/r2r_ragged_tests.log 2>&1; echo "ragged kernel tests rc=$?"
  tail -3 gpurun_out/r2r_ragged_tests.log | cut -c1-300
  timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2r_n${N}_tests.log 2>&1; echo "multi-GPU tests rc=$?"
  grep -E "Error|assert" gpurun_out/r2r_n${N}_tests.log | head -8 | cut -c1-300
  tail -3 gpurun_out/r2r_n${N}_tests.log | cut -c1-200
fi
if [ $WHAT = bench ] || [ $WHAT = all ]; then
  for tag in ${TAGS:-rows rows_noov rows_dense rows_stats}; do
    fl="--backward-parallel rows"
    [ $tag = rows_noov ] && fl="$fl --no-overlap"
    [ $tag = rows_dense ] && fl="$fl --no-unit-rows"
    [ $tag = rows_stats ] && fl="$fl --rows-hess-stats"
    par="--no-parity"; [ $tag = rows ] && par=""
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
        bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-e2e --no-cpu-baseline $par $fl > gpurun_out/r2r_n${N}_$tag.log 2>gpurun_out/r2r_n${N}_$tag.err
    echo "$tag rc=$?"; P=$((P + 1))
    grep -E "LgnnError|Error|File \"/.*laplace_gnn_b200|File \"/.*bench.py" gpurun_out/r2r_n${N}_$tag.err | head -12 | cut -c1-260
    python - "gpurun_out/r2r_n${N}_$tag.log" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ", round(d["value"]), "nodes/s", round(d["ms_per_step"], 1), "ms  marglik", d["marglik"], "parity", (d.get("parity") or {}).get("ok"),
          d["roofline"]["ms_per_step_by_kind"], "group", d["config"].get("column_group"))
except Exception as e:
    print("    no bench line:", e)
PY
  done
fi
