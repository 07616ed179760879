#!/bin/bash
# 8-GPU pass: NCCL parity test + products bench at N=8 (both backward layouts) and N=4.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu8.txt
python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/t8.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t8.log
run() { # name nproc extra...
  name=$1; np=$2; shift 2
  NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np --steps 3 --warmup 3 --no-e2e "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/$name.log; echo
}
run b8_cols 8 --backward-parallel columns
run b8_rows 8 --backward-parallel rows
run b8_rows_noov 8 --backward-parallel rows --no-overlap
run b4_cols 4 --backward-parallel columns
run b4_rows 4 --backward-parallel rows
