#!/bin/bash
# 8-GPU pass: products bench at N=8 and N=4 (default column-parallel backward), NCCL parity test.
set -u
mkdir -p gpurun_out
run() { # name nproc extra...
  name=$1; np=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np --steps 3 --warmup 3 --no-e2e "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name rc=$?"; tail -c 400 gpurun_out/$name.log; echo
}
run c8_cols 8
run c4_cols 4
python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/t8.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t8.log
