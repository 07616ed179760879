#!/bin/bash
# 8-GPU pass: the driver's exact scaling commands at N=8 (own arm with e2e, then the reference arm).
set -u
mkdir -p gpurun_out
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/d8.log 2>gpurun_out/d8.err; echo "N=8 default rc=$? in $(( $(date +%s) - S )) s"
tail -c 1500 gpurun_out/d8.log; tail -3 gpurun_out/d8.err
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/d8ref.log 2>gpurun_out/d8ref.err; echo "N=8 reference rc=$? in $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/d8ref.log
