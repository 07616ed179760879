#!/bin/bash
# 2-GPU pass: all GPU tests (incl. NCCL + predictive), then the driver's exact N=2 bench command.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t1.log | cut -c1-200
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/d2.log 2>gpurun_out/d2.err; echo "N=2 default rc=$? in $(( $(date +%s) - S )) s"
tail -c 1200 gpurun_out/d2.log; tail -3 gpurun_out/d2.err
S=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/d2ref.log 2>gpurun_out/d2ref.err; echo "N=2 reference rc=$? in $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/d2ref.log
