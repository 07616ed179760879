#!/bin/bash
# 2-GPU pass: parity tests (incl. NCCL), SYRK lab, products bench at N=2 in the three backward layouts.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t1.log
timeout 300 python tools/syrk_lab.py --n 256,48 --impl tcgen05 > gpurun_out/syrk_lab.log 2>&1; cat gpurun_out/syrk_lab.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --backward-parallel rows > gpurun_out/b2_rows.log 2>gpurun_out/b2_rows.err; echo "rows rc=$?"
$TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --backward-parallel rows --no-overlap > gpurun_out/b2_rows_noov.log 2>gpurun_out/b2_rows_noov.err; echo "rows-noov rc=$?"
$TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --backward-parallel columns > gpurun_out/b2_cols.log 2>gpurun_out/b2_cols.err; echo "cols rc=$?"
python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/b1.log 2>gpurun_out/b1.err; echo "n1 rc=$?"
for f in b1 b2_rows b2_rows_noov b2_cols; do echo "== $f"; tail -c 900 gpurun_out/$f.log; tail -3 gpurun_out/$f.err; done
