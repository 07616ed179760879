#!/bin/bash
for fl in "" "--no-fused-linear" "--no-fused-gemm" "--no-fused-linear --syrk simt"; do
timeout 200 python bench.py --workload arxiv --steps 1 --warmup 1 --no-e2e $fl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv [$fl]:', d['parity']['vs_oracle']['per_block_rel'], d['parity']['marglik_gpu'])"
done
