"""One fused-GEMM configuration per process (a launch failure poisons the CUDA context): python tools/gemm_repro.py k n m reps"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laplace_gnn_b200 import ops
k, n, m, reps = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda:0")
ld = (k + 3) // 4 * 4
x = torch.randn(m, ld, device=dev)[:, :k]
w = torch.randn(k, n, device=dev) / k ** 0.5
act = torch.randn((m + 11) // 12, n, device=dev)
out = torch.empty(m, n, device=dev)
wp = ops.gemm_mask_prepare(w)
for _ in range(reps):
    ops.gemm_mask(x, wp, act, 12, out=out)
torch.cuda.synchronize()
ref = (x[:4096].double() @ w.double()) * (act[: (4096 + 11) // 12].repeat_interleave(12, 0)[:4096] > 0)
err = float((out[:4096].double() - ref).abs().max() / ref.abs().max())
print(f"k={k} n={n} m={m} reps={reps} tpc={os.environ.get('LGNN_GEMM_TILES_PER_CLUSTER', 'auto')}: ok, rel err {err:.1e}", flush=True)
