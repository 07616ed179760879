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
def check(lo, hi):
    ref = (x[lo:hi].double() @ w.double()) * (act.repeat_interleave(12, 0)[lo:hi] > 0)
    got = out[lo:hi].double()
    err = float((got - ref).abs().max() / ref.abs().max())
    scale = float((got * ref).sum() / (ref * ref).sum()) - 1.0      # systematic shrink / growth of the product
    bad_rows = int(((got - ref).abs().max(1).values > 1e-4 * ref.abs().max()).sum())
    return err, scale, bad_rows
parts = [(0, min(m, 4096)), (max(0, m // 2 - 2048), min(m, m // 2 + 2048)), (max(0, m - 4096), m)]
res = [check(lo, hi) for lo, hi in parts]
print(f"k={k} n={n} m={m} reps={reps}: ok, rel err first/middle/last 4096 rows " + " / ".join(f"{e:.1e}" for e, _, _ in res) +
      "  scale bias " + " / ".join(f"{b:+.1e}" for _, b, _ in res) + "  rows off by > 1e-4: " + " / ".join(str(r) for _, _, r in res), flush=True)
