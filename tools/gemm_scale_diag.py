"""Does the 3xTF32 compensation of the fused GEMM depend on the operands' magnitudes?  (GPU box)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laplace_gnn_b200 import ops
dev = torch.device("cuda:0")
m = 200_000
for k, n in ((40, 256), (256, 256)):
    for sa, sw, dist in ((1.0, 1.0, "randn"), (1e-2, 1.0, "randn"), (1.0, 0.06, "randn"), (1e-2, 0.06, "randn"), (2e-3, 0.06, "uniform"),
                         (1e-3, 1e-3, "randn"), (1e2, 1e2, "randn")):
        g = torch.Generator(device=dev).manual_seed(k)
        x = torch.randn(m, k, device=dev, generator=g) * sa
        w = (torch.randn(k, n, device=dev, generator=g) if dist == "randn" else (torch.rand(k, n, device=dev, generator=g) * 2 - 1)) * sw
        wp = ops.gemm_mask_prepare(w)
        out = ops.gemm_mask(x, wp, None, 1)
        ref = x.double() @ w.double()
        scale = float((out.double() * ref).sum() / (ref * ref).sum()) - 1.0
        rms = float(((out.double() - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())
        # the same product with only hi.hi (plain tf32 truncation of both operands), for scale
        xt = (x.view(torch.int32) & -8192).view(torch.float32).double(); wt = (w.view(torch.int32) & -8192).view(torch.float32).double()
        r1 = xt @ wt
        s1 = float((r1 * ref).sum() / (ref * ref).sum()) - 1.0
        r2 = xt @ wt + (x.double() - xt) @ wt      # a compensated, w not
        s2 = float((r2 * ref).sum() / (ref * ref).sum()) - 1.0
        print(f"k={k} |a|~{sa:g} |w|~{sw:g} {dist}: kernel scale bias {scale:+.2e} rms {rms:.1e} | plain tf32 truncation of both would give {s1:+.2e}, of w only {s2:+.2e}", flush=True)
