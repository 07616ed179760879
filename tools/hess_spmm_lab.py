"""Output-layer SpMM of the KFAC backward on the products-shaped graph: materialised right-hand sides
(lgnn_hess_rhs_f32 + lgnn_spmm_f32 over the masked CSR) against the on-the-fly kernel (lgnn_hess_stats_f32 once +
lgnn_spmm_hess_f32 per group).  Run on the GPU box:  python tools/hess_spmm_lab.py [g ...]   (default 16 8 6)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops
dev = torch.device("cuda:0")
n, pairs, C = 2_449_029, 61_859_140, 47
scale = float(os.environ.get("LAB_SCALE", "1"))
n, pairs = int(n * scale), int(pairs * scale)
gen = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, n, (pairs,), device=dev, generator=gen); dst = torch.randint(0, n, (pairs,), device=dev, generator=gen)
g = L.Graph.from_edge_index(torch.stack([torch.cat([src, dst]), torch.cat([dst, src])]), n, assume_undirected=True)
del src, dst
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cp = (C + 3) // 4 * 4
logits = torch.zeros(n, cp, device=dev); logits[:, :C] = torch.randn(n, C, device=dev, generator=gen)
idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
keep = torch.zeros(n, dtype=torch.uint8, device=dev); keep[idx] = 1
masked = ops.csr_with_masked_sources(g.ahat_t, keep)
live = masked.nnz_gathered
stats = torch.empty(n, 5 * cp, device=dev)
t_stats = timed(lambda: ops.hess_stats(logits, idx, "reference", C, out=stats))
print(f"n={n} nnz={g.nnz} gathered edges={live} peak={peak} GB/s; hess_stats {t_stats:.2f} ms (once per fit)", flush=True)
for gg in [int(a) for a in sys.argv[1:]] or [16, 8, 6]:
    gq = (gg + 3) // 4 * 4
    delta = torch.zeros(n, gq * cp, device=dev); y = torch.empty(n, gq * cp, device=dev); y2 = torch.empty_like(y)
    def materialised():
        delta.zero_(); ops.hess_rhs(logits, idx, 0, gg, delta, cp, "reference", C); ops.spmm(masked, delta, out=y)
    t_rhs = timed(lambda: (delta.zero_(), ops.hess_rhs(logits, idx, 0, gg, delta, cp, "reference", C)))
    t_old = timed(materialised)
    t_reg = timed(lambda: ops.spmm_hess(masked, stats, C, 0, gg, gq, out=y2, staged=False))
    t_new = timed(lambda: ops.spmm_hess(masked, stats, C, 0, gg, gq, out=y2, staged=True))
    by_old = g.nnz * 8 + (n + 1) * 8 + live * gq * cp * 4 + n * gq * cp * 4
    by_new = g.nnz * 8 + (n + 1) * 8 + live * (2 * cp + 3 * gg) * 4 + n * gq * cp * 4
    err = float((y2 - y).abs().max() / y.abs().max())
    print(f"g={gg}: materialised {t_old:7.2f} ms (rhs {t_rhs:5.2f} + spmm {t_old - t_rhs:6.2f}: {by_old / (t_old - t_rhs) / 1e6:5.0f} GB/s)"
          f" | on the fly: register kernel {t_reg:7.2f} ms, staged {t_new:7.2f} ms ({by_new / t_new / 1e6:5.0f} GB/s of its own bytes, {by_new / t_new / 1e6 / peak:4.2f} of peak)"
          f" | speed-up {t_old / t_new:4.2f}x | max rel diff {err:.1e}", flush=True)
    del delta, y, y2
