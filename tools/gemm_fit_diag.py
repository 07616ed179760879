"""Accuracy of the fused GEMM on the data of a real fit (the arxiv-shaped 1/16 sample of bench.py's parity block):
every lgnn_gemm_mask_f32 call of one fit is repeated in float64 and compared.  Run on the GPU box."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops
dev = torch.device("cuda:0")
n, u, f, c, h, l = 169_343 // 16, 1_166_243 // 16, 128, 40, 256, 3
gen = torch.Generator(device=dev).manual_seed(1)
src = torch.randint(0, n, (u,), device=dev, generator=gen); dst = torch.randint(0, n, (u,), device=dev, generator=gen)
ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
X = torch.randn(n, f, device=dev, generator=gen)
idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
y = torch.randint(0, c, (idx.numel(),), device=dev, generator=gen)
torch.manual_seed(0)
model = L.SparseGCN(f, h, c, l, X, L.Graph.from_edge_index(ei, n, assume_undirected=True)).to(dev)
real = ops.gemm_mask
def spy(a, w, act, group, out=None, m_rows=None):
    res = real(a, w, act, group, out=out, m_rows=m_rows)
    m = a.shape[0] if m_rows is None else m_rows
    W = w.wt_hi.double()[:, : w.k].t()          # [k, n]: wt_hi is W itself (the tensor core truncates it), wt_lo = W - trunc(W)
    ref = a[:m, : w.k].double() @ W
    if act is not None:
        ref = ref * (act.repeat_interleave(group, 0)[:m, : w.n] > 0)
    got = res[:m, : w.n].double()
    terms = a[:m, : w.k].double().abs() @ W.abs()
    scale = float((got * ref).sum() / (ref * ref).sum()) - 1.0
    print(f"gemm k={w.k} n={w.n} m={m}: max rel err {float((got - ref).abs().max() / ref.abs().max()):.1e}  rms rel {float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()):.1e}"
          f"  scale bias {scale:+.1e}  cancellation sum|terms| / |result| (rms) {float((terms ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()):.1f}"
          f"  |a| max {float(a[:m, : w.k].abs().max()):.2e} rms {float((a[:m, : w.k] ** 2).mean().sqrt()):.2e}  zero rows of a {int((a[:m, : w.k].abs().max(1).values == 0).sum())}", flush=True)
    # which term is missing?  compare the kernel's output with the product under single-operand truncation
    tr = lambda t: (t.contiguous().view(torch.int32) & -8192).view(torch.float32).double()
    A64, At, Wt = a[:m, : w.k].double(), tr(a[:m, : w.k].float()), tr(W.float())
    def dist(r):
        if act is not None:
            r = r * (act.repeat_interleave(group, 0)[:m, : w.n] > 0)
        return float(((got - r) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())
    print(f"      rms distance of the kernel's output to: exact {dist(A64 @ W):.1e} | a truncated, W exact {dist(At @ W):.1e} | a exact, W truncated {dist(A64 @ Wt):.1e} | both truncated {dist(At @ Wt):.1e}", flush=True)
    # same data through fresh buffers / freshly prepared weights
    a2 = a[:m, : w.k].clone()
    w2 = ops.gemm_mask_prepare(W.float().contiguous())
    r2 = real(a2, w2, None, 1)[:, : w.n].double()
    ref2 = A64 @ W
    print(f"      fresh buffers + freshly prepared weights, no mask: scale bias {float((r2 * ref2).sum() / (ref2 * ref2).sum()) - 1.0:+.1e}; "
          f"prepared lo equal: {bool(torch.equal(w2.wt_lo, w.wt_lo))} hi equal: {bool(torch.equal(w2.wt_hi, w.wt_hi))}; |wt_lo| max {float(w.wt_lo.abs().max()):.2e}", flush=True)
    return res
ops.gemm_mask = spy
real_bias = ops.gemm_bias
def spy_bias(a, w, bias, out, m_rows=None):
    res = real_bias(a, w, bias, out, m_rows=m_rows)
    m = a.shape[0] if m_rows is None else m_rows
    W = w.wt_hi.double()[:, : w.k].t()
    ref = a[:m, : w.k].double() @ W + (bias.double() if bias is not None else 0.0)
    got = res[:m, : w.n].double()
    err = (got - ref).abs()
    print(f"gemm_bias k={w.k} n={w.n} m={m}: max rel err {float(err.max() / ref.abs().max()):.1e}  rms rel {float((err ** 2).mean().sqrt() / (ref ** 2).mean().sqrt()):.1e}"
          f"  rows with an error > 1e-5 of max: {int((err.max(1).values > 1e-5 * ref.abs().max()).sum())}  a pitch {a.stride(0)} out pitch {out.stride(0)}"
          f"  zeros in a {float((a[:m, : w.k] == 0).float().mean()):.2f}", flush=True)
    return res
ops.gemm_bias = spy_bias
for kw in ({}, {"unit_slabs": False}):
    print("backend kwargs", kw, flush=True)
    la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs=kw)
    la.fit(L.TensorBatchLoader(idx, y))
    print("marglik", float(la.log_marginal_likelihood()), flush=True)
