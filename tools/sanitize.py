"""Small-shape walk through every kernel of liblgnn.so, meant to run under
    compute-sanitizer --tool memcheck python tools/sanitize.py
(sanitizers on the small shapes).  Shapes hit the tails: odd widths, rows past a tile,
hub rows, empty rows, K not a multiple of 8/32, n not a multiple of 16."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops
from laplace_gnn_b200.ops import CSR
from conftest import Golden
from helpers import build_model, loader_for

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
n = 8000
ei = torch.randint(0, n, (2, 40_000), device=dev, generator=gen)
star = torch.stack([torch.zeros(6000, dtype=torch.int64, device=dev), torch.randperm(n, device=dev)[:6000]])
ei = torch.cat([ei, star, star.flip(0)], 1)
for sym in (False, True):
    g = L.Graph.from_edge_index(ei, n, symmetric=sym)
    b = g.partition_bounds(3)
    lo, hi = int(b[1]), int(b[2])
    ops.halo_columns(g.ahat_t, lo, hi)
    ops.csr_slice_remap(g.ahat, lo, hi, b, int((b[1:] - b[:-1]).max()))
print("graph kernels ok", g.nnz, int(g.deg.max()))
for d in (1, 7, 47, 48, 100, 256, 516, 960):
    x = torch.randn(n, d, device=dev, generator=gen)
    for relu in (False, True):
        ops.spmm(g.ahat, x, relu=relu)
        ops.spmm(g.ahat_t, x, relu=relu, impl="ldg")
for d in (512, 3076, 3840, 4100):
    x = torch.randn(n, d, device=dev, generator=gen)
    ops.spmm(g.ahat, x, impl="bulk")
    ops.spmm(g.ahat, x, relu=True, impl="bulk")
# rectangular CSR with empty rows and a hub beyond the bulk budget
cnt = np.random.default_rng(0).integers(0, 9, 200); cnt[[0, 5, 199]] = 0; cnt[50] = 40_000
rp = np.zeros(201, np.int64); np.cumsum(cnt, out=rp[1:])
a = CSR(200, 300, torch.from_numpy(rp).to(dev), torch.randint(0, 300, (int(rp[-1]),), device=dev, dtype=torch.int32),
        torch.randn(int(rp[-1]), device=dev))
x = torch.randn(300, 3900, device=dev)
ops.spmm(a, x, d=3840, impl="bulk"); ops.spmm(a, x, d=3840, impl="ldg"); ops.spmm(a, x, d=200)
keep = (torch.rand(300, device=dev) > 0.5).to(torch.uint8)
ops.spmm(ops.csr_with_masked_sources(a, keep), x, d=640)
print("spmm ok")
for C in (2, 7, 47):
    logits = torch.randn(n, C, device=dev)
    idx = torch.randperm(n, device=dev)[:3000].sort().values
    y = torch.randint(0, C, (3000,), device=dev)
    ops.softmax_ce_sum(logits, idx, y)
    cp = (C + 3) // 4 * 4
    delta = torch.zeros(n, C * cp, device=dev)
    for mode in ("reference", "ggn"):
        ops.hess_rhs(logits, idx, 0, C, delta, cp, mode)
ops.relu_mask_mul(torch.randn(n * 3, 20, device=dev), torch.randn(n, 20, device=dev), 3)
ops.relu_mask_mul(torch.randn(n * 3, 21, device=dev), torch.randn(n, 21, device=dev), 3)
print("hess ok")
for k, nn in ((1, 8), (17, 16), (1000, 47), (5000, 100), (3001, 130), (2049, 256), (70_000, 256)):
    ld = (nn + 3) // 4 * 4
    x = torch.randn(k, ld, device=dev)[:, :nn]
    ops.syrk(x, impl="tcgen05"); ops.syrk(x, impl="simt")
    ops.syrk(x, alpha=0.5, beta=1.0, out=torch.zeros(nn, nn, device=dev))
ops.syrk(torch.randn(999, 1433, device=dev))
print("syrk ok")
for m, k, nn, grp in ((1, 8, 64, 1), (129, 47, 256, 3), (5000, 200, 256, 4), (4097, 256, 128, 7), (128 * 33 + 1, 64, 64, 5)):
    ld = (k + 3) // 4 * 4
    x = torch.randn(m, ld, device=dev)[:, :k]
    w = torch.randn(k, nn, device=dev)
    act = torch.randn((m + grp - 1) // grp, nn, device=dev)
    wp = ops.gemm_mask_prepare(w)
    ops.gemm_mask(x, wp, act, grp); ops.gemm_mask(x, wp, None, grp)
print("gemm_mask ok")
for name in ("tiny_directed_3l", "tiny_directed_dups_2l", "pubmed_shape"):
    gd = Golden(name)
    model = build_model(gd, dev)
    la = L.Laplace(model, "classification", backend=L.B200GGN)
    la.fit(loader_for(gd, dev))
    ml = float(la.log_marginal_likelihood())
    assert abs(ml - gd.marglik) <= 1e-3 * abs(gd.marglik)
torch.cuda.synchronize()
print("SANITIZE_WALK_OK")
