"""GPU tests of the output-layer SpMM that rebuilds its Hessian-sqrt right-hand sides per edge
(``B200GGN(fused_hess_spmm=True)``, csrc/spmm_hess.cu; the default since round 2) against the materialised path."""
import numpy as np
import pytest
import torch

from conftest import max_rel_err
from oracle import gcn_kfac_oracle as O

pytestmark = [pytest.mark.gpu]

DEV = "cuda:0"


@pytest.mark.parametrize("C", [2, 3, 7, 32, 40, 47, 64])
@pytest.mark.parametrize("mode", ["reference", "ggn"])
def test_spmm_hess_matches_materialised_path(C, mode):
    """lgnn_hess_stats_f32 + lgnn_spmm_hess_f32 against lgnn_hess_rhs_f32 + lgnn_spmm_f32 on the same inputs
    (column groups of every width class, a zero-padded last group, repeated batch nodes, masked and unmasked
    edge values), and against the oracle's float64 SpMM of the float64 right-hand sides."""
    from laplace_gnn_b200 import ops
    import laplace_gnn_b200 as L
    n = 4000
    ei = O.synthetic_edges(n, 30_000, seed=C, directed=True)
    G = L.Graph.from_edge_index(torch.from_numpy(ei).to(DEV), n)
    R = O.build_graph(ei, n)
    gen = torch.Generator().manual_seed(C)
    cp = (C + 3) // 4 * 4
    logits = torch.zeros(n, cp)
    logits[:, :C] = 3 * torch.randn(n, C, generator=gen)
    idx = torch.randperm(n, generator=gen)[: int(0.6 * n)].sort().values
    idx = torch.cat([idx, idx[:7]])
    lg, ix = logits.to(DEV), idx.to(DEV)
    keep = torch.zeros(n, dtype=torch.uint8, device=DEV)
    keep[ix] = 1
    masked = ops.csr_with_masked_sources(G.ahat_t, keep)
    stats = ops.hess_stats(lg, ix, mode, C)
    Vref = O.hess_sqrt_rhs(logits[idx][:, :C].double(), mode)                 # [m, C, C] float64
    for width in (1, 4, 7, 12, 16):
        for c0 in range(0, C, width):
            ncols = min(width, C - c0)
            gq = (ncols + 3) // 4 * 4 if width > 1 else 1
            delta = torch.zeros(n, gq * cp, device=DEV)
            ops.hess_rhs(lg, ix, c0, ncols, delta, cp, mode, C)
            want = ops.spmm(masked, delta)
            for a in (masked, G.ahat_t):
                for staged in (False, True):          # register kernel / cp.async-ring kernel (c0 % 4 == 0 only)
                    got = ops.spmm_hess(a, stats, C, c0, ncols, gq, staged=staged)
                    assert max_rel_err(got.cpu().numpy(), want.cpu().numpy()) <= 1e-5, (width, c0, staged)
            d64 = torch.zeros(n, gq, cp, dtype=torch.float64)
            d64[:, :ncols, :C].index_add_(0, idx, Vref[:, c0:c0 + ncols, :])
            ref = O.spmm(R, d64.reshape(n, gq * cp).numpy(), transpose=True, dtype=torch.float64).numpy()
            assert max_rel_err(got.cpu().numpy(), ref) <= 1e-5, (width, c0)
            assert bool((got.view(n, gq, cp)[:, ncols:] == 0).all())
            if width > 8:
                break                                                            # one wide group is enough


@pytest.mark.parametrize("h,C,layers", [(64, 10, 3), (256, 47, 3), (32, 3, 2)])
def test_fused_hess_spmm_gives_the_same_factors(h, C, layers):
    import laplace_gnn_b200 as L
    n, U, F = 3000, 15_000, 20
    ei = torch.from_numpy(O.synthetic_edges(n, U, seed=h + C)).to(DEV)
    graph = L.Graph.from_edge_index(ei, n)
    gen = torch.Generator().manual_seed(h)
    X = torch.randn(n, F, generator=gen).to(DEV)
    torch.manual_seed(C)
    model = L.SparseGCN(F, h, C, layers, X, graph).to(DEV)
    idx = torch.randperm(n, generator=gen)[: int(0.6 * n)].sort().values.to(DEV)
    y = torch.randint(0, C, (idx.numel(),), generator=gen).to(DEV)
    for mode in ("reference", "ggn"):
        l1, k1 = L.B200GGN(model, "classification", hess_sqrt=mode, fused_hess_spmm=True).kron(idx, y, N=len(y))
        l2, k2 = L.B200GGN(model, "classification", hess_sqrt=mode, fused_hess_spmm=False).kron(idx, y, N=len(y))
        assert float(l1) == float(l2)
        for fa, fb in zip(k1.kfacs, k2.kfacs):
            for a, b in zip(fa, fb):
                assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5

