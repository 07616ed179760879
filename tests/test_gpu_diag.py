"""GPU: `B200GGN(diag_mode="node_factorised")` (SURVEY 8a row a13 / config 5 — the labelled approximation of the
diagonal GGN for graphs on which the exact diagonal, and the reference's own (M, C, P) Jacobian of
laplace/curvature/curvature.py:412-432, are out of reach) against its definition evaluated by brute force from the
oracle's float64 pieces on a 2,000-node graph, and against the EXACT diagonal where the two coincide (no edges)."""
import numpy as np
import pytest
import torch

from conftest import max_rel_err
from oracle import gcn_kfac_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _brute_force(R, X, Ws, bs, idx, n, C):
    """sum_n sum_c gZ_{l,c}[n, i]^2 H_{l-1}[n, j]^2 and sum_n sum_c gZ_{l,c}[n, i]^2 with the textbook Hessian square root."""
    hs, ps = O.forward(R, X, Ws, bs, torch.float64)
    V = O.hess_sqrt_rhs(ps[-1][idx], "ggn")
    at = torch.zeros(n, n, dtype=torch.float64)
    rows = np.repeat(np.arange(n), np.diff(R.t_rowptr))
    at[rows, R.t_col.astype(np.int64)] = torch.from_numpy(R.t_val.astype(np.float64))
    L_ = len(Ws)
    dw = [torch.zeros(w.shape, dtype=torch.float64) for w in Ws]
    db = [torch.zeros(w.shape[0], dtype=torch.float64) for w in Ws]
    for c in range(C):
        delta = torch.zeros(n, C, dtype=torch.float64).index_add(0, torch.from_numpy(idx), V[:, c, :])
        for l in range(L_ - 1, -1, -1):
            gz = at @ delta
            dw[l] += (gz ** 2).T @ (hs[l] ** 2)
            db[l] += (gz ** 2).sum(0)
            if l > 0:
                delta = (gz @ torch.from_numpy(Ws[l]).double()) * (ps[l - 1] > 0)
    return torch.cat([t for l in range(L_) for t in (dw[l].reshape(-1), db[l])])


@pytest.mark.parametrize("layers,h,C", [(2, 64, 5), (3, 256, 12)])
def test_node_factorised_diag_matches_its_definition(layers, h, C):
    import laplace_gnn_b200 as L
    n, F = 2000, 20
    ei = O.synthetic_edges(n, 9000, seed=layers, directed=True)
    gen = torch.Generator().manual_seed(layers)
    X = torch.randn(n, F, generator=gen)
    idx = torch.randperm(n, generator=gen)[: int(0.6 * n)].sort().values
    y = torch.randint(0, C, (idx.numel(),), generator=gen)
    torch.manual_seed(layers)
    model = L.SparseGCN(F, h, C, layers, X.to(DEV), L.Graph.from_edge_index(torch.from_numpy(ei).to(DEV), n)).to(DEV)
    Ws = [c.lin.weight.detach().cpu().numpy() for c in model.convs]
    bs = [c.lin.bias.detach().cpu().numpy() for c in model.convs]
    want = _brute_force(O.build_graph(ei, n), X.numpy(), Ws, bs, idx.numpy(), n, C)
    # default column grouping (unit-compacted slabs where they apply) and a budget that forces several groups
    for kw in ({}, {"rhs_tile_bytes": 2 * n * h * 4 * 4, "unit_min_width": 0}, {"unit_slabs": False}):
        be = L.B200GGN(model, "classification", diag_mode="node_factorised", **kw)
        loss, d = be.diag(idx.to(DEV), y.to(DEV))
        assert d.is_cuda and d.dtype == torch.float32 and d.numel() == sum(p.numel() for p in model.parameters())
        assert max_rel_err(d.cpu().numpy(), want.numpy()) <= 1e-4, kw
    # the reference's DiagLaplace algebra runs on it
    la = L.Laplace(model, "classification", hessian_structure="diag", backend=L.B200GGN,
                   backend_kwargs={"diag_mode": "node_factorised"})
    la.fit(L.TensorBatchLoader(idx.to(DEV), y.to(DEV)))
    assert torch.isfinite(la.log_marginal_likelihood())


def test_node_factorised_diag_is_the_exact_diagonal_without_edges():
    import laplace_gnn_b200 as L
    n, F, h, C = 300, 7, 64, 5
    gen = torch.Generator().manual_seed(1)
    X = torch.randn(n, F, generator=gen).to(DEV)
    idx = torch.randperm(n, generator=gen)[:180].sort().values.to(DEV)
    y = torch.randint(0, C, (180,), generator=gen).to(DEV)
    torch.manual_seed(1)
    model = L.SparseGCN(F, h, C, 3, X, L.Graph.from_edge_index(torch.zeros(2, 0, dtype=torch.int64, device=DEV), n)).to(DEV)
    exact = L.B200GGN(model, "classification").diag(idx, y)
    approx = L.B200GGN(model, "classification", diag_mode="node_factorised", unit_min_width=0).diag(idx, y)
    assert float(exact[0]) == float(approx[0])
    assert max_rel_err(approx[1].cpu().numpy(), exact[1].cpu().numpy()) <= 1e-4
