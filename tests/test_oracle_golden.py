"""CPU: the oracle (oracle/gcn_kfac_oracle.py) against the golden vectors produced by the
reference's own classes (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from conftest import GgnGolden, Golden, GOLDEN_KERNEL_SHAPES, GOLDEN_SMALL, max_rel_err
from oracle import gcn_kfac_oracle as O


def _graph(g):
    return O.build_graph(g.edge_index, g.n, g.symmetric)


def test_normalised_adjacency_matches_reference_dense(golden_small):
    g = golden_small
    G = _graph(g)
    dense = np.zeros((g.n, g.n), dtype=np.float32)
    rows = np.repeat(np.arange(g.n), np.diff(G.rowptr))
    dense[rows, G.col] = G.val
    ref = g.z["ahat_dense"]
    assert np.array_equal(dense != 0, ref != 0)          # pattern: bit-exact
    assert np.abs(dense - ref).max() <= 2 * np.finfo(np.float32).eps  # values: <= 2 ulp of pow(-0.5)
    # Â^T really is the transpose
    dense_t = np.zeros_like(dense)
    rows_t = np.repeat(np.arange(g.n), np.diff(G.t_rowptr))
    dense_t[rows_t, G.t_col] = G.t_val
    assert np.array_equal(dense_t, dense.T)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 2e-6)])
def test_factors_loss_marglik_match_reference(golden, dtype, tol):
    g = golden
    G = _graph(g)
    bs = None if g.batch_size == len(g.idx) else g.batch_size
    loss, kfacs, ml = O.fit_and_marglik(G, g.x, g.Ws, g.bs, g.idx, g.y, 1.0, "reference", dtype, bs)
    assert len(kfacs) == len(g.kfacs)
    for blk, ref_blk in zip(kfacs, g.kfacs):
        assert len(blk) == len(ref_blk)
        for h, ref in zip(blk, ref_blk):
            assert max_rel_err(h.numpy(), ref) <= tol
    assert abs(float(loss) - g.loss) <= 1e-5 * abs(g.loss)
    assert abs(float(ml) - g.marglik) <= 1e-5 * abs(g.marglik)


@pytest.mark.parametrize("name", GOLDEN_KERNEL_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_factors_loss_marglik_match_reference_at_kernel_shapes(name, dtype):
    """3 layers, h = 256, C = 40 — the shapes the fused GEMM, the tcgen05 SYRK and the unit-compacted slabs
    take on the GPU — against the reference's own fit (O2 tier)."""
    g = Golden(name)
    loss, kfacs, ml = O.fit_and_marglik(_graph(g), g.x, g.Ws, g.bs, g.idx, g.y, 1.0, "reference", dtype)
    for blk, ref_blk in zip(kfacs, g.kfacs):
        for h, ref in zip(blk, ref_blk):
            assert max_rel_err(h.numpy(), ref) <= 5e-6
    assert abs(float(loss) - g.loss) <= 1e-5 * abs(g.loss)
    assert abs(float(ml) - g.marglik) <= 1e-5 * abs(g.marglik)
    hs, ps = O.forward(_graph(g), g.x, g.Ws, g.bs)
    assert max_rel_err(ps[-1][torch.from_numpy(g.idx)].numpy(), g.z["logits"]) <= 1e-5


def test_logits_match_reference(golden_small):
    g = golden_small
    hs, ps = O.forward(_graph(g), g.x, g.Ws, g.bs)
    assert max_rel_err(ps[-1][torch.from_numpy(g.idx)].numpy(), g.z["logits"]) <= 1e-5


def test_ggn_mode_differs_from_fork_and_matches_textbook(golden_small):
    """hess_sqrt='ggn' is NOT what the fork computes; it equals J^T Λ J."""
    g = golden_small
    G = _graph(g)
    _, kf_ref, _ = O.fit_and_marglik(G, g.x, g.Ws, g.bs, g.idx, g.y, 1.0, "reference", torch.float64)
    _, kf_ggn, _ = O.fit_and_marglik(G, g.x, g.Ws, g.bs, g.idx, g.y, 1.0, "ggn", torch.float64)
    assert max_rel_err(kf_ggn[0][0].numpy(), kf_ref[0][0].numpy()) > 1e-3
    # textbook identity on the last layer's bias block: G_L = sum_n Â^T-propagated Λ
    V = O.hess_sqrt_rhs(O.forward(G, g.x, g.Ws, g.bs, torch.float64)[1][-1][torch.from_numpy(g.idx)], "ggn")
    p = torch.softmax(O.forward(G, g.x, g.Ws, g.bs, torch.float64)[1][-1][torch.from_numpy(g.idx)], 1)
    lam = torch.diag_embed(p) - p.unsqueeze(2) * p.unsqueeze(1)
    assert torch.allclose(torch.einsum("nck,ncj->nkj", V, V), lam, atol=1e-12)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_ggn_mode_matches_upstream_curvlinops_arithmetic(golden_small, dtype):
    """hess_sqrt="ggn" pinned: the reference's own classes run with upstream curvlinops' ``out.detach()`` restored
    on the Hessian-sqrt input (the one expression the fork changed, curvlinops/kfac.py:631-642) — factors, loss and
    marglik of that run (oracle/make_golden_ggn.py), multi-batch case included."""
    g = golden_small
    gg = GgnGolden(g.name)
    assert abs(gg.fork_marglik - g.marglik) <= 1e-6 * abs(g.marglik) and gg.marglik != g.marglik
    bs = None if g.batch_size == len(g.idx) else g.batch_size
    loss, kfacs, ml = O.fit_and_marglik(_graph(g), g.x, g.Ws, g.bs, g.idx, g.y, 1.0, "ggn", dtype, bs)
    assert len(kfacs) == len(gg.kfacs)
    for blk, ref_blk in zip(kfacs, gg.kfacs):
        for h, ref in zip(blk, ref_blk):
            assert max_rel_err(h.numpy(), ref) <= 2e-6
    assert abs(float(loss) - gg.loss) <= 1e-5 * abs(gg.loss)
    assert abs(float(ml) - gg.marglik) <= 1e-5 * abs(gg.marglik)


@pytest.mark.parametrize("name", [n for n in GOLDEN_SMALL if n.startswith("tiny")])
def test_exact_diag_ggn_matches_reference(name):
    g = Golden(name)
    loss, dg = O.diag_ggn(_graph(g), g.x, g.Ws, g.bs, g.idx, g.y)
    assert max_rel_err(dg.numpy(), g.z["diag_H"]) <= 1e-5


def test_row_partition_and_halo():
    ei = O.synthetic_edges(500, 2000, seed=3)
    G = O.build_graph(ei, 500)
    for parts in (1, 2, 3, 8):
        b = O.row_partition(G.rowptr, parts)
        assert b[0] == 0 and b[-1] == 500 and np.all(np.diff(b) >= 0)
        nnz = np.diff(G.rowptr[b])
        assert nnz.sum() == G.nnz
        assert nnz.max() - nnz.min() <= 2 * np.diff(G.rowptr).max()
    b = O.row_partition(G.rowptr, 4)
    halo = O.halo_columns(G.rowptr, G.col, int(b[1]), int(b[2]))
    assert np.all((halo < b[1]) | (halo >= b[2])) and np.all(np.diff(halo) > 0)


def test_edge_cases_empty_and_isolated():
    G = O.build_graph(np.zeros((2, 0), dtype=np.int64), 5)
    assert G.nnz == 5 and np.array_equal(G.col, np.arange(5)) and np.all(G.val == 1.0)
    with pytest.raises(ValueError):
        O.coo_to_adj_csr(np.array([[0], [7]]), 5)
