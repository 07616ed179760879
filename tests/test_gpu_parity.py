"""GPU parity tests: every call goes through the C-ABI (liblgnn.so) on cuda:0 and is compared with
the CPU oracle on the same seeded inputs and with the committed golden fixtures (reference outputs).

Tolerances (BASELINE.json north star): bit-exact for CSR / normalisation indices / partitioning;
<= 1e-5 relative for SpMM; <= 1e-4 relative for Kronecker factors; <= 1e-3 relative for the log
marginal likelihood.  "relative" = max |a-b| / max |b| over the tensor.
"""
import numpy as np
import pytest
import torch

from conftest import Golden, max_rel_err
from helpers import build_model, check_against_golden, loader_for
from oracle import gcn_kfac_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from laplace_gnn_b200 import ops
    return ops


def _dev_graph(ei, n, symmetric=False):
    import laplace_gnn_b200 as L
    return L.Graph.from_edge_index(torch.from_numpy(ei).to(DEV), n, symmetric=symmetric)


def _assert_graph_bit_exact(G, R):
    for name, a, b in [("rowptr", G.ahat.rowptr, R.rowptr), ("col", G.ahat.col, R.col),
                       ("t_rowptr", G.ahat_t.rowptr, R.t_rowptr), ("t_col", G.ahat_t.col, R.t_col),
                       ("deg", G.deg, R.deg)]:
        assert np.array_equal(a.cpu().numpy(), b), name
    # fp32 values are bit-exact too: IEEE sqrt / div / mul on both sides
    assert np.array_equal(G.dis.cpu().numpy().view(np.uint32), R.dis.view(np.uint32))
    assert np.array_equal(G.ahat.val.cpu().numpy().view(np.uint32), R.val.view(np.uint32))
    assert np.array_equal(G.ahat_t.val.cpu().numpy().view(np.uint32), R.t_val.view(np.uint32))


# ---------------------------------------------------------------------------------- integer kernels
@pytest.mark.parametrize("n,U,directed,symmetric,rmat", [
    (1, 0, False, False, False),            # single node, no edges
    (7, 0, True, False, False),             # empty edge list: self loops only
    (40, 70, False, False, False),
    (50, 400, True, False, False),          # directed, duplicates likely
    (36, 60, True, True, False),            # symmetrise on the device
    (2708, 5278, False, False, False),      # Cora shape
    (19717, 44324, False, False, False),    # Pubmed shape
    (4096, 200000, True, False, True),      # R-MAT: rows > 256 (block sort) and heavy duplicates
    (300, 120000, True, False, False),      # dense-ish: every row long
])
def test_csr_build_bit_exact(n, U, directed, symmetric, rmat):
    ei = O.synthetic_edges(n, U, seed=n + U, directed=directed, rmat=rmat)
    G = _dev_graph(ei, n, symmetric)
    R = O.build_graph(ei, n, symmetric)
    _assert_graph_bit_exact(G, R)


def test_csr_build_very_long_row_global_sort():
    # one hub with > 8192 distinct neighbours -> in-place global-memory bitonic path
    n = 20000
    rng = np.random.Generator(np.random.PCG64(5))
    hub_dst = rng.permutation(n)[:12000]
    ei = np.concatenate([np.stack([np.zeros(12000, np.int64), hub_dst]),
                         O.synthetic_edges(n, 30000, seed=9, directed=True)], axis=1)
    _assert_graph_bit_exact(_dev_graph(ei, n), O.build_graph(ei, n))


def test_csr_build_rejects_out_of_range():
    ei = torch.tensor([[0, 1], [1, 9]], dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError):
        _ops().csr_from_edge_index(ei, 5)


def test_golden_graphs_bit_exact(golden):
    g = golden
    _assert_graph_bit_exact(_dev_graph(g.edge_index, g.n, g.symmetric), O.build_graph(g.edge_index, g.n, g.symmetric))


@pytest.mark.parametrize("parts", [1, 2, 3, 8, 64])
def test_row_partition_and_halo_bit_exact(parts):
    ops = _ops()
    ei = O.synthetic_edges(5000, 40000, seed=1, rmat=True)
    G = _dev_graph(ei, 5000)
    R = O.build_graph(ei, 5000)
    b = ops.row_partition(G.ahat.rowptr, parts).cpu().numpy()
    assert np.array_equal(b, O.row_partition(R.rowptr, parts))
    r = min(1, parts - 1)
    lo, hi = int(b[r]), int(b[r + 1])
    assert np.array_equal(ops.halo_columns(G.ahat, lo, hi).cpu().numpy(), O.halo_columns(R.rowptr, R.col, lo, hi))
    # padded all-gather remap: decoding the remapped columns gives the original ones back
    pad = int(np.diff(b).max())
    sl = ops.csr_slice_remap(G.ahat, lo, hi, torch.from_numpy(b).to(DEV), pad)
    c = sl.col.cpu().numpy().astype(np.int64)
    decoded = b[c // pad] + c % pad
    assert np.array_equal(decoded, R.col[R.rowptr[lo]:R.rowptr[hi]])
    assert np.array_equal(sl.rowptr.cpu().numpy(), R.rowptr[lo:hi + 1] - R.rowptr[lo])
    assert np.array_equal(sl.val.cpu().numpy(), R.val[R.rowptr[lo]:R.rowptr[hi]])


# ---------------------------------------------------------------------------------- SpMM
@pytest.mark.parametrize("d", [1, 3, 7, 16, 47, 64, 100, 128, 256, 376, 576, 720, 1024, 2048])
@pytest.mark.parametrize("relu", [False, True])
def test_spmm_matches_oracle(d, relu):
    ops = _ops()
    n = 3000
    ei = O.synthetic_edges(n, 20000, seed=d, directed=True, rmat=(d % 2 == 0))
    G, R = _dev_graph(ei, n), O.build_graph(ei, n)
    rng = np.random.Generator(np.random.PCG64(d))
    x = rng.standard_normal((n, d)).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    for transpose, csr in ((False, G.ahat), (True, G.ahat_t)):
        ref = O.spmm(R, x, transpose=transpose, dtype=torch.float64)
        if relu:
            ref = torch.relu(ref)
        y = ops.spmm(csr, xd, relu=relu)
        assert max_rel_err(y.cpu().numpy(), ref.numpy()) <= 1e-5


@pytest.mark.parametrize("d", [512, 516, 960, 1024, 3072, 3076, 4096, 7000])
@pytest.mark.parametrize("relu", [False, True])
def test_spmm_bulk_kernel_matches_oracle_and_ldg_kernel(d, relu):
    """Wide (multi-RHS) SpMM: the bulk-async ring kernel vs the oracle and vs the warp-per-row kernel
    (summation order over a row's neighbours is the same in both, so they agree to the bit)."""
    ops = _ops()
    n = 2000
    ei = O.synthetic_edges(n, 9000, seed=d, directed=True, rmat=(d % 8 == 0))
    G, R = _dev_graph(ei, n), O.build_graph(ei, n)
    rng = np.random.Generator(np.random.PCG64(d))
    x = rng.standard_normal((n, d)).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    for transpose, csr in ((False, G.ahat), (True, G.ahat_t)):
        ref = O.spmm(R, x, transpose=transpose, dtype=torch.float64)
        if relu:
            ref = torch.relu(ref)
        y = ops.spmm(csr, xd, relu=relu, impl="bulk")
        assert max_rel_err(y.cpu().numpy(), ref.numpy()) <= 1e-5
        assert torch.equal(y, ops.spmm(csr, xd, relu=relu, impl="ldg"))
        assert torch.equal(y, ops.spmm(csr, xd, relu=relu))          # auto: whichever kernel the width selects


def test_spmm_bulk_kernel_empty_rows_hub_rows_and_leading_dimensions():
    """Rectangular CSR with empty rows (leading, interior, trailing), a row longer than one CTA's nnz
    budget (kept whole), two hub rows long enough to be split across CTAs (added with red.global —
    one of them starting a budget, one ending the matrix), and ldx / ldy larger than d."""
    ops = _ops()
    from laplace_gnn_b200.ops import CSR
    rng = np.random.Generator(np.random.PCG64(7))
    n_rows, n_cols, d, ld = 300, 500, 3840, 3900      # a width "auto" gives to the bulk kernel
    counts = rng.integers(0, 12, n_rows)
    counts[[0, 1, 57, 58, 298, 299]] = 0
    counts[100] = 9000                                    # > BULK_NNZ_PER_CTA, below the hub threshold
    counts[200] = 50_000                                  # hub: split over ~12 CTA budgets
    counts[297] = 20_000                                  # hub that is the last non-empty row
    rowptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(counts, out=rowptr[1:])
    col = rng.integers(0, n_cols, rowptr[-1]).astype(np.int32)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    x = rng.standard_normal((n_cols, ld)).astype(np.float32)
    a = CSR(n_rows, n_cols, torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV),
            torch.from_numpy(val).to(DEV))
    import scipy.sparse as sp
    ref = sp.csr_matrix((val.astype(np.float64), col, rowptr), shape=(n_rows, n_cols)) @ x[:, :d].astype(np.float64)
    xd = torch.from_numpy(x).to(DEV)
    for impl in ("bulk", "auto"):
        out = torch.full((n_rows, ld), 7.0, device=DEV)
        ops.spmm(a, xd, out=out, d=d, impl=impl)
        assert max_rel_err(out[:, :d].cpu().numpy(), ref) <= 1e-5
        assert float((out[:, d:] - 7.0).abs().max()) == 0.0   # columns beyond d untouched
        assert float(out[[0, 1, 57, 58, 298, 299], :d].abs().max()) == 0.0
        whole = np.setdiff1d(np.arange(n_rows), [100, 200, 297])   # rows neither kernel splits: bit-equal
        ldg = ops.spmm(a, xd, d=d, impl="ldg")
        assert torch.equal(out[whole, :d], ldg[whole])
    relu = ops.spmm(a, xd, d=d, relu=True, impl="bulk")   # fused relu: hub rows are not split
    assert max_rel_err(relu.cpu().numpy(), np.maximum(ref, 0)) <= 1e-5


@pytest.mark.parametrize("d", [16, 64, 256, 576, 720, 1024, 2256])
@pytest.mark.parametrize("relu", [False, True])
def test_spmm_hub_rows_of_power_law_graphs(d, relu):
    """A star on top of a random graph: rows of 30 k and 9 k non-zeros are split into 2048-entry
    segments (warp-per-row path) and summed with red.global; relu is applied after the full sum."""
    ops = _ops()
    n = 40_000
    rng = np.random.Generator(np.random.PCG64(d))
    base = O.synthetic_edges(n, 60_000, seed=d)
    star0 = np.stack([np.zeros(30_000, np.int64), rng.permutation(n)[:30_000]])
    star1 = np.stack([np.full(9_000, 17, np.int64), rng.permutation(n)[:9_000]])
    stars = np.concatenate([star0, star1], axis=1)
    ei = np.concatenate([base, stars, stars[::-1]], axis=1)
    G, R = _dev_graph(ei, n), O.build_graph(ei, n)
    assert int(np.diff(R.rowptr).max()) >= 30_000          # the star (+ its self loop unless the permutation holds node 0)
    x = rng.standard_normal((n, d)).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    ref = O.spmm(R, x, dtype=torch.float64)
    if relu:
        ref = torch.relu(ref)
    y = ops.spmm(G.ahat, xd, relu=relu, impl="ldg")
    assert max_rel_err(y.cpu().numpy(), ref.numpy()) <= 1e-5
    # rows below the hub threshold are bit-reproducible
    small = torch.from_numpy(np.flatnonzero(np.diff(R.rowptr) <= 4096)).to(DEV)
    assert torch.equal(y[small], ops.spmm(G.ahat, xd, relu=relu, impl="ldg")[small])


def test_spmm_strided_and_unaligned_operands():
    ops = _ops()
    n, d = 1000, 24
    ei = O.synthetic_edges(n, 6000, seed=2)
    G, R = _dev_graph(ei, n), O.build_graph(ei, n)
    big = torch.randn(n, 40, device=DEV)
    ref = O.spmm(R, big[:, :d].cpu().numpy(), dtype=torch.float64).numpy()
    out = torch.zeros(n, 32, device=DEV)
    ops.spmm(G.ahat, big, out=out, d=d)                      # ld != d, vector path
    assert max_rel_err(out[:, :d].cpu().numpy(), ref) <= 1e-5 and float(out[:, d:].abs().max()) == 0.0
    odd = torch.randn(n, 41, device=DEV)[:, 1:]              # misaligned base -> scalar path
    ref2 = O.spmm(R, odd[:, :d].cpu().numpy(), dtype=torch.float64).numpy()
    assert max_rel_err(ops.spmm(G.ahat, odd, d=d).cpu().numpy(), ref2) <= 1e-5


def test_spmm_linearity_and_transpose_adjoint_at_scale():
    """Size-independent properties on a graph far too big for the oracle's comfort:
    Â(ax + by) = aÂx + bÂy and <Âx, y> = <x, Â^T y>."""
    ops = _ops()
    n, d = 400_000, 256
    gen = torch.Generator(device=DEV).manual_seed(0)
    ei = torch.randint(0, n, (2, 4_000_000), device=DEV, generator=gen)
    import laplace_gnn_b200 as L
    G = L.Graph.from_edge_index(ei, n)
    x = torch.randn(n, d, device=DEV, generator=gen)
    y = torch.randn(n, d, device=DEV, generator=gen)
    lhs = ops.spmm(G.ahat, 0.5 * x - 2.0 * y)
    rhs = 0.5 * ops.spmm(G.ahat, x) - 2.0 * ops.spmm(G.ahat, y)
    assert float((lhs - rhs).abs().max()) <= 1e-5 * float(rhs.abs().max())
    a = (ops.spmm(G.ahat, x).double() * y.double()).sum()
    b = (x.double() * ops.spmm(G.ahat_t, y).double()).sum()
    assert abs(float(a - b)) <= 1e-6 * abs(float(a))
    # rows of a row-stochastic-like check: Â 1 has entries sum_j dis_i dis_j over in-neighbours
    ones = torch.ones(n, 4, device=DEV)
    deg_in = ops.spmm(G.ahat, ones)[:, 0]
    assert torch.isfinite(deg_in).all() and float(deg_in.min()) > 0


# ---------------------------------------------------------------------------------- loss / Hessian sqrt / mask
@pytest.mark.parametrize("C", [2, 3, 7, 40, 47, 200])
@pytest.mark.parametrize("mode", ["reference", "ggn"])
def test_hess_rhs_and_loss_match_oracle(C, mode):
    ops = _ops()
    n, m = 500, 300
    rng = np.random.Generator(np.random.PCG64(C))
    logits = (3 * rng.standard_normal((n, C))).astype(np.float32)
    idx = np.sort(rng.permutation(n)[:m]).astype(np.int64)
    idx[5] = idx[4]                                       # duplicate train index accumulates
    y = rng.integers(0, C, m).astype(np.int64)
    ld = (C + 3) // 4 * 4
    lg = torch.zeros(n, ld, device=DEV)
    lg[:, :C] = torch.from_numpy(logits).to(DEV)
    idx_d, y_d = torch.from_numpy(idx).to(DEV), torch.from_numpy(y).to(DEV)
    loss, hits = ops.softmax_ce_sum(lg, idx_d, y_d, C=C)
    f = torch.from_numpy(logits)[torch.from_numpy(idx)]
    ref_loss = O.cross_entropy_sum(f.double(), y)
    assert abs(float(loss) - float(ref_loss)) <= 1e-6 * abs(float(ref_loss))
    assert int(hits) == int((f.argmax(1).numpy() == y).sum())
    V = O.hess_sqrt_rhs(f.double(), mode)                 # [m, C, C]
    c0, g = (1, min(3, C - 1))
    delta = torch.zeros(n, g * ld, device=DEV)
    ops.hess_rhs(lg, idx_d, c0, g, delta, ld, mode, C=C)
    ref = torch.zeros(n, g, ld, dtype=torch.float64)
    for k in range(g):
        ref[:, k, :C].index_add_(0, torch.from_numpy(idx), V[:, c0 + k, :])
    assert max_rel_err(delta.cpu().numpy(), ref.view(n, g * ld).numpy()) <= 1e-5


def test_relu_mask_mul():
    ops = _ops()
    n, g, d = 777, 5, 64
    act = torch.relu(torch.randn(n, d, device=DEV))
    x = torch.randn(n * g, d, device=DEV)
    ref = x * (act > 0).float().repeat_interleave(g, dim=0)
    assert torch.equal(ops.relu_mask_mul(x.clone(), act, g), ref)
    x2 = torch.randn(n * g, 7, device=DEV)                # scalar path
    act2 = torch.relu(torch.randn(n, 7, device=DEV))
    assert torch.equal(ops.relu_mask_mul(x2.clone(), act2, g), x2 * (act2 > 0).float().repeat_interleave(g, dim=0))


# ---------------------------------------------------------------------------------- SYRK (CUDA cores)
@pytest.mark.parametrize("k,n", [(1, 5), (100, 3), (2708, 16), (5000, 47), (19717, 64), (3000, 256),
                                  (2708, 1433), (50_000, 100), (33, 130)])
def test_syrk_simt_matches_fp64(k, n):
    ops = _ops()
    x = torch.randn(k, n, device=DEV)
    ref = (x.double().T @ x.double())
    c = ops.syrk(x, impl="simt")
    assert max_rel_err(c.cpu().numpy(), ref.cpu().numpy()) <= 1e-5
    assert torch.equal(c, c.T)
    c2 = ops.syrk(x, alpha=0.5, beta=1.0, out=c.clone(), impl="simt")
    assert max_rel_err(c2.cpu().numpy(), (1.5 * ref).cpu().numpy()) <= 1e-5
    # run-to-run bit reproducibility (fixed-order split-K reduction)
    assert torch.equal(ops.syrk(x, impl="simt"), c)


# ---------------------------------------------------------------------------------- end to end vs the reference goldens
@pytest.mark.parametrize("impl", ["simt"])
def test_kron_fit_marglik_match_reference_goldens(golden, impl):
    import laplace_gnn_b200 as L
    g = golden
    model = build_model(g, DEV)
    la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs={"syrk_impl": impl})
    la.fit(loader_for(g, DEV))
    check_against_golden(g, la.loss, la.H_facs.kfacs, la.log_marginal_likelihood())
    assert la.H_facs.kfacs[0][0].is_cuda


def test_column_grouping_and_ggn_mode_match_oracle():
    import laplace_gnn_b200 as L
    g = Golden("tiny_directed_3l")
    model = build_model(g, DEV)
    idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
    R = O.build_graph(g.edge_index, g.n)
    for mode in ("reference", "ggn"):
        _, ref = O.kron_factors(R, g.x, g.Ws, g.bs, g.idx, g.y, len(g.y), mode, torch.float64)
        for budget in (None, 1):   # 1 byte -> one Hessian-sqrt column per group
            be = L.B200GGN(model, "classification", hess_sqrt=mode, rhs_tile_bytes=budget, syrk_impl="simt")
            _, kron = be.kron(idx, y, N=len(y))
            for fa, fb in zip(kron.kfacs, ref):
                for a, b in zip(fa, fb):
                    assert max_rel_err(a.cpu().numpy(), b.numpy()) <= 1e-4


def test_training_step_gradients_match_dense_autograd():
    """GCNConvFunction backward vs the reference formulation (dense Â, torch autograd)."""
    g = Golden("tiny_directed_dups_2l")
    model = build_model(g, DEV)
    model.eval()
    idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
    torch.nn.functional.cross_entropy(model(idx), y).backward()
    ahat = torch.from_numpy(g.z["ahat_dense"]).double()
    Ws = [torch.from_numpy(w).double().requires_grad_(True) for w in g.Ws]
    bs = [torch.from_numpy(b).double().requires_grad_(True) for b in g.bs]
    h = torch.from_numpy(g.x).double()
    for l in range(g.L):
        h = ahat @ (h @ Ws[l].T + bs[l])
        if l < g.L - 1:
            h = torch.relu(h)
    torch.nn.functional.cross_entropy(h[torch.from_numpy(g.idx)], torch.from_numpy(g.y)).backward()
    for l, conv in enumerate(model.convs):
        assert max_rel_err(conv.lin.weight.grad.cpu().numpy(), Ws[l].grad.numpy()) <= 1e-4
        assert max_rel_err(conv.lin.bias.grad.cpu().numpy(), bs[l].grad.numpy()) <= 1e-4


def test_arxiv_shape_properties():
    """ogbn-arxiv-shaped synthetic graph (too slow for the reference): structural properties of the
    result — symmetric PSD factors, A_l = H^T H / N reproduced by torch, finite marglik — and
    agreement of the two Hessian-sqrt column groupings."""
    import laplace_gnn_b200 as L
    n, F, C, h, Lyr = 169_343, 128, 40, 256, 3
    gen = torch.Generator(device=DEV).manual_seed(0)
    src = torch.randint(0, n, (1_166_243,), device=DEV, generator=gen)
    dst = torch.randint(0, n, (1_166_243,), device=DEV, generator=gen)
    ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
    graph = L.Graph.from_edge_index(ei, n, assume_undirected=True)
    X = torch.randn(n, F, device=DEV, generator=gen)
    torch.manual_seed(0)
    model = L.SparseGCN(F, h, C, Lyr, X, graph).to(DEV)
    idx = torch.randperm(n, device=DEV, generator=gen)[: int(0.6 * n)].sort().values
    y = torch.randint(0, C, (idx.numel(),), device=DEV, generator=gen)
    be = L.B200GGN(model, "classification", syrk_impl="simt")
    loss, kron = be.kron(idx, y, N=idx.numel())
    # a budget of 7 columns: dense slabs take groups of 7, unit-compacted slabs round down to 6 (even groups)
    be2 = L.B200GGN(model, "classification", syrk_impl="simt", rhs_tile_bytes=2 * n * 256 * 4 * 7, unit_slabs=False)
    _, kron2 = be2.kron(idx, y, N=idx.numel())
    be3 = L.B200GGN(model, "classification", syrk_impl="simt", rhs_tile_bytes=2 * n * 256 * 4 * 7)
    _, kron3 = be3.kron(idx, y, N=idx.numel())
    assert be2.last_stats["group"] == 7 and be3.last_stats["group"] == 6 and be.last_stats["group"] > 7
    assert be.last_stats["unit_slabs"] > 0 and be3.last_stats["unit_slabs"] == 14 and be2.last_stats["unit_slabs"] == 0
    for fa, fb in zip(kron.kfacs, kron3.kfacs):
        for a, b in zip(fa, fb):
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4
    for fa, fb in zip(kron.kfacs, kron2.kfacs):
        for a, b in zip(fa, fb):
            assert torch.equal(a, a.T)
            assert float(torch.linalg.eigvalsh(a.double()).min()) >= -1e-4 * float(a.abs().max())
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4
    A0 = (X.double().T @ X.double() / idx.numel()).float()
    assert max_rel_err(kron.kfacs[0][1].cpu().numpy(), A0.cpu().numpy()) <= 1e-4
    assert torch.isfinite(loss)


# ---------------------------------------------------------------------------------- exact diag GGN
@pytest.mark.parametrize("name", ["tiny_undirected_2l", "tiny_directed_3l", "tiny_directed_dups_2l",
                                  "tiny_symmetrised_2l"])
def test_diag_laplace_matches_reference_golden(name):
    """hessian_structure="diag": backend.diag() vs the reference's DiagLaplace (GGNInterface.diag,
    curvature.py:412-432) stored by oracle/make_golden.py.  diag GGN <= 1e-4 rel, marglik <= 1e-3 rel."""
    import laplace_gnn_b200 as L
    from laplace_gnn_b200.diag import diag_ggn_exact
    g = Golden(name)
    if "diag_H" not in g.z.files:
        pytest.skip("fixture has no diag reference")
    model = build_model(g, DEV)
    la = L.Laplace(model, "classification", hessian_structure="diag", backend=L.B200GGN)
    la.fit(loader_for(g, DEV))
    assert max_rel_err(la.H.cpu().numpy(), g.z["diag_H"]) <= 1e-4
    ref_ml = float(g.z["diag_marglik"])
    assert abs(float(la.log_marginal_likelihood()) - ref_ml) <= 1e-3 * abs(ref_ml)
    idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
    _, d1 = diag_ggn_exact(la.backend, idx, y, tile_bytes=1)             # one train node per tile
    if g.batch_size == len(g.idx):
        assert max_rel_err(d1.cpu().numpy(), g.z["diag_H"]) <= 1e-4


# ---------------------------------------------------------------------------------- fused (gZ W) ⊙ relu'
@pytest.mark.parametrize("m,k,n,group", [(1000, 47, 256, 3), (129, 256, 256, 1), (5000, 64, 64, 5),
                                           (70_001, 256, 128, 7), (128 * 32 * 3 + 5, 200, 256, 4),
                                           (1, 8, 64, 1), (300_000, 256, 256, 12)])
def test_gemm_mask_fused_kernel_matches_fp64(m, k, n, group):
    """out = (A W) ⊙ (act > 0) on tcgen05 (3xTF32, cluster multicast) vs fp64, and vs the two-step
    cuBLAS + mask-kernel path it replaces."""
    ops = _ops()
    gen = torch.Generator(device=DEV).manual_seed(m + k + n)
    ld = (k + 3) // 4 * 4
    a = torch.randn(m, ld, device=DEV, generator=gen)[:, :k]
    w = torch.randn(k, n, device=DEV, generator=gen) / k ** 0.5
    nodes = (m + group - 1) // group
    act = torch.randn(nodes, n, device=DEV, generator=gen)
    assert ops.gemm_mask_supported(k, n)
    wp = ops.gemm_mask_prepare(w)
    out = ops.gemm_mask(a, wp, act, group)
    mask = (act > 0).repeat_interleave(group, dim=0)[:m]
    ref = (a.double() @ w.double()) * mask
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err <= 5e-6, err                              # fp32-faithful (plain TF32 would be ~5e-4)
    two_step = ops.relu_mask_mul(torch.mm(a, w), act, group) if m == nodes * group else None
    if two_step is not None:
        assert float((out - two_step).abs().max() / ref.abs().max()) <= 5e-6
    plain = ops.gemm_mask(a, wp, None, group)           # no mask
    assert float((plain.double() - a.double() @ w.double()).abs().max() / ref.abs().max()) <= 5e-6
    assert torch.equal(out, ops.gemm_mask(a, wp, act, group))   # deterministic


def test_fused_and_two_step_backends_agree():
    import laplace_gnn_b200 as L
    g = Golden("pubmed_shape")
    model = build_model(g, DEV)
    idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
    _, k1 = L.B200GGN(model, "classification", fused_gemm=True).kron(idx, y, N=len(y))
    _, k2 = L.B200GGN(model, "classification", fused_gemm=False).kron(idx, y, N=len(y))
    for fa, fb in zip(k1.kfacs, k2.kfacs):
        for a, b in zip(fa, fb):
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5


def test_marglik_training_epoch_loop_on_device():
    """The caller of the hot path (gnn/marglik_training.py:159-329): Adam step through GCNConvFunction,
    fit + marglik per epoch on the B200 backend with the input factor cached, validation forward."""
    import laplace_gnn_b200 as L
    g = Golden("cora_shape")
    model = build_model(g, DEV)
    idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
    rest = np.setdiff1d(np.arange(g.n), g.idx)
    val_idx = torch.from_numpy(rest[: len(rest) // 2]).to(DEV)
    val_y = torch.randint(0, g.C, (val_idx.numel(),), device=DEV)
    res = L.marglik_training(model, idx, y, val_idx, val_y, n_epochs=5, lr=0.01)
    assert len(res.neg_margliks) == 5 and all(np.isfinite(res.neg_margliks))
    assert res.losses[-1] < res.losses[0]
    # the last epoch's marglik equals a fresh, uncached fit of the final weights
    la = L.Laplace(model, "classification", backend=L.B200GGN)
    la.fit(L.TensorBatchLoader(idx, y))
    ref = -float(la.log_marginal_likelihood())
    assert abs(res.neg_margliks[-1] - ref) <= 1e-5 * abs(ref)


def test_masked_output_layer_spmm_changes_nothing():
    """Zeroing the edges that read non-train rows of the output-layer slab only skips gathers of rows
    that are zero anyway: identical loss and factors."""
    import laplace_gnn_b200 as L
    g = Golden("pubmed_shape")
    model = build_model(g, DEV)
    idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
    be1 = L.B200GGN(model, "classification")
    be2 = L.B200GGN(model, "classification")
    be2.skip_zero_rows = False
    l1, k1 = be1.kron(idx, y, N=len(y))
    l2, k2 = be2.kron(idx, y, N=len(y))
    assert float(l1) == float(l2)
    for fa, fb in zip(k1.kfacs, k2.kfacs):
        for a, b in zip(fa, fb):
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-6
    ops = _ops()
    keep = torch.zeros(g.n, dtype=torch.uint8, device=DEV)
    keep[idx] = 1
    masked = ops.csr_with_masked_sources(model.graph.ahat_t, keep)
    ref = model.graph.ahat_t.val * keep[model.graph.ahat_t.col.long()].float()
    assert torch.equal(masked.val, ref)


# ---------------------------------------------------------------------------------- unit-compacted slabs
def _masked_slab(n, g, h, density, seed, pitch_extra=0):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    act = torch.randn(n, h, device=DEV, generator=gen)
    act *= (torch.rand(n, h, device=DEV, generator=gen) < density)          # <= 0 <-> dead unit
    act[3] = 0                                                              # a node with no live unit
    if n > 5:
        act[5] = 1                                                          # ... and one with all of them
    slab = torch.randn(n, g * h + pitch_extra, device=DEV, generator=gen)
    slab[:, : g * h].view(n, g, h).mul_((act > 0)[:, None, :])
    return slab, act


@pytest.mark.parametrize("g,h", [(4, 32), (8, 64), (12, 256), (16, 256), (12, 96), (4, 1024)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_unit_pack_layout(g, h, density):
    """lgnn_unit_pack_f32: header words and the in-place [slot][g] layout, against a numpy restatement."""
    ops = _ops()
    n = 257
    slab, act = _masked_slab(n, g, h, density, seed=g * 1000 + h, pitch_extra=4)
    dense = slab[:, : g * h].view(n, g, h).cpu().numpy().copy()
    live = (act > 0).cpu().numpy()
    us = ops.unit_pack(slab, act, g)
    hdr = us.hdr.cpu().numpy().view(np.uint32)
    out = slab.cpu().numpy()
    for r in range(n):
        units = np.nonzero(live[r])[0]
        first = 0
        for w in range(h // 32):
            blk = live[r, 32 * w: 32 * w + 32]
            mask = int(sum(1 << i for i in range(32) if blk[i]))
            assert hdr[r, w, 0] == mask and hdr[r, w, 1] == first
            first += int(blk.sum())
        want = dense[r][:, units].T.reshape(-1)                   # [slot][g]
        assert np.array_equal(out[r, : want.size], want)
    assert us.live_units().cpu().numpy().tolist() == live.sum(1).tolist()


@pytest.mark.parametrize("g,h", [(4, 32), (8, 64), (12, 256), (16, 256), (12, 96), (8, 512)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_unit_spmm_is_bit_identical_to_dense(g, h, density):
    """unit_pack + spmm_units against the dense SpMM of the same masked slab: identical bits (same
    neighbour order, dead units are the dense kernel's multiplications by zero); every unroll variant."""
    ops = _ops()
    n = 5000
    ei = O.synthetic_edges(n, 40_000, seed=g + h)
    G = _dev_graph(ei, n)
    slab, act = _masked_slab(n, g, h, density, seed=g * 7 + h, pitch_extra=8)
    dense = ops.spmm(G.ahat, slab, d=g * h, impl="ldg")
    us = ops.unit_pack(slab, act, g)
    for variant in [0, 1] + list(range(8, 16)):      # default, simple kernel, every staged configuration
        y = ops.spmm_units(G.ahat, us, variant=variant)
        assert torch.equal(y, dense), variant
    # rows of a directed pattern (Â^T != Â) and an output with a wider pitch
    out = torch.full((n, g * h + 12), -1.0, device=DEV)
    ops.spmm_units(G.ahat_t, us, out=out)
    assert torch.equal(out[:, : g * h], ops.spmm(G.ahat_t, _unpack(us), impl="ldg"))
    assert bool((out[:, g * h:] == -1).all())


def _unpack(us):
    """Dense [n, g*h] slab of a UnitSlab (test helper, torch on the device)."""
    n, g, h = us.n_rows, us.g, us.h
    live = us.act[:, :h] > 0
    slot = torch.cumsum(live, 1) - 1                                        # slot of each live unit
    comp = us.slab[:, : g * h].reshape(n, h, g)                             # [slot][g] (tail unused)
    vals = torch.gather(comp, 1, slot.clamp(min=0)[:, :, None].expand(n, h, g))
    return (vals * live[:, :, None]).permute(0, 2, 1).reshape(n, g * h).contiguous()


def test_unit_slab_arguments_are_checked():
    ops = _ops()
    from laplace_gnn_b200._lib import LgnnError
    assert ops.unit_slabs_supported(12, 256) and ops.unit_slabs_supported(16, 1024)
    assert not ops.unit_slabs_supported(11, 256) and not ops.unit_slabs_supported(12, 100)
    assert not ops.unit_slabs_supported(20, 256) and not ops.unit_slabs_supported(4, 2048)
    slab, act = _masked_slab(64, 4, 48, 0.5, seed=1)                        # h = 48: not a multiple of 32
    with pytest.raises(LgnnError):
        ops.unit_pack(slab, act, 4, hdr=torch.empty(64, 2, 2, dtype=torch.int32, device=DEV))


def test_unit_compacted_backward_gives_the_same_factors():
    """The KFAC backward with unit-compacted slabs below the output layer against the dense slabs:
    same loss, factors equal to rounding of the SYRK's zero-padded column groups, reference goldens."""
    import laplace_gnn_b200 as L
    for name in ("pubmed_shape", "tiny_directed_3l"):
        g = Golden(name)
        model = build_model(g, DEV)
        idx, y = torch.from_numpy(g.idx).to(DEV), torch.from_numpy(g.y).to(DEV)
        be1 = L.B200GGN(model, "classification", unit_slabs=True, unit_min_width=0)
        be2 = L.B200GGN(model, "classification", unit_slabs=False)
        l1, k1 = be1.kron(idx, y, N=len(y))
        l2, k2 = be2.kron(idx, y, N=len(y))
        if g.h % 32 == 0:
            assert be1.last_stats["unit_slabs"] > 0, name
        assert float(l1) == float(l2)
        for fa, fb in zip(k1.kfacs, k2.kfacs):
            for a, b in zip(fa, fb):
                assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5
        check_against_golden(g, l1, k1.kfacs, torch.tensor(g.marglik))


@pytest.mark.parametrize("h,C,layers", [(64, 10, 3), (256, 47, 3), (128, 6, 2)])
def test_unit_compacted_backward_products_like_shapes(h, C, layers):
    """Same comparison on synthetic models whose hidden widths take the fused tcgen05 GEMM (no mask in
    its epilogue when unit_pack follows) and whose class count needs a zero-padded last group."""
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
    res = []
    for units in (True, False):
        be = L.B200GGN(model, "classification", unit_slabs=units, unit_min_width=0)
        loss, kron = be.kron(idx, y, N=len(y))
        assert (be.last_stats["unit_slabs"] > 0) == units
        res.append((float(loss), kron.kfacs))
    assert res[0][0] == res[1][0]
    for fa, fb in zip(res[0][1], res[1][1]):
        for a, b in zip(fa, fb):            # same slabs bit for bit; the column grouping changes the SYRK's summation order
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5


# ---------------------------------------------------------------------------------- linear layer on the fused GEMM
@pytest.mark.parametrize("d_in,d_out,bias", [(100, 256, True), (256, 256, True), (256, 47, True), (128, 40, False),
                                              (16, 7, True), (3, 64, True), (255, 129, True)])
def test_linear_layer_on_the_tensor_core_gemm(d_in, d_out, bias):
    """ops.gemm_bias (lgnn_gemm_bias_f32: Z = H W^T + b of gnn/models/layers.py:45 on the 3xTF32 tcgen05 kernel, bias
    in the epilogue, output width zero-padded to 64 / 128 / 256) against float64; pitched input, row tail."""
    ops = _ops()
    m = 70_001
    gen = torch.Generator(device=DEV).manual_seed(d_in + d_out)
    h = torch.randn(m, d_in + 4, device=DEV, generator=gen)[:, :d_in] if d_in % 4 == 0 else \
        torch.randn(m, (d_in + 3) // 4 * 4, device=DEV, generator=gen)[:, :d_in]
    lin = torch.nn.Linear(d_in, d_out, bias=bias).to(DEV)
    wp = ops.linear_prepare(lin.weight)
    assert wp is not None and wp.n in (64, 128, 256) and wp.n >= d_out and wp.k == d_in
    bp = None
    if bias:
        bp = torch.zeros(wp.n, device=DEV)
        bp[:d_out] = lin.bias.detach()
    out = torch.full((m, wp.n + 8), -3.0, device=DEV)
    ops.gemm_bias(h, wp, bp, out[:, : wp.n])
    want = h.double() @ lin.weight.detach().double().t() + (lin.bias.detach().double() if bias else 0.0)
    assert max_rel_err(out[:, :d_out].cpu().numpy(), want.cpu().numpy()) <= 5e-6          # 3xTF32: fp32-level
    assert bool((out[:, d_out:wp.n] == 0).all()) and bool((out[:, wp.n:] == -3.0).all())
    assert ops.linear_prepare(torch.randn(7, 1433, device=DEV)) is None          # d_in beyond the kernel: cuBLAS path
