"""GPU tests of the unit-compacted slabs for column groups with g % 4 == 2 (csrc/spmm_units_even.cu: the kernels
behind ``B200GGN(unit_even_groups=True)``, the default since round 2 — a rank of the 8-GPU column split owns 6 of
the products shape's 47 columns), of ``unit_hub_split=True`` (hub rows cut into pieces for the unit SpMM), and the GPU
leg of the differential fuzz against the oracle."""
import numpy as np
import pytest
import torch

from conftest import max_rel_err
from helpers import unit_layout
from oracle import gcn_kfac_oracle as O

pytestmark = [pytest.mark.gpu]

DEV = "cuda:0"


def _masked_slab(n, g, h, density, seed, pitch_extra=0):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    act = torch.randn(n, h, device=DEV, generator=gen)
    act *= (torch.rand(n, h, device=DEV, generator=gen) < density)
    act[3] = 0                                                              # a node with no live unit
    if n > 5:
        act[5] = 1                                                          # ... and one with all of them
    slab = torch.randn(n, g * h + pitch_extra, device=DEV, generator=gen)
    slab[:, : g * h].view(n, g, h).mul_((act > 0)[:, None, :])
    return slab, act


@pytest.mark.parametrize("g,h", [(2, 32), (6, 64), (6, 256), (10, 256), (14, 256), (6, 96), (2, 1024)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_unit_pack_even_layout(g, h, density):
    """Header words and the in-place [slot][g] layout with every block's run on an even slot."""
    from laplace_gnn_b200 import ops
    n = 257
    slab, act = _masked_slab(n, g, h, density, seed=g * 1000 + h, pitch_extra=4)
    dense = slab[:, : g * h].view(n, g, h).cpu().numpy().copy()
    live = (act > 0).cpu().numpy()
    us = ops.unit_pack(slab, act, g)
    hdr = us.hdr.cpu().numpy().view(np.uint32)
    out = slab.cpu().numpy()
    for r in range(n):
        want_hdr, slot = unit_layout(live[r], g)
        for w, (mask, first) in enumerate(want_hdr):
            assert hdr[r, w, 0] == mask and hdr[r, w, 1] == first and first % 2 == 0
        for u in np.nonzero(live[r])[0]:
            assert np.array_equal(out[r, slot[u] * g: slot[u] * g + g], dense[r][:, u])


@pytest.mark.parametrize("g,h", [(2, 32), (6, 64), (6, 256), (10, 256), (14, 256), (6, 96), (10, 512)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_unit_spmm_even_is_bit_identical_to_dense(g, h, density):
    from laplace_gnn_b200 import ops
    import laplace_gnn_b200 as L
    n = 5000
    ei = O.synthetic_edges(n, 40_000, seed=g + h)
    G = L.Graph.from_edge_index(torch.from_numpy(ei).to(DEV), n)
    slab, act = _masked_slab(n, g, h, density, seed=g * 7 + h, pitch_extra=8)
    masked = slab.clone()
    dense = ops.spmm(G.ahat, masked, d=g * h, impl="ldg")
    dense_t = ops.spmm(G.ahat_t, masked, d=g * h, impl="ldg")
    us = ops.unit_pack(slab, act, g)
    for variant in [0, 1, 2] + list(range(8, 16)) + [16, 17, 24, 28]:   # 16+: one unit block per warp forced
        y = ops.spmm_units(G.ahat, us, variant=variant)
        assert torch.equal(y, dense), variant
    out = torch.full((n, g * h + 12), -1.0, device=DEV)
    ops.spmm_units(G.ahat_t, us, out=out)
    assert torch.equal(out[:, : g * h], dense_t)
    assert bool((out[:, g * h:] == -1).all())


@pytest.mark.parametrize("h,C,layers", [(64, 10, 3), (256, 6, 3), (256, 5, 2), (128, 14, 2)])
def test_even_groups_give_the_same_factors(h, C, layers):
    """B200GGN(unit_even_groups=True) against dense slabs: same loss, factors equal to SYRK rounding."""
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
    be1 = L.B200GGN(model, "classification", unit_slabs=True, unit_min_width=0, unit_even_groups=True)
    be2 = L.B200GGN(model, "classification", unit_slabs=False)
    l1, k1 = be1.kron(idx, y, N=len(y))
    l2, k2 = be2.kron(idx, y, N=len(y))
    assert be1.last_stats["unit_slabs"] > 0 and be1.last_stats["group"] % 2 == 0
    if C in (5, 6, 10, 14):
        assert be1.last_stats["group"] % 4 == 2          # the group really took the even-g kernels
    assert float(l1) == float(l2)
    for fa, fb in zip(k1.kfacs, k2.kfacs):
        for a, b in zip(fa, fb):
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("limit", [64, 1000])
def test_hub_split_gives_the_same_factors(limit):
    """B200GGN(unit_hub_split=True) on an R-MAT graph whose hub rows exceed the (lowered) row limit of the unit
    SpMM: pieces as extra output rows + the gather SpMM, against dense slabs."""
    import laplace_gnn_b200 as L
    n, C, h = 6000, 12, 256
    ei = torch.from_numpy(O.synthetic_edges(n, 120_000, seed=11, rmat=True, directed=True)).to(DEV)
    graph = L.Graph.from_edge_index(ei, n)
    assert graph.ahat_t.max_row_nnz > 1000
    gen = torch.Generator().manual_seed(2)
    X = torch.randn(n, 24, generator=gen).to(DEV)
    torch.manual_seed(2)
    model = L.SparseGCN(24, h, C, 3, X, graph).to(DEV)
    idx = torch.randperm(n, generator=gen)[: int(0.6 * n)].sort().values.to(DEV)
    y = torch.randint(0, C, (idx.numel(),), generator=gen).to(DEV)
    be0 = L.B200GGN(model, "classification", unit_slabs=False)
    l0, k0 = be0.kron(idx, y, N=len(y))
    be1 = L.B200GGN(model, "classification", unit_hub_split=True)
    be1.unit_row_limit = limit
    l1, k1 = be1.kron(idx, y, N=len(y))
    assert be1.last_stats["unit_slabs"] > 0 and graph.meta["_split_t"][limit].n_extra > 0
    assert float(l0) == float(l1)
    for fa, fb in zip(k1.kfacs, k0.kfacs):
        for a, b in zip(fa, fb):
            assert max_rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5


# ---------------------------------------------------------------------------------- on-the-fly Hessian-sqrt SpMM
@pytest.mark.parametrize("seed", list(range(12)))
def test_random_small_configurations_match_the_oracle(seed):
    """The default backend on random small shapes (directed / symmetrised graphs, duplicate and self-loop edges,
    isolated nodes, 1-4 layers, 2-9 classes, hidden widths 1-96 incl. multiples of 32, uneven batches, repeated
    train nodes) against the oracle — the GPU leg of oracle/fuzz_against_reference.py.  Behind LGNN_LAB until its
    first run on a B200; then it joins the default suite."""
    import laplace_gnn_b200 as L
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(6, 400))
    directed = bool(rng.integers(0, 2))
    symmetric = bool(directed and rng.integers(0, 3) == 0)
    e = int(rng.integers(0, 6 * n))
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    if not directed:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    if rng.integers(0, 2) and e > 0:
        iso = rng.permutation(n)[: max(1, n // 8)]
        keep = ~(np.isin(src, iso) | np.isin(dst, iso))
        src, dst = src[keep], dst[keep]
    ei = np.stack([src, dst]).astype(np.int64)
    layers, C, F = int(rng.integers(1, 5)), int(rng.integers(2, 10)), int(rng.integers(1, 40))
    h = int(rng.choice([1, 3, 8, 32, 33, 64, 96]))
    m = int(rng.integers(1, n + 1))
    idx = np.sort(rng.permutation(n)[:m]).astype(np.int64)
    if rng.integers(0, 4) == 0 and m > 1:
        idx = np.concatenate([idx, idx[:2]])
    y = rng.integers(0, C, idx.shape[0]).astype(np.int64)
    bs = int(idx.shape[0]) if rng.integers(0, 2) else int(rng.integers(1, idx.shape[0] + 1))
    x = rng.standard_normal((n, F)).astype(np.float32)
    graph = L.Graph.from_edge_index(torch.from_numpy(ei).to(DEV), n, symmetric=symmetric)
    torch.manual_seed(seed)
    model = L.SparseGCN(F, h, C, layers, torch.from_numpy(x).to(DEV), graph).to(DEV)
    Ws = [c.lin.weight.detach().cpu().numpy() for c in model.convs]
    bs_ = [c.lin.bias.detach().cpu().numpy() for c in model.convs]
    from torch.utils.data import DataLoader, TensorDataset
    la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs={"unit_min_width": 0})
    la.fit(DataLoader(TensorDataset(torch.from_numpy(idx).to(DEV), torch.from_numpy(y).to(DEV)), batch_size=bs))
    loss, kfacs, ml = O.fit_and_marglik(O.build_graph(ei, n, symmetric), x, Ws, bs_, idx, y, 1.0, "reference",
                                        torch.float64, None if bs == len(idx) else bs)
    for blk, ref_blk in zip(la.H_facs.kfacs, kfacs):
        for a, b in zip(blk, ref_blk):
            assert float((a.cpu().double() - b).abs().max()) <= 1e-4 * max(float(b.abs().max()), 1e-12)
    assert abs(float(la.loss) - float(loss)) <= 1e-4 * abs(float(loss))
    assert abs(float(la.log_marginal_likelihood()) - float(ml)) <= 1e-3 * abs(float(ml))
