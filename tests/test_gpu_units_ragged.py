"""GPU tests of the RAGGED unit-compacted rows (lgnn_unit_pack_ragged_f32 / lgnn_spmm_units_ragged_f32): the layout
the ranks of the row-partitioned backward exchange (``B200GGN(backward_parallel="rows", unit_rows=True)``) — rows
back to back at absolute slots, header words carrying absolute slots, the SpMM kernels of the pitched layout run with
pitch 0.  Bit-exact against the dense SpMM and against the in-place pack."""
import numpy as np
import pytest
import torch

from helpers import unit_layout
from oracle import gcn_kfac_oracle as O
from test_gpu_units_even import _masked_slab

pytestmark = [pytest.mark.gpu]

DEV = "cuda:0"


def _plan(act, g, base, shuffle_seed=None):
    """row_first for rows laid back to back from slot ``base`` on — optionally in a shuffled row order, which the
    layout allows (a row is found through its header alone)."""
    from laplace_gnn_b200 import ops
    slots = ops.unit_row_slots(act, g)
    n = slots.numel()
    order = torch.arange(n, device=DEV)
    if shuffle_seed is not None:
        order = torch.randperm(n, generator=torch.Generator().manual_seed(shuffle_seed)).to(DEV)
    s = slots[order]
    first = torch.empty(n, dtype=torch.int64, device=DEV)
    first[order] = torch.cumsum(s, 0) - s + base
    return first.contiguous(), int(slots.sum().item())


@pytest.mark.parametrize("g,h", [(2, 32), (4, 64), (6, 256), (8, 128), (10, 256), (12, 256), (14, 96), (16, 256),
                                 (16, 1024)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_ragged_pack_layout_and_headers(g, h, density):
    from laplace_gnn_b200 import ops
    n, base = 301, 40
    slab, act = _masked_slab(n, g, h, density, seed=g * 100 + h, pitch_extra=4)
    dense = slab[:, : g * h].view(n, g, h).cpu().numpy().copy()
    live = (act > 0).cpu().numpy()
    first, total = _plan(act, g, base, shuffle_seed=g + h)
    flat = torch.full(((base + total) * g + 8,), -7.0, device=DEV)
    hdr = torch.zeros(n, h // 32, 2, dtype=torch.int32, device=DEV)
    ops.unit_pack_ragged(slab, act, g, first, flat, hdr)
    hdr_only = torch.zeros_like(hdr)
    ops.unit_pack_ragged(None, act, g, first, None, hdr_only)                 # header-only call: same words
    assert torch.equal(hdr, hdr_only)
    flat2 = torch.full_like(flat, -7.0)
    ops.unit_pack_ragged(slab, act, g, first, flat2, None)                    # value-only call: same values
    assert torch.equal(flat, flat2)
    assert torch.equal(slab[:, : g * h].view(n, g, h).cpu(), torch.from_numpy(dense))   # src untouched
    H, F, R0 = hdr.cpu().numpy().view(np.uint32), flat.cpu().numpy(), first.cpu().numpy()
    written = np.zeros(F.shape[0], dtype=bool)
    for r in range(n):
        want_hdr, slot = unit_layout(live[r], g)
        for w, (mask, rel) in enumerate(want_hdr):
            assert H[r, w, 0] == mask and H[r, w, 1] == R0[r] + rel
        for u in np.nonzero(live[r])[0]:
            at = (R0[r] + slot[u]) * g
            assert np.array_equal(F[at: at + g], dense[r][:, u])
            written[at: at + g] = True
    assert np.all(F[~written] == -7.0)             # nothing outside the rows' own slots is touched
    assert written[: base * g].sum() == 0


@pytest.mark.parametrize("g,h", [(2, 64), (4, 64), (6, 256), (8, 256), (10, 256), (12, 256), (14, 256), (16, 256),
                                 (16, 512)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_ragged_unit_spmm_is_bit_identical_to_dense(g, h, density):
    from laplace_gnn_b200 import ops
    import laplace_gnn_b200 as L
    n = 5000
    ei = O.synthetic_edges(n, 40_000, seed=g + h)
    G = L.Graph.from_edge_index(torch.from_numpy(ei).to(DEV), n)
    slab, act = _masked_slab(n, g, h, density, seed=g * 7 + h, pitch_extra=8)
    dense = ops.spmm(G.ahat, slab, d=g * h, impl="ldg")
    dense_t = ops.spmm(G.ahat_t, slab, d=g * h, impl="ldg")
    for base, seed in ((0, None), (1000, 3)):                                # in row order / shuffled behind a gap
        first, total = _plan(act, g, base, seed)
        flat = torch.zeros((base + total) * g + 4, device=DEV)
        hdr = torch.zeros(n, h // 32, 2, dtype=torch.int32, device=DEV)
        ops.unit_pack_ragged(slab, act, g, first, flat, hdr)
        us = ops.UnitSlab(n, g, h, flat, hdr, act, ragged=True)
        variants = [0, 1, 8, 12, 13, 14] + ([16, 28] if g % 4 else [])
        for variant in variants:
            assert torch.equal(ops.spmm_units(G.ahat, us, variant=variant), dense), (variant, base)
        out = torch.full((n, g * h + 12), -1.0, device=DEV)
        ops.spmm_units(G.ahat_t, us, out=out)
        assert torch.equal(out[:, : g * h], dense_t)
        assert bool((out[:, g * h:] == -1).all())


def test_ragged_entry_points_reject_bad_arguments():
    from laplace_gnn_b200 import ops
    from laplace_gnn_b200._lib import LgnnError
    n, g, h = 64, 4, 64
    slab, act = _masked_slab(n, g, h, 0.5, seed=1)
    first, total = _plan(act, g, 0)
    flat = torch.zeros(total * g + 4, device=DEV)
    hdr = torch.zeros(n, h // 32, 2, dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        ops.unit_pack_ragged(slab, act, g, first.to(torch.int32), flat, hdr)          # row_first must be int64
    with pytest.raises(ValueError):
        ops.unit_pack_ragged(slab, act, g, first, None, hdr)                          # values without a destination
    with pytest.raises(LgnnError):
        ops.unit_pack_ragged(None, act, g, first, None, None)                         # nothing asked for
    with pytest.raises(LgnnError):
        ops.unit_pack_ragged(slab, act, 3, first, flat, hdr)                          # odd group
