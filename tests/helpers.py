"""Shared helpers: build the sparse model from a golden fixture on a given device."""
import numpy as np
import torch


def build_model(g, device="cpu", assume_undirected=False):
    import laplace_gnn_b200 as L
    ei = torch.from_numpy(g.edge_index).to(device)
    graph = L.Graph.from_edge_index(ei, g.n, symmetric=g.symmetric, assume_undirected=assume_undirected)
    X = torch.from_numpy(g.x).to(device)
    model = L.SparseGCN(g.F, g.h, g.C, g.L, X, graph).to(device)
    with torch.no_grad():
        for l, conv in enumerate(model.convs):
            conv.lin.weight.copy_(torch.from_numpy(g.Ws[l]))
            conv.lin.bias.copy_(torch.from_numpy(g.bs[l]))
    return model


def loader_for(g, device="cpu"):
    from torch.utils.data import DataLoader, TensorDataset
    ds = TensorDataset(torch.from_numpy(g.idx).to(device), torch.from_numpy(g.y).to(device))
    return DataLoader(ds, batch_size=g.batch_size, shuffle=False)


def check_against_golden(g, loss, kfacs, marglik, fac_tol=1e-4, ml_tol=1e-3):
    """north-star tolerances: Kronecker factors <= 1e-4 rel, log marglik <= 1e-3 rel."""
    from conftest import max_rel_err
    assert len(kfacs) == len(g.kfacs)
    for blk, ref_blk in zip(kfacs, g.kfacs):
        assert len(blk) == len(ref_blk)
        for h, ref in zip(blk, ref_blk):
            err = max_rel_err(h.detach().cpu().numpy(), ref)
            assert err <= fac_tol, f"{g.name}: factor rel err {err}"
    assert abs(float(loss) - g.loss) <= 1e-4 * abs(g.loss)
    assert abs(float(marglik) - g.marglik) <= ml_tol * abs(g.marglik)


def unit_layout(live_row: np.ndarray, g: int):
    """Numpy restatement of the in-place unit-compacted row layout (csrc/spmm_units.cu, spmm_units_even.cu).

    live_row: bool [h].  Returns (hdr, slot): hdr[w] = (mask, first slot of block w), slot[u] = slot of live
    unit u (-1 for dead units).  For g % 4 == 0 the runs of the 32-unit blocks are packed back to back; for
    g % 4 == 2 (8*odd bytes per slot) every block starts on an EVEN slot so that its run is 16-byte aligned."""
    h = live_row.shape[0]
    even_start = g % 4 != 0
    hdr, slot = [], np.full(h, -1, dtype=np.int64)
    first = 0
    for w in range(h // 32):
        blk = live_row[32 * w: 32 * w + 32]
        mask = int(sum(1 << i for i in range(32) if blk[i]))
        hdr.append((mask, first))
        k = 0
        for i in range(32):
            if blk[i]:
                slot[32 * w + i] = first + k
                k += 1
        first += (k + 1) // 2 * 2 if even_start else k
    return hdr, slot
