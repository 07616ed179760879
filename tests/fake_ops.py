"""TEST DOUBLE for ``laplace_gnn_b200.ops`` — CPU implementations built on the oracle, with the
same signatures.  Installed only by the ``fake_ops`` fixture (tests/conftest.py) so that the
host-side logic (backend orchestration, Kron packing, Laplace drivers, the partitioned
multi-rank path under gloo) can be exercised on a machine without a GPU.  Never shipped, never
imported by the package.
"""
from __future__ import annotations

import numpy as np
import torch

from laplace_gnn_b200.ops import CSR
from oracle import gcn_kfac_oracle as O


def _np(t):
    return t.detach().cpu().numpy()


def csr_from_edge_index(edge_index, num_nodes, symmetric=False):
    rp, col = O.coo_to_adj_csr(_np(edge_index), num_nodes, symmetric)
    return CSR(num_nodes, num_nodes, torch.from_numpy(rp), torch.from_numpy(col), None)


def csr_transpose(a):
    rp, col, _ = O.csr_transpose_pattern(_np(a.rowptr), _np(a.col), a.n_cols)
    return CSR(a.n_cols, a.n_rows, torch.from_numpy(rp), torch.from_numpy(col), None)


def degree_norm(a):
    deg = np.diff(_np(a.rowptr)).astype(np.int64)
    return torch.from_numpy(deg), torch.from_numpy(O.inv_sqrt_degree(deg))


def edge_values(a, dis, row_offset=0):
    rows = np.repeat(np.arange(a.n_rows, dtype=np.int64), np.diff(_np(a.rowptr))) + row_offset
    d = _np(dis)
    return torch.from_numpy((d[rows] * d[_np(a.col)]).astype(np.float32))


def row_partition(rowptr, nparts):
    return torch.from_numpy(O.row_partition(_np(rowptr), nparts))


def halo_columns(a, lo, hi):
    return torch.from_numpy(O.halo_columns(_np(a.rowptr), _np(a.col), lo, hi))


def csr_slice_remap(a, lo, hi, bounds, pad):
    rp, col = _np(a.rowptr), _np(a.col).astype(np.int64)
    b = _np(bounds)
    s, e = rp[lo], rp[hi]
    c = col[s:e]
    owner = np.searchsorted(b, c, side="right") - 1
    new_col = (owner * pad + (c - b[owner])).astype(np.int32)
    val = None if a.val is None else a.val[s:e].clone()
    return CSR(hi - lo, (len(b) - 1) * pad, torch.from_numpy(rp[lo:hi + 1] - s), torch.from_numpy(new_col), val,
               a.max_row_nnz)


def spmm(a, x, relu=False, out=None, d=None, impl="auto"):
    d = x.shape[1] if d is None else d
    m = torch.sparse_csr_tensor(a.rowptr, a.col.to(torch.int64), a.val, size=(a.n_rows, x.shape[0]))
    y = m @ x[:, :d].contiguous()
    if relu:
        y = torch.relu(y)
    if out is None:
        return y
    out[: a.n_rows, :d] = y
    return out


def softmax_ce_sum(logits, idx, y, C=None):
    C = logits.shape[1] if C is None else C
    f = logits[idx][:, :C]
    loss = torch.nn.functional.cross_entropy(f, y, reduction="sum").double()
    hits = (f.argmax(1) == y).sum()
    return loss, hits


def hess_rhs(logits, idx, c0, ncols, delta, ldc, mode="reference", C=None):
    C = logits.shape[1] if C is None else C
    if idx.numel() == 0 or delta.shape[0] == 0:            # a rank without train nodes / without rows
        return delta
    V = O.hess_sqrt_rhs(logits[idx][:, :C], mode)          # [m, C(col), C]
    d3 = delta.view(delta.shape[0], -1, ldc)             # a wider row = zero padding columns behind the group
    for g in range(ncols):
        d3[:, g, :C].index_add_(0, idx, V[:, c0 + g, :])
    return delta


def spmm_hess_supported(C, width):
    return 1 <= C <= 64 and 1 <= width <= 16


_STATS_MODE = ["reference"]


def hess_stats(logits, idx, mode="reference", C=None, out=None):
    """CPU double: in place of the five softmax vectors the rows keep what spmm_hess needs to rebuild the right-hand
    sides — the logits and how often the node occurs in the batch (same [n, 5 Cp] shape, so the rows layout can
    all-gather it like the real thing)."""
    C = logits.shape[1] if C is None else C
    cp = (C + 3) // 4 * 4
    n = logits.shape[0]
    if out is None:
        out = torch.zeros(n, 5 * cp, dtype=torch.float32)
    out.zero_()
    out[:n, :C] = logits[:, :C]
    out[:n, cp].index_add_(0, idx, torch.ones(idx.numel(), dtype=torch.float32))
    _STATS_MODE[0] = mode
    return out


def spmm_hess(a, stats, C, c0, ncols, width, out=None, staged=None):
    cp = (C + 3) // 4 * 4
    n = a.n_cols
    idx = torch.repeat_interleave(torch.arange(n), stats[:n, cp].round().to(torch.int64))
    delta = torch.zeros(n, width * cp, dtype=torch.float32)
    hess_rhs(stats[:n, :C].contiguous(), idx, c0, ncols, delta, cp, _STATS_MODE[0], C)
    return spmm(a, delta, out=out)


def relu_mask_mul(inp, act, group, out=None, d=None):
    d = inp.shape[1] if d is None else d
    out = inp if out is None else out
    n = act.shape[0]
    mask = (act[:, :d] > 0).to(inp.dtype).repeat_interleave(group, dim=0)
    out[: n * group, :d] = inp[: n * group, :d] * mask
    return out


def syrk(x, n=None, alpha=1.0, beta=0.0, out=None, impl="auto", k_rows=None):
    n = x.shape[1] if n is None else n
    k_rows = x.shape[0] if k_rows is None else k_rows
    xs = x[:k_rows, :n]
    c = alpha * (xs.T @ xs)
    if out is None:
        return c
    out[:n, :n] = beta * out[:n, :n] + c if beta != 0.0 else c
    return out


def csr_with_masked_sources(a, keep):
    val = a.val * keep[a.col.to(torch.int64)].to(a.val.dtype)
    return CSR(a.n_rows, a.n_cols, a.rowptr, a.col, val, a.max_row_nnz, masked=True)


def unit_slabs_supported(g, h):
    return 2 <= g <= 16 and g % 2 == 0 and 32 <= h <= 1024 and h % 32 == 0


def unit_pack(slab, act, g, hdr=None):
    """Same in-place [live unit slot][g] layout as csrc/spmm_units.cu (header words left to the GPU test)."""
    from laplace_gnn_b200.ops import UnitSlab
    n, h = act.shape
    live = act > 0
    dense = slab[:n, : g * h].reshape(n, g, h).clone()
    order = torch.argsort((~live).to(torch.int8), dim=1, stable=True)       # live units first, ascending
    comp = torch.gather(dense.permute(0, 2, 1), 1, order[:, :, None].expand(n, h, g))   # [n, slot, g]
    slab[:n, : g * h] = comp.reshape(n, g * h)           # slots beyond the live count hold garbage, never read
    return UnitSlab(n, g, h, slab, hdr if hdr is not None else torch.zeros(n, h // 32, 2, dtype=torch.int32), act)


def _i32_bits(x):
    """int64 values in [0, 2^32) -> the int32 holding the same 32 bits (what the device writes as uint32)."""
    return ((x + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32)


def unit_pack_ragged(src, act, g, row_first, dst, hdr):
    """csrc/spmm_units.cu's ragged layout, header words included: node n's slots start at absolute slot
    row_first[n] of the flat ``dst``; hdr[n][w] = (mask, absolute slot of block w's first live unit)."""
    n, h = act.shape
    nb = h // 32
    blk = (act > 0).view(n, nb, 32)
    cnt = blk.sum(2)
    slots = cnt if g % 4 == 0 else (cnt + 1) // 2 * 2
    first = row_first[:n, None] + torch.cumsum(slots, 1) - slots
    if hdr is not None:
        hv = hdr.view(-1, nb, 2)
        hv[:n, :, 0] = _i32_bits((blk.to(torch.int64) << torch.arange(32)).sum(2))
        hv[:n, :, 1] = _i32_bits(first)
    if src is not None:
        slot = (first[:, :, None] + torch.cumsum(blk, 2) - 1).view(n, h)
        rows, units = torch.nonzero(blk.view(n, h), as_tuple=True)
        dense = src[:n, : g * h].reshape(n, g, h)
        at = slot[rows, units][:, None] * g + torch.arange(g)
        dst.view(-1)[at.reshape(-1)] = dense[rows, :, units].reshape(-1)


def _ragged_to_dense(us):
    n, g, h = us.n_rows, us.g, us.h
    nb = h // 32
    hv = us.hdr.view(-1, nb, 2)[:n].to(torch.int64) & 0xFFFFFFFF
    blk = ((hv[:, :, 0, None] >> torch.arange(32)) & 1).bool()
    slot = (hv[:, :, 1, None] + torch.cumsum(blk, 2) - 1).view(n, h)
    live = blk.view(n, h)
    dense = torch.zeros(n, g, h, dtype=us.slab.dtype)
    rows, units = torch.nonzero(live, as_tuple=True)
    at = slot[rows, units][:, None] * g + torch.arange(g)
    dense[rows, :, units] = us.slab.view(-1)[at.reshape(-1)].view(-1, g)
    return dense.reshape(n, g * h)


def spmm_units(a, us, out=None, variant=0):
    if us.ragged:
        return spmm(a, _ragged_to_dense(us), out=out)
    n, g, h = us.n_rows, us.g, us.h
    live = us.act[:, :h] > 0
    slot = (torch.cumsum(live, 1) - 1).clamp(min=0)
    comp = us.slab[:n, : g * h].reshape(n, h, g)
    dense = (torch.gather(comp, 1, slot[:, :, None].expand(n, h, g)) * live[:, :, None]).permute(0, 2, 1)
    return spmm(a, dense.reshape(n, g * h).contiguous(), out=out)


def sddmm(rows, cols, u, v, d=None, out=None, accumulate=False):
    d = min(u.shape[1], v.shape[1]) if d is None else d
    s = (u[rows.long(), :d] * v[cols.long(), :d]).sum(1)
    if out is None:
        return s
    out[: s.numel()] = out[: s.numel()] + s if accumulate else s
    return out


def gemm_mask_supported(k, n):
    return False          # the CPU double keeps the two-step path (torch.mm + relu_mask_mul)


ALL = ["spmm_hess_supported", "hess_stats", "spmm_hess", "sddmm", "unit_slabs_supported", "unit_pack", "unit_pack_ragged", "spmm_units", "gemm_mask_supported", "csr_with_masked_sources", "csr_from_edge_index", "csr_transpose", "degree_norm", "edge_values", "row_partition",
       "halo_columns", "csr_slice_remap", "spmm", "softmax_ce_sum", "hess_rhs", "relu_mask_mul", "syrk"]
