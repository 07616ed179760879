"""Exact diagonal GGN for ``hessian_structure="diag"`` (the ``CurvatureInterface.diag`` contract,
laplace/curvature/curvature.py:267-289; reference implementation ``GGNInterface.diag``,
curvature.py:412-432, with the true Λ = diag(p) - p p^T of :365-372 — no T1 quirk on this path).

The reference materialises the (M, C, P) Jacobian with ``torch.func.jacrev``; here the Jacobian
rows are produced tile by tile as extra right-hand-side columns of the same multi-RHS SpMM kernel
that serves ``kron``: column j of a tile is the one-hot seed (train node m, class k), it is pushed
down the layers exactly like a Hessian-sqrt column, and the per-column weight gradient is one GEMM
``gZ_l^T H_{l-1}``.  Memory is O(tile x P), not O(M x C x P), but the work is still M*C backward
columns — like the reference this is a small-graph path."""
from __future__ import annotations

import torch

from . import ops


def diag_ggn_exact(backend, x: torch.Tensor, y: torch.Tensor, tile_bytes: int = 256 << 20):
    if backend.process_group is not None:
        raise NotImplementedError("diag() is a single-device small-graph path")
    model = backend.model
    g = model.graph
    n = g.n
    Ws, bs = backend._layers()
    L = len(Ws)
    idx = x.to(torch.int64).contiguous()
    yy = y.to(torch.int64).contiguous()
    Hs, logits = backend._forward(Ws, bs)
    dev = logits.device
    C = logits.shape[1]
    loss, _ = ops.softmax_ce_sum(logits, idx, yy)
    p = torch.softmax(logits[idx], dim=1)                                   # [M, C]
    dims = [w.shape[0] for w in Ws]
    c_pad = (C + 3) // 4 * 4
    sizes, offs, P = [], [], 0
    for l in range(L):
        offs.append(P)
        P += Ws[l].numel() + (0 if bs[l] is None else dims[l])
    diag = torch.zeros(P, dtype=torch.float32, device=dev)
    dmax = max([c_pad] + dims[:-1])
    M = int(idx.numel())
    nodes_per_tile = max(1, min(M, tile_bytes // (4 * C * max(P, 2 * n * dmax))))
    for m0 in range(0, M, nodes_per_tile):
        nt = min(nodes_per_tile, M - m0)
        T = nt * C
        slab = torch.zeros(n, T * c_pad, dtype=torch.float32, device=dev)
        rows = idx[m0:m0 + nt].repeat_interleave(C)                           # seed (m, k) -> column m*C + k
        cols = torch.arange(T, device=dev) * c_pad + torch.arange(C, device=dev).repeat(nt)
        slab[rows, cols] = 1.0
        J = torch.empty(T, P, dtype=torch.float32, device=dev)
        width, ld = C, c_pad
        for l in range(L - 1, -1, -1):
            gz = ops.spmm(g.ahat_t, slab)                                     # [n, T*ld]
            gz3 = gz.view(n, T, ld)[:, :, :width]
            d_in = Ws[l].shape[1]
            jw = torch.einsum("ntd,ne->tde", gz3, Hs[l])                      # [T, d_l, d_{l-1}]
            J[:, offs[l]:offs[l] + width * d_in] = jw.reshape(T, -1)
            if bs[l] is not None:
                J[:, offs[l] + width * d_in: offs[l] + width * d_in + width] = gz3.sum(0)
            if l > 0:
                d_prev = dims[l - 1]
                nxt = torch.mm(gz.view(n * T, ld)[:, :width], Ws[l])          # [n*T, d_prev]
                ops.relu_mask_mul(nxt, Hs[l], T)
                slab = nxt.view(n, T * d_prev)
                width, ld = d_prev, d_prev
        J3 = J.view(nt, C, P)
        pt = p[m0:m0 + nt]
        lam = torch.diag_embed(pt) - pt.unsqueeze(2) * pt.unsqueeze(1)        # [nt, C, C]
        diag += (torch.bmm(lam, J3) * J3).sum(dim=(0, 1))
    return (backend.factor * loss).to(torch.float32), diag
