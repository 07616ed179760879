"""Sampling predictive on the Kron posterior: the reference runs one full-graph
forward per weight sample (``_nn_predictive_classification``, laplace/baselaplace.py:1183-1199, called
by ``mc_eval``, gnn/marglik_training.py:341-353).  Here a tile of samples goes through the graph
together: their per-sample linear layers are written side by side into one slab [N, S*d] and ONE
multi-RHS SpMM aggregates all of them (the same kernels as the KFAC backward)."""
from __future__ import annotations

import torch

from . import ops
from .gcn import SparseGCN


def _unpack(samples: torch.Tensor, model: SparseGCN):
    Ws, bs, cur = [], [], 0
    S = samples.shape[0]
    for conv in model.convs:
        d_out, d_in = conv.lin.weight.shape
        Ws.append(samples[:, cur:cur + d_out * d_in].reshape(S, d_out, d_in))
        cur += d_out * d_in
        if conv.lin.bias is not None:
            bs.append(samples[:, cur:cur + d_out])
            cur += d_out
        else:
            bs.append(None)
    if cur != samples.shape[1]:
        raise ValueError(f"samples have {samples.shape[1]} parameters, the model has {cur}")
    return Ws, bs


def mc_predictive(model: SparseGCN, samples: torch.Tensor, idx: torch.Tensor,
                  tile_bytes: int | None = None) -> torch.Tensor:
    """mean_s softmax(f_{theta_s}(idx)) for weight samples [S, P] (parameter order of
    ``named_parameters()``: weight then bias per layer); eval mode."""
    if not isinstance(model, SparseGCN):
        raise TypeError("mc_predictive needs a laplace_gnn_b200.SparseGCN model")
    g = model.graph
    X = model.X.float().contiguous()
    samples = samples.to(torch.float32)
    n, dev = g.n, X.device
    S = samples.shape[0]
    Ws, bs = _unpack(samples, model)
    L = len(Ws)
    dims = [w.shape[1] for w in Ws]                      # d_out per layer
    C = dims[-1]
    lds = [(d + 3) // 4 * 4 for d in dims]               # 16-byte rows for the 128-bit SpMM paths
    if tile_bytes is None:
        if dev.type == "cuda":
            free, total = torch.cuda.mem_get_info(dev)
            free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
            tile_bytes = min(int(0.5 * free), int(0.4 * total))
        else:
            tile_bytes = 1 << 30
    tile = max(1, min(S, int(tile_bytes // (2 * n * max(lds) * 4))))
    idx = idx.to(torch.int64)
    py = torch.zeros(idx.numel(), C, dtype=torch.float32, device=dev)
    for s0 in range(0, S, tile):
        sg = min(tile, S - s0)
        h = None
        for l in range(L):
            d_out, ld = dims[l], lds[l]
            alloc = torch.zeros if ld != d_out else torch.empty      # pad columns must be zero
            z = alloc(n, sg * ld, dtype=torch.float32, device=dev)
            if l == 0 and ld == d_out:                   # one GEMM for all samples: X @ [W_1; ...; W_sg]^T
                wcat = Ws[0][s0:s0 + sg].reshape(sg * d_out, -1)
                if bs[0] is None:
                    torch.mm(X, wcat.t(), out=z)
                else:
                    torch.addmm(bs[0][s0:s0 + sg].reshape(-1), X, wcat.t(), out=z)
            else:
                d_in = Ws[l].shape[2]
                ld_in = lds[l - 1] if l > 0 else d_in
                for j in range(sg):
                    src = X if l == 0 else h[:, j * ld_in: j * ld_in + d_in]
                    zj = torch.mm(src, Ws[l][s0 + j].t())
                    if bs[l] is not None:
                        zj += bs[l][s0 + j]
                    z[:, j * ld: j * ld + d_out] = zj
            h = ops.spmm(g.ahat, z, relu=(l < L - 1))
        logits = h[idx].view(idx.numel(), sg, lds[-1])[:, :, :C]
        py += torch.softmax(logits, dim=-1).sum(dim=1)
    return py / S
