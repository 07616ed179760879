"""TEST INFRASTRUCTURE — CPU oracle for SURVEY §8(f) row 3: the gradient of the KFAC-GGN Laplace log
marginal likelihood with respect to the entries of the (binarised) adjacency matrix.

The reference obtains it by keeping the autograd graph alive through ``la.fit()`` and calling
``(-la.log_marginal_likelihood()).backward()``, which fills the dense ``model.adj.grad``
(gnn/marglik_training.py:197-224; STEGCN.forward_adj, gnn/models/models.py:100-115: straight-through
binarisation, ``fill_diagonal_(1)``, ``normalize_adj``; the vendored curvlinops back-propagates with
``create_graph=True``, curvlinops/kfac.py:655-661, and accumulates the factors without detaching,
kfac.py:790, 837).  The straight-through estimator passes the gradient of the binarised matrix to
``adj`` unchanged (gnn/models/utils.py:42-86 with mask=None, sign_grad=False), so ``adj.grad[i, j]``
= d(-marglik)/dA[i, j] at the current 0/1 matrix A; the diagonal is overwritten by ``fill_diagonal_``
and receives no gradient.

This file restates the whole path as ONE differentiable function of a dense A in float64 (torch
autograd on the CPU does the differentiation — the oracle states WHAT is differentiated, the package
derives the adjoint by hand) and is pinned to the reference's own ``adj.grad`` by
oracle/make_golden_adjgrad.py -> tests/golden/adjgrad_*.npz (tests/test_structure.py).
Only tests/ may import it.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

from . import gcn_kfac_oracle as O


def normalized_adj_dense(A: torch.Tensor) -> torch.Tensor:
    """gnn/models/models.py:113-115 + gnn/models/utils.py:106-112: diagonal set to 1, row-sum degrees,
    Â = (A D^-1/2)^T D^-1/2, i.e. Â[i, j] = dis_i A[j, i] dis_j."""
    n = A.shape[0]
    eye = torch.eye(n, dtype=A.dtype)
    A1 = A * (1.0 - eye) + eye                       # fill_diagonal_(1): constant, no gradient
    d = A1.sum(1)
    dis = d.pow(-0.5)
    dis = torch.where(torch.isinf(dis), torch.zeros_like(dis), dis)
    return (A1 * dis[None, :]).T * dis[None, :]


def marglik_of_dense_adj(A: torch.Tensor, x, weights: Sequence, biases: Sequence, idx, y,
                         prior_prec: float = 1.0, mode: str = "reference", batch_size=None) -> torch.Tensor:
    """log marginal likelihood of ``Laplace(model, "classification", "all", "kron")`` after ``fit`` as a
    differentiable function of the dense 0/1 adjacency A (float64).  Follows oracle/gcn_kfac_oracle.py
    (kron_factors, fit_and_marglik, log_marglik) line by line, dense instead of CSR.  ``batch_size``:
    the reference's per-batch accumulation ``loss += loss_b; H += H_b`` (baselaplace.py:778-854): every
    batch runs the full-graph pipeline on its own train nodes, A_b = H^T H / N each."""
    dt = A.dtype
    ahat = normalized_adj_dense(A)
    idx_t = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
    y_t = torch.as_tensor(np.asarray(y), dtype=torch.int64)
    Ws = [torch.as_tensor(np.asarray(w)).to(dt) for w in weights]
    bs = [None if b is None else torch.as_tensor(np.asarray(b)).to(dt) for b in biases]
    L = len(Ws)
    h = torch.as_tensor(np.asarray(x)).to(dt)
    hs, ps = [h], []
    for l in range(L):                                # base_gnn.py:136-161, eval mode
        z = h @ Ws[l].T
        if bs[l] is not None:
            z = z + bs[l]
        p = ahat @ z
        ps.append(p)
        if l < L - 1:
            h = torch.relu(p)
            hs.append(h)
    M = idx_t.numel()
    bsz = M if batch_size is None else int(batch_size)
    n = A.shape[0]
    C = ps[-1].shape[1]
    loss = torch.zeros((), dtype=dt)
    A_fac = [torch.zeros(hh.shape[1], hh.shape[1], dtype=dt) for hh in hs]
    G_fac = [torch.zeros(w.shape[0], w.shape[0], dtype=dt) for w in Ws]
    for s in range(0, M, bsz):
        ib, yb = idx_t[s:s + bsz], y_t[s:s + bsz]
        f = ps[-1][ib]
        loss = loss + torch.nn.functional.cross_entropy(f, yb, reduction="sum")
        V = O.hess_sqrt_rhs(f, mode)                  # [m, C, C], differentiable in f (fork: not detached)
        A_fac = [a + (hh.T @ hh) / M for a, hh in zip(A_fac, hs)]      # (1/M_b) * (M_b/N), N = M
        for c in range(C):
            delta = torch.zeros(n, C, dtype=dt).index_add(0, ib, V[:, c, :])
            for l in range(L - 1, -1, -1):
                gz = ahat.T @ delta
                G_fac[l] = G_fac[l] + gz.T @ gz
                if l > 0:
                    delta = (gz @ Ws[l]) * (ps[l - 1] > 0).to(dt)
    logdet = torch.zeros((), dtype=dt)
    n_params = 0
    theta_sq = torch.zeros((), dtype=dt)
    for l in range(L):
        lg = torch.linalg.eigvalsh(G_fac[l], UPLO="U").clamp(min=0.0)
        la = torch.linalg.eigvalsh(A_fac[l], UPLO="U").clamp(min=0.0)
        logdet = logdet + torch.log(torch.outer(lg, la) + prior_prec).sum()
        n_params += Ws[l].numel()
        theta_sq = theta_sq + (Ws[l] ** 2).sum()
        if bs[l] is not None:
            logdet = logdet + torch.log(lg + prior_prec).sum()
            n_params += bs[l].numel()
            theta_sq = theta_sq + (bs[l] ** 2).sum()
    return -loss - 0.5 * (logdet - n_params * math.log(prior_prec) + prior_prec * theta_sq)


def marglik_adj_grad(adj01: np.ndarray, x, weights, biases, idx, y, prior_prec: float = 1.0,
                     mode: str = "reference", batch_size=None):
    """(marglik, d marglik / dA as a dense [n, n] float64 array; zero diagonal)."""
    A = torch.tensor(np.asarray(adj01), dtype=torch.float64, requires_grad=True)
    ml = marglik_of_dense_adj(A, x, weights, biases, idx, y, prior_prec, mode, batch_size)
    (g,) = torch.autograd.grad(ml, A)
    return float(ml), g.numpy()


def dense_adj01(edge_index: np.ndarray, n: int, symmetric: bool = False) -> np.ndarray:
    """Binary adjacency of the edge list without the diagonal convention applied
    (gnn/utils.py:325-330 + clamp, marglik_training.py:403-405; optional A + A^T clamp)."""
    a = np.zeros((n, n), dtype=np.float64)
    a[edge_index[0], edge_index[1]] = 1.0
    if symmetric:
        a = np.minimum(a + a.T, 1.0)
    return a
