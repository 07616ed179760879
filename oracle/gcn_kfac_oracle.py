"""TEST INFRASTRUCTURE — CPU restatement (the "oracle") of the reference's
GCN + KFAC-GGN Laplace hot path.  NOT product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  The product package never does.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the reference's own
classes (``/root/reference`` through ``oracle/ref_loader.py``) on small graphs
and on the Cora-/Pubmed-shaped configs and stores inputs + outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here
against those files (factors <= 1e-5 rel, marglik <= 1e-5 rel in fp32; tighter
in fp64).  The ``hess_sqrt="ggn"`` mode is pinned on upstream curvlinops'
arithmetic: ``oracle/make_golden_ggn.py`` runs the same reference classes with
upstream's ``out.detach()`` restored at run time on the Hessian-sqrt input (the
one expression the fork changed, curvlinops/kfac.py:631-642) and stores
``tests/golden/ggn_*.npz``.  The alternate backends the north star names
(asdl, backpack) are not installed: against those, parity unpinned.

Integer / index work is numpy (bit-exact contract).  Floating-point work uses
torch CPU tensors (fp32 like the reference, or fp64) because the reference's
arithmetic *is* torch CPU; no autograd is used anywhere.

Reference lines restated (all relative to /root/reference):
  gnn/utils.py:325-330, gnn/marglik_training.py:403-405   edge list -> 0/1 adjacency
  gnn/models/models.py:23, gnn/models/base_gnn.py:68-73   self loops, symmetrise
  gnn/models/utils.py:106-112                             normalize_adj
  gnn/models/layers.py:45-46, gnn/models/base_gnn.py:136-161   forward
  curvlinops/kfac_utils.py:122-126, curvlinops/kfac.py:628-661 Hessian sqrt, C backward passes
  curvlinops/kfac.py:777-817, :819-875                    G and A accumulation
  laplace/curvature/curvlinops.py:46-108                  packing, M/N rescale, loss
  laplace/utils/matrix.py:118-145, :371-394, laplace/utils/utils.py:193-226   decompose, logdet
  laplace/baselaplace.py:210-232, :856-973                marglik algebra
  laplace/curvature/curvature.py:365-372, :412-432        exact diag GGN
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# (a1) edge list -> adjacency pattern with self loops          INTEGER, bit-exact contract
# --------------------------------------------------------------------------------------


def coo_to_adj_csr(edge_index: np.ndarray, num_nodes: int, symmetric: bool = False
                   ) -> Tuple[np.ndarray, np.ndarray]:
    """CSR pattern (rowptr int64[N+1], col int32[nnz]) of the binary matrix A.

    A[s, d] = 1 for every edge (s -> d) in ``edge_index`` (duplicates collapse: scipy
    sums them and the caller clamps to 1, gnn/utils.py:325-330 +
    marglik_training.py:403-405); optional ``A = clamp(A + A^T)``
    (base_gnn.py:68-73); then the diagonal is *set* to 1 (models.py:23).
    Columns are sorted ascending inside each row.
    """
    ei = np.asarray(edge_index, dtype=np.int64)
    assert ei.ndim == 2 and ei.shape[0] == 2
    n = int(num_nodes)
    src, dst = ei[0], ei[1]
    if src.size and (src.min() < 0 or dst.min() < 0 or src.max() >= n or dst.max() >= n):
        raise ValueError("edge index out of range")
    if symmetric:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    loops = np.arange(n, dtype=np.int64)
    keys = np.concatenate([src * n + dst, loops * n + loops])
    keys = np.unique(keys)  # sorted + deduplicated
    rows = keys // n
    cols = (keys - rows * n).astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    return rowptr, cols


def csr_transpose_pattern(rowptr: np.ndarray, col: np.ndarray, n: int
                          ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Pattern transpose; also returns perm with col_t[k] coming from entry perm[k]."""
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    order = np.lexsort((rows, col.astype(np.int64)))  # by new row (=col), then new col (=row)
    t_rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(col, minlength=n), out=t_rowptr[1:])
    return t_rowptr, rows[order].astype(np.int32), order


@dataclass
class NormAdj:
    """Â and Â^T in CSR.  Â[i, j] = dis_i * A[j, i] * dis_j  (utils.py:106-112)."""
    n: int
    deg: np.ndarray      # int64 [N]   row sums of A (out-degree incl. self loop)
    dis: np.ndarray      # fp32  [N]   deg^-1/2 (0 where deg == 0)
    rowptr: np.ndarray   # Â    : row i holds the in-neighbours j of i
    col: np.ndarray
    val: np.ndarray
    t_rowptr: np.ndarray  # Â^T : row i holds the out-neighbours (= pattern of A)
    t_col: np.ndarray
    t_val: np.ndarray

    @property
    def nnz(self) -> int:
        return int(self.col.shape[0])


def inv_sqrt_degree(deg: np.ndarray) -> np.ndarray:
    """fp32 1/sqrt(deg) with IEEE-rounded sqrt and divide (inf -> 0).

    The reference uses ``rowsum.pow(-0.5)`` (utils.py:107-109); ATen's result is
    within 2 ulp of this, while IEEE sqrt+div is reproducible to
    the bit on the GPU (``__fsqrt_rn`` / ``__fdiv_rn``) — that is the contract.
    """
    d = deg.astype(np.float32)
    with np.errstate(divide="ignore"):
        out = np.float32(1.0) / np.sqrt(d)
    out[deg == 0] = np.float32(0.0)
    return out.astype(np.float32)


def normalize_adj_csr(a_rowptr: np.ndarray, a_col: np.ndarray, n: int) -> NormAdj:
    deg = np.diff(a_rowptr).astype(np.int64)
    dis = inv_sqrt_degree(deg)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    # Â^T shares A's pattern:  Â^T[i, j] = Â[j, i] = dis_j * A[i, j] * dis_i
    t_val = (dis[rows] * dis[a_col]).astype(np.float32)
    rowptr, col, perm = csr_transpose_pattern(a_rowptr, a_col, n)
    val = t_val[perm]
    return NormAdj(n, deg, dis, rowptr, col, val, a_rowptr.copy(), a_col.copy(), t_val)


def build_graph(edge_index: np.ndarray, num_nodes: int, symmetric: bool = False) -> NormAdj:
    rp, col = coo_to_adj_csr(edge_index, num_nodes, symmetric)
    return normalize_adj_csr(rp, col, num_nodes)


def row_partition(rowptr: np.ndarray, nparts: int) -> np.ndarray:
    """nnz-balanced contiguous row blocks: bounds[r] = first row i with
    rowptr[i] >= floor(r * nnz / nparts); bounds[0] = 0, bounds[nparts] = N."""
    n = rowptr.shape[0] - 1
    nnz = int(rowptr[-1])
    bounds = np.zeros(nparts + 1, dtype=np.int64)
    for r in range(1, nparts):
        target = (r * nnz) // nparts
        bounds[r] = np.searchsorted(rowptr, target, side="left")
    bounds[nparts] = n
    return np.minimum(bounds, n)


def halo_columns(rowptr: np.ndarray, col: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Sorted unique column ids referenced by rows [lo, hi) that lie outside [lo, hi)."""
    c = col[rowptr[lo]:rowptr[hi]].astype(np.int64)
    c = c[(c < lo) | (c >= hi)]
    return np.unique(c)


# --------------------------------------------------------------------------------------
# float pipeline (torch CPU tensors; dtype fp32 like the reference, or fp64)
# --------------------------------------------------------------------------------------


def _csr_tensor(rowptr, col, val, n, dtype):
    return torch.sparse_csr_tensor(
        torch.from_numpy(np.ascontiguousarray(rowptr)).to(torch.int64),
        torch.from_numpy(np.ascontiguousarray(col)).to(torch.int64),
        torch.from_numpy(np.ascontiguousarray(val)).to(dtype), size=(n, n))


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.detach().to("cpu", dtype)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def spmm(g: NormAdj, x, transpose: bool = False, dtype=torch.float32) -> torch.Tensor:
    """Y = Â X  (layers.py:46) or Â^T X (its autograd backward)."""
    x = _t(x, dtype)
    if transpose:
        m = _csr_tensor(g.t_rowptr, g.t_col, g.t_val, g.n, dtype)
    else:
        m = _csr_tensor(g.rowptr, g.col, g.val, g.n, dtype)
    return m @ x


def forward(g: NormAdj, x, weights: Sequence, biases: Sequence, dtype=torch.float32):
    """Eval-mode forward (base_gnn.py:136-161): P_l = Â (H_{l-1} W_l^T + b_l),
    H_l = relu(P_l) for l < L.  Returns (Hs = [H_0..H_{L-1}], Ps = [P_1..P_L])."""
    m = _csr_tensor(g.rowptr, g.col, g.val, g.n, dtype)
    h = _t(x, dtype)
    hs, ps = [h], []
    L = len(weights)
    for l in range(L):
        z = h @ _t(weights[l], dtype).T
        if biases[l] is not None:
            z = z + _t(biases[l], dtype)
        p = m @ z
        ps.append(p)
        if l < L - 1:
            h = torch.relu(p)
            hs.append(h)
    return hs, ps


def hess_sqrt_rhs(f: torch.Tensor, mode: str = "reference") -> torch.Tensor:
    """V[n, c, :] = vector injected at the logits of sample n for Hessian-sqrt column c.

    mode="ggn":        v = sqrt(p_c) (e_c - p)            (kfac_utils.py:122-126, detached)
    mode="reference":  the fork back-propagates (f ⊙ S_c(f)).sum() with S *not*
                       detached (kfac.py:631-642, :653-661), which adds Σ_i f_i ∂S_ic/∂f:
        v = sqrt(p_c) [ (e_c - p)(1 + ½(f_c - f̄)) - p ⊙ (f - f̄) ],  f̄ = p·f
    """
    p = torch.softmax(f, dim=1)
    n, C = f.shape
    sp = p.sqrt()
    eye = torch.eye(C, dtype=f.dtype)
    ec_minus_p = eye.unsqueeze(0) - p.unsqueeze(1)          # [n, c, k] = δ_ck - p_k
    if mode == "ggn":
        return sp.unsqueeze(2) * ec_minus_p
    if mode != "reference":
        raise ValueError(f"unknown hess_sqrt mode {mode!r}")
    fbar = (p * f).sum(1, keepdim=True)                       # [n,1]
    fc = f - fbar                                             # [n,k] = f_k - f̄
    a = 1.0 + 0.5 * fc                                        # indexed by c
    v = ec_minus_p * a.unsqueeze(2) - (p * fc).unsqueeze(1)
    return sp.unsqueeze(2) * v


def cross_entropy_sum(f: torch.Tensor, y) -> torch.Tensor:
    y = torch.as_tensor(np.asarray(y), dtype=torch.int64)
    return torch.nn.functional.cross_entropy(f, y, reduction="sum")


def kron_factors(g: NormAdj, x, weights, biases, idx, y, n_data: int,
                 mode: str = "reference", dtype=torch.float32):
    """One call of ``backend.kron(idx, y, N=n_data)`` (curvlinops.py:77-108).

    Returns (loss, kfacs) with kfacs = [[G_1, A_1], [G_1], [G_2, A_2], [G_2], ...]
    (output-side factor first, bias block = [G]; curvlinops.py:55-75).
      A_l = H_{l-1}^T H_{l-1} / M * (M / N)   over ALL graph nodes (kfac.py:870, curvlinops.py:46-53)
      G_l = Σ_c gZ_{l,c}^T gZ_{l,c},  gZ_{l,c} = Â^T δ_{l,c}          (kfac.py:777-817)
      δ_{L,c} = V[:, c, :] scattered to rows idx;  δ_{l-1,c} = (gZ_{l,c} W_l) ⊙ 1[P_{l-1} > 0]
    """
    idx_t = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
    M = int(idx_t.numel())
    hs, ps = forward(g, x, weights, biases, dtype)
    L = len(weights)
    f = ps[-1][idx_t]
    loss = cross_entropy_sum(f, y)
    V = hess_sqrt_rhs(f, mode)                                # [M, C, C]
    C = f.shape[1]
    mt = _csr_tensor(g.t_rowptr, g.t_col, g.t_val, g.n, dtype)
    Ws = [_t(w, dtype) for w in weights]
    A = [(h.T @ h) / M * (M / n_data) for h in hs]
    G = [torch.zeros(w.shape[0], w.shape[0], dtype=dtype) for w in Ws]
    for c in range(C):
        delta = torch.zeros(g.n, C, dtype=dtype)
        # duplicates in idx accumulate, as autograd's index backward does
        delta.index_add_(0, idx_t, V[:, c, :])
        for l in range(L - 1, -1, -1):
            gz = mt @ delta
            G[l] += gz.T @ gz
            if l > 0:
                delta = (gz @ Ws[l]) * (ps[l - 1] > 0).to(dtype)
    kfacs = []
    for l in range(L):
        if biases[l] is not None:
            kfacs.append([G[l], A[l]])
            kfacs.append([G[l].clone()])
        else:
            kfacs.append([G[l], A[l]])
    return loss, kfacs


def symeig(m: torch.Tensor):
    """utils.py:193-226: eigh(UPLO='U'); if the solver does not converge, the reference's jitter fallback
    (decompose M + I, take 1 off the eigenvalues, :209-216); eigenvalues clamped >= 0, NaN -> 0."""
    try:
        lam, q = torch.linalg.eigh(m, UPLO="U")
    except RuntimeError:
        lam, q = torch.linalg.eigh(m + torch.eye(m.shape[0], dtype=m.dtype), UPLO="U")
        lam = lam - 1.0
    return torch.nan_to_num(lam.clamp(min=0.0)), torch.nan_to_num(q)


def kron_logdet(kfacs, delta) -> torch.Tensor:
    """matrix.py:118-145 + :371-394 with a prior precision δ per block (scalar or list)."""
    out = 0.0
    for b, F in enumerate(kfacs):
        d = delta if not isinstance(delta, (list, tuple)) else delta[b]
        lams = [symeig(Hi)[0] for Hi in F]
        if len(lams) == 1:
            out = out + torch.log(lams[0] + d).sum()
        else:
            out = out + torch.log(torch.outer(lams[0], lams[1]) + d).sum()
    return out


def log_marglik(loss, kfacs, params: Sequence, prior_prec: float = 1.0,
                dtype=torch.float32) -> torch.Tensor:
    """baselaplace.py:938-973 for classification, temperature 1, scalar prior:
       -loss - ½ [ logdet(H + δI) - P log δ + δ ‖θ‖² ]."""
    theta = torch.cat([_t(p, dtype).reshape(-1) for p in params])
    P = theta.numel()
    logdet_post = kron_logdet([[Hi.to(dtype) for Hi in F] for F in kfacs], prior_prec)
    logdet_prior = P * math.log(prior_prec)
    scatter = prior_prec * (theta @ theta)
    return -loss - 0.5 * (logdet_post - logdet_prior + scatter)


def params_in_order(weights, biases) -> List:
    """named_parameters() order minus adj: convs.0.lin.weight, convs.0.lin.bias, ..."""
    out = []
    for w, b in zip(weights, biases):
        out.append(w)
        if b is not None:
            out.append(b)
    return out


def fit_and_marglik(g: NormAdj, x, weights, biases, idx, y, prior_prec=1.0,
                    mode="reference", dtype=torch.float32, batch_size=None):
    """KronLaplace.fit + log_marginal_likelihood (baselaplace.py:778-854, 1580-1610).
    ``batch_size=None`` = a single full batch; otherwise the reference's per-batch
    accumulation ``loss += loss_b; H += H_b`` (SURVEY §8a multi-batch semantics)."""
    idx = np.asarray(idx)
    y = np.asarray(y)
    n = idx.shape[0]
    bs = n if batch_size is None else batch_size
    loss, kfacs = None, None
    for s in range(0, n, bs):
        lb, kb = kron_factors(g, x, weights, biases, idx[s:s + bs], y[s:s + bs], n, mode, dtype)
        if kfacs is None:
            loss, kfacs = lb, kb
        else:
            loss = loss + lb
            kfacs = [[a + b for a, b in zip(Fa, Fb)] for Fa, Fb in zip(kfacs, kb)]
    ml = log_marglik(loss, kfacs, params_in_order(weights, biases), prior_prec, dtype)
    return loss, kfacs, ml


# --------------------------------------------------------------------------------------
# predictive path on the Kron posterior
# --------------------------------------------------------------------------------------


def kron_bmm(kfacs, W: torch.Tensor, exponent: float, delta: float = 1.0) -> torch.Tensor:
    """(H + δI)^exponent @ W^T for the block-diagonal Kron H, row-wise on W [S, P]
    (KronDecomposed._bmm, matrix.py:396-446; weight blocks are flattened [d_out, d_in] row-major
    and their factors are [G (d_out), A (d_in)])."""
    out, cur = [], 0
    for F in kfacs:
        if len(F) == 1:
            lam, Q = symeig(F[0].to(W.dtype))
            p = lam.numel()
            Wp = W[:, cur:cur + p].T
            out.append((Q @ (torch.pow(lam + delta, exponent).reshape(-1, 1) * (Q.T @ Wp))).T)
        else:
            (l1, Q1), (l2, Q2) = symeig(F[0].to(W.dtype)), symeig(F[1].to(W.dtype))
            p = l1.numel() * l2.numel()
            Wp = W[:, cur:cur + p].reshape(-1, l1.numel(), l2.numel())
            Wp = (Q1.T @ Wp @ Q2) * torch.pow(torch.outer(l1, l2) + delta, exponent).unsqueeze(0)
            out.append((Q1 @ Wp @ Q2.T).reshape(-1, p))
        cur += p
    return torch.cat(out, dim=1)


def posterior_samples(mean, kfacs, eps, prior_prec: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    """KronLaplace.sample for given standard-normal draws eps [S, P] (baselaplace.py:1646-1655)."""
    return _t(mean, dtype).reshape(1, -1) + kron_bmm(kfacs, _t(eps, dtype), -0.5, prior_prec)


def unpack_params(theta, shapes):
    """Parameter vector -> ([W_l], [b_l]) in named_parameters() order (weight then bias per layer)."""
    Ws, bs, cur = [], [], 0
    for (d_out, d_in) in shapes:
        Ws.append(theta[cur:cur + d_out * d_in].reshape(d_out, d_in)); cur += d_out * d_in
        bs.append(theta[cur:cur + d_out]); cur += d_out
    return Ws, bs


def mc_predictive(g: NormAdj, x, samples, shapes, idx, dtype=torch.float32) -> torch.Tensor:
    """_nn_predictive_classification (baselaplace.py:1183-1199): mean over weight samples of
    softmax(model_sample(idx)); one eval-mode full-graph forward per sample."""
    idx_t = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
    samples = _t(samples, dtype)
    py = 0.0
    for s in range(samples.shape[0]):
        Ws, bs = unpack_params(samples[s], shapes)
        _, ps = forward(g, x, Ws, bs, dtype)
        py = py + torch.softmax(ps[-1][idx_t], dim=-1) / samples.shape[0]
    return py


# --------------------------------------------------------------------------------------
# exact diag GGN (curvature.py:412-432) — small graphs only: builds the dense Jacobian
# --------------------------------------------------------------------------------------


def diag_ggn(g: NormAdj, x, weights, biases, idx, y, dtype=torch.float64):
    """diag(Σ_n J_n^T Λ_n J_n) with the TRUE Λ = diag(p) - p p^T (curvature.py:365-372).
    Dense Jacobian by brute force (one backward sweep per (n, k)); only for tiny cases."""
    idx_np = np.asarray(idx)
    hs, ps = forward(g, x, weights, biases, dtype)
    L = len(weights)
    f = ps[-1][torch.as_tensor(idx_np, dtype=torch.int64)]
    loss = cross_entropy_sum(f, y)
    p = torch.softmax(f, 1)
    C = f.shape[1]
    mt = _csr_tensor(g.t_rowptr, g.t_col, g.t_val, g.n, dtype)
    Ws = [_t(w, dtype) for w in weights]
    sizes = []
    for l in range(L):
        sizes.append(Ws[l].numel())
        if biases[l] is not None:
            sizes.append(Ws[l].shape[0])
    P = sum(sizes)
    diag = torch.zeros(P, dtype=dtype)
    for n_i, node in enumerate(idx_np):
        J = torch.zeros(C, P, dtype=dtype)
        for k in range(C):
            delta = torch.zeros(g.n, C, dtype=dtype)
            delta[node, k] = 1.0
            grads = [None] * L
            for l in range(L - 1, -1, -1):
                gz = mt @ delta
                grads[l] = (gz.T @ hs[l], gz.sum(0))
                if l > 0:
                    delta = (gz @ Ws[l]) * (ps[l - 1] > 0).to(dtype)
            parts = []
            for l in range(L):
                parts.append(grads[l][0].reshape(-1))
                if biases[l] is not None:
                    parts.append(grads[l][1])
            J[k] = torch.cat(parts)
        lam = torch.diag(p[n_i]) - torch.outer(p[n_i], p[n_i])
        diag += torch.einsum("cp,ck,kp->p", J, lam, J)
    return loss, diag


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench (numpy PCG64 → identical on every box)
# --------------------------------------------------------------------------------------


def synthetic_edges(n: int, n_undirected: int, seed: int = 0, directed: bool = False,
                    rmat: bool = False) -> np.ndarray:
    """Uniform random pairs (or R-MAT a,b,c,d=.57,.19,.19,.05), mirrored unless directed."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if rmat:
        scale = int(math.ceil(math.log2(max(n, 2))))
        src = np.zeros(n_undirected, dtype=np.int64)
        dst = np.zeros(n_undirected, dtype=np.int64)
        for _ in range(scale):
            r = rng.random(n_undirected)
            src = (src << 1) | ((r >= 0.76).astype(np.int64))
            dst = (dst << 1) | (((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).astype(np.int64)
        src %= n
        dst %= n
    else:
        src = rng.integers(0, n, n_undirected, dtype=np.int64)
        dst = rng.integers(0, n, n_undirected, dtype=np.int64)
    if directed:
        return np.stack([src, dst])
    return np.stack([np.concatenate([src, dst]), np.concatenate([dst, src])])
