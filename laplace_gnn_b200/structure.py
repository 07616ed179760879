"""Gradient of the KFAC-GGN Laplace log marginal likelihood with respect to the adjacency entries —
the quantity the reference's graph-structure learning descends on.

The reference keeps the autograd graph alive through ``la.fit()`` (curvlinops/kfac.py:655-661
back-propagates with ``create_graph=True``; the factors are accumulated without detaching, kfac.py:790,
837) and calls ``(-la.log_marginal_likelihood()).backward()`` to fill the DENSE ``model.adj.grad``
(gnn/marglik_training.py:197-224; STEGCN.forward_adj, gnn/models/models.py:100-115: straight-through
binarisation, ``fill_diagonal_(1)``, ``normalize_adj``).  That needs C retained N x N autograd graphs.

Here the adjoint is derived by hand and evaluated on a SPARSE set of entries — the existing edges of
A and any list of candidate (non-)edges — with the kernels of the hot path plus one SDDMM:

  marglik = -loss - 1/2 [ sum_l logdet(G_l (x) A_l + delta) (+ bias blocks) - P log delta + delta |theta|^2 ]

  factor adjoints   Gbar_l = -1/2 U diag( sum_j a_j / (g_i a_j + delta) [+ 1 / (g_i + delta)] ) U^T,
                    Abar_l likewise, from the eigendecompositions the marglik needs anyway
  G part, per group of Hessian-sqrt columns (the KFAC backward is recomputed, its slabs kept):
                    gZbar_l   = 2 gZ_l Gbar_l + (dbar_{l-1} * relu') W_l^T
                    dAhat    += SDDMM(gZbar_l, delta_l)            (Â enters as gZ_l = Â^T delta_l)
                    dbar_l    = Â gZbar_l                           (SpMM with Â instead of Â^T)
                    fbar     += J_v(f)^T dbar_L[idx]                (the Hessian square root is NOT detached
                                                                    in the fork; closed form)
  forward part      Pbar_L = scatter(fbar - (softmax(f) - onehot(y)));   dAhat += SDDMM(Z_l, Pbar_l);
                    Zbar_l = Â^T Pbar_l;  Pbar_{l-1} = (Zbar_l W_l + 2 H_{l-1} Abar_l / N) * relu'
  normalisation     Â[i, j] = dis_i A[j, i] dis_j, dis = rowsum(A)^-1/2:
                    dA[m, k] = dAhat[k, m] Â[k, m] + r_m,  r_m = -(rowsum_m(S) + colsum_m(S)) / (2 d_m),
                    S = dAhat * Â; the diagonal is a constant (fill_diagonal_) and gets 0.

Scope: one full batch or the reference's per-batch accumulation, scalar or per-block prior precision, the
weights are constants.
Graphs built with ``symmetric=True`` (the model-side (A + A^T) symmetrisation) are not covered.
Cost: about four fits (every SpMM of the KFAC backward once more with Â, plus an SDDMM that reads
both slabs per edge).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import ops
from .curvature import B200GGN
from .gcn import SparseGCN


@dataclass
class EdgeGradient:
    """d marglik / dA on the pattern of A (aligned with ``graph.ahat_t``: entry e is A[rows[e], cols[e]])
    and on the candidate entries."""
    marglik: torch.Tensor            # 0-d
    loss: torch.Tensor               # 0-d, sum CE
    rows: torch.Tensor               # int32 [nnz(A)]
    cols: torch.Tensor               # int32 [nnz(A)]
    grad_edges: torch.Tensor         # fp32 [nnz(A)]; 0 on the diagonal
    grad_candidates: torch.Tensor | None
    kfacs: list


def hess_sqrt_columns(f: torch.Tensor, mode: str) -> torch.Tensor:
    """V[m, c, :] = the vector injected at the logits for Hessian-sqrt column c, as a differentiable
    function of f (same closed form as lgnn_hess_rhs_f32; curvlinops/kfac_utils.py:122-126 +
    kfac.py:631-661)."""
    p = torch.softmax(f, dim=1)
    C = f.shape[1]
    sp = p.sqrt()
    e = torch.eye(C, dtype=f.dtype, device=f.device).unsqueeze(0) - p.unsqueeze(1)      # [m, c, k]
    if mode == "ggn":
        return sp.unsqueeze(2) * e
    fc = f - (p * f).sum(1, keepdim=True)
    return sp.unsqueeze(2) * (e * (1.0 + 0.5 * fc).unsqueeze(2) - (p * fc).unsqueeze(1))


def _factor_adjoints(kfacs, has_bias, deltas):
    """(logdet of the posterior precision, [Gbar_l], [Abar_l]) in float64 -> float32."""
    logdet = 0.0
    Gbar, Abar = [], []
    b = 0
    for l, bias in enumerate(has_bias):
        G, A = kfacs[b][0].double(), kfacs[b][1].double()
        dw = deltas[b]
        lg_raw, UG = torch.linalg.eigh(G, UPLO="U")
        la_raw, UA = torch.linalg.eigh(A, UPLO="U")
        lg, la = lg_raw.clamp(min=0.0), la_raw.clamp(min=0.0)
        den = torch.outer(lg, la) + dw
        logdet = logdet + torch.log(den).sum()
        gG = (la[None, :] / den).sum(1)
        gA = (lg[:, None] / den).sum(0)
        b += 1
        if bias:
            db = deltas[b]
            logdet = logdet + torch.log(lg + db).sum()
            gG = gG + 1.0 / (lg + db)
            b += 1
        gG = torch.where(lg_raw > 0, gG, torch.zeros_like(gG))          # clamped eigenvalues are constants
        gA = torch.where(la_raw > 0, gA, torch.zeros_like(gA))
        Gbar.append((-0.5 * (UG * gG[None, :]) @ UG.T).float())
        Abar.append((-0.5 * (UA * gA[None, :]) @ UA.T).float())
    return logdet, Gbar, Abar


def marglik_edge_grad(model: SparseGCN, idx: torch.Tensor, y: torch.Tensor, prior_precision=1.0,
                      hess_sqrt: str = "reference", candidates: torch.Tensor | None = None,
                      group: int | None = None, batch_size: int | None = None) -> EdgeGradient:
    """log marginal likelihood of ``Laplace(model, "classification", "all", "kron")`` after ``fit`` on
    (idx, y), and its gradient with respect to A[m, k] for every edge of the graph and every
    ``candidates[:, e] = (m, k)`` (int64 [2, K]; entries that are not edges).  ``batch_size=None`` is one
    full batch; otherwise the reference's per-batch accumulation (``loss += loss_b; H += H_b``,
    baselaplace.py:778-854, with the driver's ``batch_size=10000``, gnn/marglik_training.py:125-127): every
    batch back-propagates its own train nodes through the whole graph, the factors add up, and so do
    the batches' adjoints (the factor adjoints come from the summed factors)."""
    if not isinstance(model, SparseGCN):
        raise TypeError("marglik_edge_grad needs a laplace_gnn_b200.SparseGCN model")
    g = model.graph
    if g.meta.get("symmetrised", False):
        raise NotImplementedError("graphs symmetrised on the model side (symmetric=True) are not covered")
    be = B200GGN(model, "classification", hess_sqrt=hess_sqrt, unit_slabs=False)
    Ws, bs = be._layers()
    L, n = len(Ws), g.n
    idx = idx.to(torch.int64).contiguous()
    y = y.to(torch.int64).contiguous()
    M = int(idx.numel())
    dev = Ws[0].device
    bsz = M if batch_size is None else max(1, int(batch_size))
    batches = [(s0, min(M, s0 + bsz)) for s0 in range(0, M, bsz)]
    loss, kron = None, None
    for s0, s1 in batches:                                   # ParametricLaplace.fit
        lb, kb = be.kron(idx[s0:s1], y[s0:s1], N=M)
        loss, kron = (lb, kb) if kron is None else (loss + lb, kron + kb)
    Hs, logits = be._forward(Ws, bs)
    C = logits.shape[1]
    c_pad = (C + 3) // 4 * 4
    dims = [w.shape[0] for w in Ws]
    has_bias = [b is not None for b in bs]
    n_blocks = sum(2 if hb else 1 for hb in has_bias)
    pp = torch.as_tensor(prior_precision, dtype=torch.float64, device=dev).reshape(-1)
    if pp.numel() == 1:
        pp = pp.expand(n_blocks)
    if pp.numel() != n_blocks:
        raise ValueError("prior_precision must be a scalar or one value per parameter tensor")

    # ---- marglik value and the adjoints of the factors
    logdet, Gbar, Abar = _factor_adjoints(kron.kfacs, has_bias, pp)
    params = []
    for w, b in zip(Ws, bs):
        params.append(w)
        if b is not None:
            params.append(b)
    quad = sum(float(d) * float((p.double() ** 2).sum()) for d, p in zip(pp, params))
    logdet_prior = sum(p.numel() * math.log(float(d)) for d, p in zip(pp, params))
    marglik = -loss.double() - 0.5 * (logdet - logdet_prior + quad)

    # ---- entries: the pattern of A = CSR of Â^T (row m, column k <-> A[m, k] <-> Â[k, m])
    at = g.ahat_t
    rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32),
                                   (at.rowptr[1:] - at.rowptr[:-1]))
    cols = at.col
    gt = torch.zeros(at.nnz, dtype=torch.float32, device=dev)              # dL/dÂ[k, m] per entry (m, k)
    cand_rows = cand_cols = gc_t = None
    if candidates is not None and candidates.numel() > 0:
        cand_rows = candidates[0].to(torch.int32).contiguous()
        cand_cols = candidates[1].to(torch.int32).contiguous()
        gc_t = torch.zeros(cand_rows.numel(), dtype=torch.float32, device=dev)

    def edge_dots(U, V, d):
        """dAhat[k, m] += U[m, :d] . V[k, :d] on every requested entry (m, k)."""
        ops.sddmm(rows, cols, U, V, d, out=gt, accumulate=True)
        if gc_t is not None:
            ops.sddmm(cand_rows, cand_cols, U, V, d, out=gc_t, accumulate=True)

    # ---- G part: per group of Hessian-sqrt columns
    if group is None:
        per_col = 2 * n * (c_pad + 2 * sum(dims[:-1])) * 4 + 1
        if dev.type == "cuda":
            free, _ = torch.cuda.mem_get_info(dev)
            free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        else:
            free = 1 << 30
        group = max(1, min(C, int(0.4 * free // per_col)))
    f_train = logits[idx][:, :C].contiguous()
    fbar = torch.zeros(M, C, dtype=torch.float32, device=dev)
    for (s0, s1), c0 in ((b, c) for b in batches for c in range(0, C, group)):
        gc = min(group, C - c0)
        idx_b = idx[s0:s1]
        deltas, gzs = [None] * L, [None] * L
        delta = torch.zeros(n, gc * c_pad, dtype=torch.float32, device=dev)
        ops.hess_rhs(logits, idx_b, c0, gc, delta, c_pad, hess_sqrt)
        width, ld = C, c_pad
        for l in range(L - 1, -1, -1):                       # the KFAC backward, slabs kept
            gz = ops.spmm(at, delta)
            deltas[l], gzs[l] = delta, gz
            if l > 0:
                d_prev = dims[l - 1]
                nxt = torch.mm(gz.view(n * gc, ld)[:, :width], Ws[l])
                ops.relu_mask_mul(nxt, Hs[l], gc)
                delta = nxt.view(n, gc * d_prev)
                width = ld = d_prev
        dbar = None
        for l in range(L):                                   # its adjoint, bottom up
            width = dims[l]
            ld = c_pad if l == L - 1 else width
            gz_rows = gzs[l].view(n * gc, ld)
            gzbar = torch.zeros_like(gz_rows)
            gzbar[:, :width] = torch.mm(gz_rows[:, :width], 2.0 * Gbar[l])
            if l > 0:
                t = dbar.view(n * gc, dims[l - 1])
                ops.relu_mask_mul(t, Hs[l], gc)
                gzbar[:, :width] += torch.mm(t, Ws[l].t())
            gzbar_slab = gzbar.view(n, gc * ld)
            edge_dots(gzbar_slab, deltas[l], gc * ld)
            dbar = ops.spmm(g.ahat, gzbar_slab)
            deltas[l] = gzs[l] = None
        # top: delta_L[idx[m], c, :] = v_c(f_m)  ->  fbar_m += sum_c J_{v_c}(f_m)^T dbar_L[idx[m], c, :]
        cot = dbar.view(n, gc, c_pad)[idx_b][:, :, :C]
        step = max(1, (64 << 20) // max(1, C * C * 4))
        for s in range(s0, s1, step):
            e = min(s1, s + step)
            fs = f_train[s:e].detach().requires_grad_(True)
            with torch.enable_grad():
                V = hess_sqrt_columns(fs, hess_sqrt)[:, c0:c0 + gc, :]
                (gf,) = torch.autograd.grad((V * cot[s - s0:e - s0]).sum(), fs)
            fbar[s:e] += gf
        del dbar, cot

    # ---- forward part (single right-hand side)
    p = torch.softmax(f_train, dim=1)
    p[torch.arange(M, device=dev), y] -= 1.0                 # d loss / d f
    pbar = torch.zeros(n, c_pad, dtype=torch.float32, device=dev)
    pbar[:, :C].index_add_(0, idx, fbar - p)                 # marglik = -loss + ...
    for l in range(L - 1, -1, -1):
        width = dims[l]
        ld = pbar.shape[1]
        z = torch.zeros(n, ld, dtype=torch.float32, device=dev)
        z[:, :width] = torch.mm(Hs[l], Ws[l].t()) if bs[l] is None else torch.addmm(bs[l], Hs[l], Ws[l].t())
        edge_dots(z, pbar, ld)                               # Â enters as P_l = Â Z_l
        if l > 0:
            zbar = ops.spmm(at, pbar)
            hbar = torch.mm(zbar[:, :width], Ws[l])
            hbar += torch.mm(Hs[l], Abar[l]) * (2.0 * len(batches) / M)     # A_l = sum_b H_{l-1}^T H_{l-1} / N
            ops.relu_mask_mul(hbar, Hs[l], 1)
            pbar = hbar

    # ---- through the normalisation to A
    st = gt * at.val
    both = torch.zeros(n, dtype=torch.float32, device=dev)
    both.index_add_(0, rows.long(), st)
    both.index_add_(0, cols.long(), st)
    r = -0.5 * both / g.deg.to(torch.float32)
    grad_edges = st + r[rows.long()]
    grad_edges[rows == cols] = 0.0                           # fill_diagonal_(1): constant
    grad_cand = None
    if gc_t is not None:
        grad_cand = gc_t * g.dis[cand_rows.long()] * g.dis[cand_cols.long()] + r[cand_rows.long()]
        grad_cand[cand_rows == cand_cols] = 0.0
    return EdgeGradient(marglik.float(), loss, rows, cols, grad_edges, grad_cand, kron.kfacs)


class _EdgeMarglik(torch.autograd.Function):
    @staticmethod
    def forward(ctx, edge_param, model, idx, y, prior_precision, hess_sqrt):
        res = marglik_edge_grad(model, idx, y, prior_precision, hess_sqrt)
        ctx.save_for_backward(res.grad_edges)
        return res.marglik.clone()

    @staticmethod
    def backward(ctx, grad_out):
        (ge,) = ctx.saved_tensors
        return grad_out * ge, None, None, None, None, None


def log_marginal_likelihood_of_edges(model: SparseGCN, idx, y, edge_param: torch.Tensor,
                                     prior_precision=1.0, hess_sqrt: str = "reference") -> torch.Tensor:
    """The reference's usage pattern on a sparse parameter: ``edge_param`` is a leaf tensor with one
    entry per edge of A (``graph.ahat_t`` order; its values are not read — the straight-through
    estimator evaluates at the binarised graph), and
    ``(-log_marginal_likelihood_of_edges(...)).backward()`` fills ``edge_param.grad`` the way
    ``neg_marglik.backward()`` fills ``model.adj.grad`` (gnn/marglik_training.py:213)."""
    if edge_param.numel() != model.graph.ahat_t.nnz:
        raise ValueError("edge_param needs one entry per edge of A (graph.ahat_t.nnz, self loops included)")
    return _EdgeMarglik.apply(edge_param, model, idx, y, prior_precision, hess_sqrt)


# ------------------------------------------------------------------------------------------------
# straight-through edge scores: the sparse analogue of STEGCN's dense ``adj`` parameter
# ------------------------------------------------------------------------------------------------
class EdgeScores(torch.nn.Module):
    """One real-valued score per TRACKED entry A[m, k] (m != k) — the existing edges (score 1) and a
    caller-chosen set of candidate entries (score 0) — binarised by a threshold in the forward pass and
    updated with the gradient of the binarised graph (straight-through), like the reference's dense
    ``adj`` parameter (STEGCN, gnn/models/models.py:65-118; BinarizeSTE, gnn/models/utils.py:42-86).
    The reference tracks all N^2 entries; here the untracked ones stay 0 forever."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, candidates: torch.Tensor | None = None,
                 threshold: float = 0.5):
        super().__init__()
        n = int(num_nodes)
        dev = edge_index.device
        parts = [(edge_index.to(torch.int64), 1.0)]
        if candidates is not None and candidates.numel() > 0:
            parts.append((candidates.to(torch.int64).to(dev), 0.0))
        key = torch.cat([p[0][0] * n + p[0][1] for p in parts])
        val = torch.cat([torch.full((p[0].shape[1],), p[1], device=dev) for p in parts])
        off = (key // n) != (key % n)                               # the diagonal is a constant 1
        key, val = key[off], val[off]
        # unique entries sorted by (m, k); an entry listed as edge and as candidate is an edge
        order = torch.argsort(key * 2 + (1 - val.long()), stable=True)
        key, val = key[order], val[order]
        first = torch.ones_like(key, dtype=torch.bool)
        first[1:] = key[1:] != key[:-1]
        key, val = key[first], val[first]
        self.n = n
        self.threshold = float(threshold)
        self.register_buffer("entries", torch.stack([key // n, key % n]))
        self.score = torch.nn.Parameter(val.to(torch.float32))

    @property
    def active(self) -> torch.Tensor:
        return self.score.detach() > self.threshold

    def edge_index(self) -> torch.Tensor:
        return self.entries[:, self.active]

    def graph(self):
        from .graph import Graph
        return Graph.from_edge_index(self.edge_index().contiguous(), self.n)

    def neg_marglik_step(self, model: SparseGCN, idx, y, optimizer: torch.optim.Optimizer, prior_precision=1.0,
                         hess_sqrt: str = "reference", grad_norm: bool = False,
                         batch_size: int | None = None) -> torch.Tensor:
        """One ``adj_optimizer.zero_grad(); neg_marglik.backward(); adj_optimizer.step()`` of
        gnn/marglik_training.py:206-220 on the tracked entries, then the model's graph is rebuilt from
        the re-binarised scores.  Returns the marglik BEFORE the step."""
        act = self.active
        res = marglik_edge_grad(model, idx, y, prior_precision, hess_sqrt,
                                candidates=self.entries[:, ~act] if bool((~act).any()) else None,
                                batch_size=batch_size)
        off = res.rows != res.cols                              # CSR order == (m, k) order of the active entries
        grad = torch.zeros_like(self.score)
        grad[act] = -res.grad_edges[off]
        if res.grad_candidates is not None:
            grad[~act] = -res.grad_candidates
        optimizer.zero_grad()
        self.score.grad = grad
        if grad_norm:
            torch.nn.utils.clip_grad_norm_([self.score], max_norm=1.0)
        optimizer.step()
        if not torch.equal(self.active, act):
            model.graph = self.graph()
        return res.marglik
