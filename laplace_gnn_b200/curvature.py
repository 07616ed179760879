"""B200GGN — curvature backend implementing the reference's ``CurvatureInterface.kron()/diag()``
contract (laplace/curvature/curvature.py:236-289) for the sparse GCN, on hand-written sm_100a
kernels.  It replaces ``CurvlinopsGGN.kron`` (laplace/curvature/curvlinops.py:77-108) and the
vendored ``KFACLinearOperator._compute_kfac`` (curvlinops/kfac.py:540-875) for this model family.

Select it exactly like any other backend::

    la = Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron",
                 backend=B200GGN, backend_kwargs={"hess_sqrt": "reference"})
    la.fit(train_loader); la.log_marginal_likelihood()

What one ``kron(idx, y, N)`` call does on the device (M = len(idx), all graph nodes take part,
SURVEY §0-T2):

  forward      Z_l = H_{l-1} W_l^T + b_l (cuBLAS fp32, a plain library GEMM), P_l = Â Z_l (SpMM
               kernel, relu fused for l < L), once — the reference runs three forwards
  A_l          = H_{l-1}^T H_{l-1} / M * (M / N)                         (SYRK kernel)
  loss         = sum CE(P_L[idx], y)                                      (fused softmax kernel)
  backward     for a group of g Hessian-sqrt columns at a time (g sized to the HBM budget):
               delta_L[idx, c, :] = v_{m,c}  (fused softmax + Hessian-sqrt kernel, both modes),
               gZ_l = Â^T delta_l as ONE multi-RHS SpMM of width g*d_l, G_l += gZ_l^T gZ_l (SYRK,
               K = N*g), delta_{l-1} = (gZ_l W_l) ⊙ 1[H_{l-1} > 0]  (cuBLAS GEMM + mask kernel)
  packing      [[G_1, A_1], [G_1], [G_2, A_2], [G_2], ...]  (curvlinops.py:55-75)

``hess_sqrt="reference"`` (default) reproduces the fork's non-detached Hessian square root
(curvlinops/kfac.py:631-661); ``"ggn"`` is the textbook GGN of upstream
curvlinops / asdl / backpack.  Returned tensors are fresh, detached, fp32, on the model's device.
"""
from __future__ import annotations

import contextlib
import sys
from typing import Any

import torch
from torch import nn

from . import ops
from .gcn import SparseGCN


# ------------------------------------------------------------------------------------------------
# persistent slab workspace
# ------------------------------------------------------------------------------------------------
# The two multi-RHS slabs of a pass take up to 45 % of HBM (2 x 40 GB on the products shape).  A new
# backend is created per fit (per epoch in the reference's loop); handing such blocks back to torch's
# caching allocator between fits lets the next forward carve its activations out of them, the next pass
# then finds no block large enough, and every fit pays a fresh 30-40 GB cudaMalloc (measured: one device
# allocation per step, 163 GB reserved after six steps, 30-140 ms of idle device per fit).  The slabs
# are therefore kept, per (device, stream, lane, slot), and only ever grow; release_workspace() frees them.
_SLABS: dict = {}


def _slab(dev, lane: int, slot: int, numel: int) -> torch.Tensor:
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0
    key = (dev.type, dev.index, stream, lane, slot)
    t = _SLABS.get(key)
    if t is None or t.numel() < numel:
        _SLABS.pop(key, None)
        del t
        t = torch.empty(max(numel, 1), dtype=torch.float32, device=dev)
        _SLABS[key] = t
    return t[:max(numel, 1)]


_LANE_STREAMS: dict = {}


def _lane_streams(dev, lanes: int):
    """The side streams of the overlapped row-partitioned backward, created once per device: per-stream
    scratch (SYRK partials, ops._workspace) is keyed by stream and would otherwise be re-created for every
    stream torch's pool hands out."""
    have = _LANE_STREAMS.setdefault((dev.type, dev.index), [])
    while len(have) < lanes:
        have.append(torch.cuda.Stream(dev))
    return have[:lanes]


def workspace_bytes(dev=None) -> int:
    """Bytes held by the persistent slab workspace (on ``dev``, or everywhere)."""
    return sum(t.numel() * 4 for k, t in _SLABS.items()
               if dev is None or (k[0], k[1]) == (torch.device(dev).type, torch.device(dev).index))


def release_workspace() -> None:
    """Free the persistent multi-RHS slabs (they are re-created by the next ``kron`` call)."""
    _SLABS.clear()


class CurvatureInterfaceLite:
    """Attribute-compatible stand-in for ``laplace.curvature.CurvatureInterface.__init__``
    (curvature.py:46-83) used when the ``laplace`` package is not installed."""

    def __init__(self, model: nn.Module, likelihood: str, last_layer: bool = False,
                 subnetwork_indices=None, dict_key_x: str = "input_ids", dict_key_y: str = "labels"):
        if likelihood not in ("classification", "regression"):
            raise ValueError(f"Invalid likelihood type {likelihood}")
        self.likelihood = likelihood
        self.model = model
        self.last_layer = last_layer
        self.subnetwork_indices = subnetwork_indices
        self.dict_key_x, self.dict_key_y = dict_key_x, dict_key_y
        if likelihood == "regression":
            self.lossfunc = nn.MSELoss(reduction="sum")
            self.factor = 0.5
        else:
            self.lossfunc = nn.CrossEntropyLoss(reduction="sum")
            self.factor = 1.0
        self.params, self.params_dict = [], {}
        for k, v in model.named_parameters():
            if v.requires_grad and "adj" not in k and "norms" not in k:
                self.params.append(v)
                self.params_dict[k] = v
        self.buffers_dict = dict(model.named_buffers())


def _kron_class():
    """The ``Kron`` of whichever ``laplace`` package is driving us, else the local stand-in."""
    mod = sys.modules.get("laplace.utils")
    if mod is not None and hasattr(mod, "Kron"):
        return mod.Kron
    from .kron import Kron
    return Kron


class _DeferredActivations(list):
    """[X, H_1, ..., H_{L-1}] of ALL nodes in the column-parallel multi-GPU backward; entry l is produced by its
    thunk (wait for the asynchronous all-gather, copy into natural node order) when it is first indexed."""

    def __init__(self, first, n: int):
        super().__init__([first] + [None] * (n - 1))
        self._thunks: dict = {}

    def defer(self, l: int, thunk) -> None:
        self._thunks[l] = thunk

    def __getitem__(self, i):
        if isinstance(i, int):
            j = i if i >= 0 else len(self) + i
            thunk = self._thunks.pop(j, None)
            if thunk is not None:
                list.__setitem__(self, j, thunk())
        return list.__getitem__(self, i)

    def __iter__(self):
        for j in range(len(self)):
            yield self[j]


class _Whole:
    """Backward layout: every graph row is local (single device, or the column-parallel backward)."""

    def __init__(self, graph):
        self.csr_t = graph.ahat_t
        self.n_local = self.total_rows = graph.n
        self.slot0 = 0
        self.communicates = False
        self.csr_t_top = None        # Â^T with the edges from non-train rows zeroed (set per kron call)
        self.split_t = None          # Â^T with its hub rows cut into pieces (graph.SplitCSR), for the unit SpMM
        self.hess_stats = None       # softmax statistics for the on-the-fly output-layer SpMM (set per kron call)

    def gather(self, slab):
        pass

    def agree_min(self, value: int) -> int:
        return value


class _Rows:
    """Backward layout: this rank's row block; SpMM inputs are all-gathered padded slabs."""

    def __init__(self, part):
        self.part = part
        self.csr_t = part.ahat_t
        self.n_local, self.total_rows, self.slot0 = part.n_local, part.total_rows, part.slot0
        self.communicates = True
        self.csr_t_top = None
        self.split_t = None
        self.hess_stats = None
        self.plans = {}              # (layer, even-slot layout) -> _RaggedPlan, built per kron call

    def gather(self, slab):
        self.part.exchange_for_spmm(slab)

    def agree_min(self, value: int) -> int:
        """Collectives need the same column grouping on every rank."""
        t = torch.tensor([value], dtype=torch.int64, device=self.part.ahat.rowptr.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=self.part.pg)
        return int(t.item())


class _RaggedPlan:
    """Where the unit-compacted rows of one hidden layer live in the buffer the ranks all-gather (rows layout).

    The live units of a node are the same for every Hessian-sqrt column and every column group (relu' of H_l), so
    the plan is made once per fit and layer: rank r packs its rows back to back from slot r*cap on (``row_first``:
    r*cap + exclusive scan of the rows' slot counts; ``cap`` = the largest per-rank total, agreed by one MAX
    all-reduce), and the header words (mask, ABSOLUTE first slot per 32 units) of all nodes are all-gathered once —
    8 bytes per 32 units against the g*4*32 bytes of a group's values."""

    def __init__(self, part, act, g: int, hdr: torch.Tensor):
        dev = act.device
        n_loc, h = int(act.shape[0]), int(act.shape[1])
        slots = ops.unit_row_slots(act, g)
        scan = torch.cumsum(slots, 0)
        top = scan[-1:].clone() if n_loc > 0 else torch.zeros(1, dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(top, op=torch.distributed.ReduceOp.MAX, group=part.pg)
        self.cap = max((int(top.item()) + 3) // 4 * 4, 4)          # slots per rank, 16-byte aligned for every g
        self.world = part.world
        self.ok = self.world * self.cap < (1 << 32)                 # header slots are 32-bit
        self.hdr = hdr
        self.live = None
        if not self.ok:
            return
        self.row_first = (scan - slots + part.rank * self.cap).contiguous()
        hdr.zero_()                                                 # pad rows: empty masks
        ops.unit_pack_ragged(None, act, g, self.row_first, None, hdr[part.slot0:part.slot0 + n_loc])
        part.all_gather_slab(hdr)
        if ops.PROFILE is not None:                                 # live units of every node, for the byte accounting
            live = torch.zeros(part.total_rows, dtype=torch.int64, device=dev)
            live[part.slot0:part.slot0 + n_loc] = (act > 0).sum(1)
            part.all_gather_slab(live)
            self.live = live

    def floats(self, g: int) -> int:
        return self.world * self.cap * g


class _B200KFAC:
    """kron()/diag() on the B200 kernels; mixed into a CurvatureInterface subclass."""
    _lanes_on_cpu = False        # tests: walk the two-lane interleaving of the rows layout with the CPU double too

    def _b200_setup(self, hess_sqrt="reference", differentiable=False, process_group=None, rhs_tile_bytes=None,
                    syrk_impl="auto", backward_parallel="columns", overlap=True, fused_gemm=True, fused_linear=True,
                    cache_input_factor=False, _shared_cache=None, diag_mode="exact",
                    # data layout of the multi-RHS slabs (defaults = what the benchmark runs; off-switches for A/B)
                    unit_slabs=True, unit_min_width=1024, unit_even_groups=True, unit_hub_split=True,
                    fused_hess_spmm=True,
                    # multi-GPU
                    shard_eigh=True, defer_gathers=True, sparse_halo=False, unit_rows=True,
                    rows_hess_stats=False):
        if hess_sqrt not in ("reference", "ggn"):
            raise ValueError(f"hess_sqrt must be 'reference' or 'ggn', got {hess_sqrt!r}")
        if diag_mode not in ("exact", "node_factorised"):
            raise ValueError(f"diag_mode must be 'exact' or 'node_factorised', got {diag_mode!r}")
        if backward_parallel not in ("rows", "columns"):
            raise ValueError(f"backward_parallel must be 'rows' or 'columns', got {backward_parallel!r}")
        if differentiable:
            raise NotImplementedError(
                "differentiable=True (autograd graphs through the factors) is not provided; the gradient of "
                "the marginal likelihood w.r.t. the adjacency entries is laplace_gnn_b200.structure."
                "marglik_edge_grad (SURVEY §8f row 3)")
        if self.likelihood != "classification":
            raise ValueError("B200GGN implements the classification (softmax CE) hot path only")
        if not isinstance(self.model, SparseGCN):
            raise TypeError("B200GGN needs a laplace_gnn_b200.SparseGCN model")
        if self.last_layer or self.subnetwork_indices is not None:
            raise NotImplementedError("last_layer / subnetwork Laplace are outside the hot path")
        self.hess_sqrt = hess_sqrt
        self.diag_mode = diag_mode
        self._layer_hook = None              # diag.diag_ggn_node_factorised taps gZ_l of every column group here
        self.process_group = process_group
        self.rhs_tile_bytes = rhs_tile_bytes
        self.syrk_impl = syrk_impl
        self.backward_parallel = backward_parallel
        self.overlap = bool(overlap)
        self.fused_gemm = bool(fused_gemm)
        # the forward linear layers Z_l = H_{l-1} W_l^T + b_l on the same tcgen05 kernel (lgnn_gemm_bias_f32);
        # False: cuBLAS fp32 (torch.addmm)
        self.fused_linear = bool(fused_linear)
        # column-parallel multi-GPU backward: all-gathers of the hidden activations issued asynchronously and waited
        # for on first use; False: gathered in front of the backward (round-1 order)
        self.defer_gathers = bool(defer_gathers)
        self.skip_zero_rows = True
        # unit-compacted slabs below the output layer (csrc/spmm_units.cu): the relu' mask is shared by all
        # columns of a node, so the slab rows keep only their live hidden units and the SpMM gathers about
        # half the bytes.  Column groups are then padded to a multiple of 4 with all-zero right-hand sides.
        self.unit_slabs = bool(unit_slabs)
        self.unit_min_width = int(unit_min_width)   # narrower slabs stay dense (one row is a few hundred bytes anyway)
        # column groups of 2, 6, 10 or 14 as well (csrc/spmm_units_even.cu): a rank of the 8-GPU column split
        # owns 6 of the products shape's 47 columns and would otherwise carry 8 (g = 6: 93.7 ms against 111 ms per
        # hidden layer, profiles/r2a_units_lab.txt; bit-identical to the dense SpMM, tests/test_gpu_units_even.py)
        self.unit_even_groups = bool(unit_even_groups)
        # with a process group the stand-in KronLaplace spreads the factor eigendecompositions over the ranks
        # (kron.Kron.decompose: one in-place all-gather hands every rank the same eigenpairs, bit for bit — NCCL test
        # tests/test_gpu_dist.py); shard_eigh=False keeps them replicated
        self.shard_eigh = bool(shard_eigh)
        # power-law graphs: the unit SpMM gives one warp group a whole row, so graphs with rows beyond
        # unit_row_limit non-zeros keep dense slabs — unless unit_hub_split cuts those rows into pieces
        # (graph.split_hub_rows; the pieces are summed by a small SpMM afterwards).  R-MAT 2^20 - 2^22, degree
        # 16 - 64: 13 - 19 % faster fits than dense slabs, same marglik to the last digit (profiles/r2h_rmat_sweep.txt)
        self.unit_hub_split = bool(unit_hub_split)
        # output-layer SpMM with its Hessian-sqrt right-hand sides rebuilt per edge from five softmax vectors per
        # node (csrc/spmm_hess.cu): 576 instead of 3072 gathered bytes per edge at g = 16, C = 47, and no
        # lgnn_hess_rhs_f32 pass.  Measured on the products shape: 18.9 ms against 43.6 ms per 16-column group, the
        # fit 1,678 against 1,786 ms (profiles/r2b_hess_spmm_lab.txt); shapes it does not take fall back
        self.fused_hess_spmm = bool(fused_hess_spmm)
        # row-partitioned passes: when fewer than half of the other ranks' rows are referenced (a graph with
        # locality, partitioned), exchange only those halo rows (all-to-all) instead of all-gathering whole slabs
        # (dist.RowPartition.exchange_for_spmm).  OFF until the all-to-all has run over NCCL (gloo-tested)
        self.sparse_halo = bool(sparse_halo)
        # unit-compacted slabs in the "rows" layout too: a rank packs the live units of its row block back to back
        # (ops.unit_pack_ragged), the packed rows are what the all-gather moves (about half the dense slab) and what
        # the SpMM then gathers from (_RaggedPlan).  False: dense slabs travel (the round-1 rows layout)
        self.unit_rows = bool(unit_rows)
        # rows layout, output layer: softmax statistics all-gathered once per fit and the right-hand sides rebuilt per
        # edge (as on one device) instead of dense right-hand-side slabs exchanged per column group.  565 against
        # 622 ms per products fit at 4 GPUs — but OPT-IN: the NCCL parity test passed with it at 4 ranks and failed
        # once at 2 ranks (factors off by 1.9e-4) in the round's last GPU seconds, unresolved
        self.rows_hess_stats = bool(rows_hess_stats)
        self.unit_row_limit = 4096
        # A_0 = X^T X does not depend on the weights: with cache_input_factor the raw Gram matrix of
        # this rank's feature rows is kept (per backend, or in a dict shared across backends by the
        # epoch loop) and only rescaled per call
        self.cache_input_factor = bool(cache_input_factor)
        self._cache = _shared_cache if _shared_cache is not None else {}
        self.n_outputs = self.model.out_channels
        self.last_stats: dict[str, Any] = {}

    # ------------------------------------------------------------------ helpers
    def _layers(self):
        Ws, bs = [], []
        for conv in self.model.convs:
            w = conv.lin.weight.detach()
            if w.dtype != torch.float32:
                raise TypeError("B200GGN computes in float32; cast the model to float32")
            Ws.append(w.contiguous())
            bs.append(None if conv.lin.bias is None else conv.lin.bias.detach().contiguous())
        return Ws, bs

    def _forward(self, Ws, bs, part=None):
        """Eval-mode forward.  Returns Hs = [X, H_1, ..., H_{L-1}] and the logits P_L — all graph
        rows on one device, this rank's row block when ``part`` is given (each layer then
        all-gathers Z_l, the slab its SpMM reads, over the process group)."""
        g = self.model.graph
        h = self.model.X
        if h.dtype != torch.float32 or not h.is_contiguous():
            h = h.float().contiguous()
        x_rows = getattr(self.model, "x_rows", None)
        if x_rows is not None:                        # the model holds this rank's node block only (sharded ingest)
            if part is None or x_rows != (part.lo, part.hi):
                raise ValueError(f"the model holds rows {x_rows} of X; the fit needs "
                                 f"{'all rows' if part is None else (part.lo, part.hi)}")
        elif part is not None:
            h = h[part.lo:part.hi]
        Hs = [h]
        L = len(Ws)
        self._fwd_out = []          # row-partitioned forward: the padded slabs holding this rank's P_l / H_l rows
        # Z_l = H_{l-1} W_l^T + b_l on the fused 3xTF32 tcgen05 GEMM (bias in its epilogue) wherever it takes the shape
        # (d_in, d_out <= 256); cuBLAS fp32 otherwise (e.g. the 1,433 input features of the Cora shape)
        lin = [self._linear_operands(Ws[l], bs[l]) if (self.fused_gemm and self.fused_linear and h.is_cuda) else None for l in range(L)]
        for l in range(L):
            d_out = Ws[l].shape[0]
            if part is None:
                # Z_l and P_l / H_l live in the persistent workspace (one pair per layer): the activations are
                # rewritten by every pass and a fresh 2.5 GB allocation per layer per fit is a cudaMalloc each
                n_rows = h.shape[0]
                ldz = (d_out + 3) // 4 * 4     # odd class count: pad the pitch so the SpMM takes its 128-bit path
                if lin[l] is not None and self._tma_ok(h):   # the kernel writes its whole (zero-padded) output width
                    wp, bias_p = lin[l]
                    zbuf = _slab(h.device, 1000 + l, 0, n_rows * wp.n).view(n_rows, wp.n)
                    ops.gemm_bias(h, wp, bias_p, zbuf, m_rows=n_rows)
                    z = zbuf[:, :ldz]
                else:
                    z = _slab(h.device, 1000 + l, 0, n_rows * ldz).view(n_rows, ldz)
                    with ops.timed("gemm_fwd", d_out, 2.0 * n_rows * Ws[l].numel()):
                        if ldz == d_out:
                            if bs[l] is None:
                                torch.mm(h, Ws[l].t(), out=z)
                            else:
                                torch.addmm(bs[l], h, Ws[l].t(), out=z)
                        else:
                            z[:, :d_out] = torch.mm(h, Ws[l].t()) if bs[l] is None else torch.addmm(bs[l], h, Ws[l].t())
                            z[:, d_out:] = 0
                rows_out = n_rows + g.extra_rows()          # hub rows of power-law graphs come back in pieces
                out = _slab(h.device, 1000 + l, 1, rows_out * ldz).view(rows_out, ldz)
                h = g.propagate(z, relu=(l < L - 1), out=out)[:, :d_out]
            else:
                # row block of this rank: Z_l goes straight into this rank's slot of the padded all-gather slab,
                # P_l / H_l into its slot of a second one (the column-parallel backward all-gathers that one in
                # place).  Both persistent, like the single-device pair above.
                ldz = (d_out + 3) // 4 * 4
                slab = _slab(h.device, 1100 + l, 0, part.total_rows * ldz).view(part.total_rows, ldz)
                z = slab[part.slot0:part.slot0 + part.n_local]
                if lin[l] is not None and lin[l][0].n == ldz and self._tma_ok(h):   # (a padded width would overrun the pitch)
                    ops.gemm_bias(h, lin[l][0], lin[l][1], z, m_rows=h.shape[0])
                else:
                    with ops.timed("gemm_fwd", d_out, 2.0 * h.shape[0] * Ws[l].numel()):
                        if ldz == d_out:
                            if bs[l] is None:
                                torch.mm(h, Ws[l].t(), out=z)
                            else:
                                torch.addmm(bs[l], h, Ws[l].t(), out=z)
                        else:
                            z[:, :d_out] = torch.mm(h, Ws[l].t()) if bs[l] is None else torch.addmm(bs[l], h, Ws[l].t())
                            z[:, d_out:] = 0
                with ops.timed("allgather", d_out, 4.0 * part.total_rows * ldz):
                    part.exchange_for_spmm(slab)          # whole slab, or the halo rows only when the halo is sparse
                out_slab = _slab(h.device, 1100 + l, 1, part.total_rows * ldz).view(part.total_rows, ldz)
                self._fwd_out.append(out_slab)
                h = ops.spmm(part.ahat, slab, relu=(l < L - 1),
                             out=out_slab[part.slot0:part.slot0 + part.n_local])[:, :d_out]
            if l < L - 1:
                Hs.append(h)
        return Hs, h

    @staticmethod
    def _tma_ok(t: torch.Tensor) -> bool:
        """The fused GEMM reads its left operand through TMA: 16-byte aligned rows (e.g. not a 7-feature matrix)."""
        return t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0 and t.stride(1) == 1

    def _linear_operands(self, W: torch.Tensor, b):
        """(prepared W^T, zero-padded bias) for ops.gemm_bias, or None when the fused GEMM does not take the layer."""
        wp = ops.linear_prepare(W)
        if wp is None:
            return None
        bias_p = None
        if b is not None:
            bias_p = torch.zeros(wp.n, dtype=torch.float32, device=W.device)
            bias_p[: b.numel()] = b
        return wp, bias_p

    def _group_size(self, rows_in: int, rows_out: int, dmax: int, C: int, device) -> int:
        budget = self.rhs_tile_bytes
        if budget is None:
            if device.type == "cuda":
                free, total = torch.cuda.mem_get_info(device)
                # blocks torch's caching allocator holds but has not handed out are ours to reuse
                free += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
                free += workspace_bytes(device)          # the slabs kept from the previous pass are reused
                budget = min(int(0.6 * free), int(0.45 * total))
            else:
                budget = 1 << 30
        g = int(budget // (max(rows_in + rows_out, 1) * dmax * 4))
        return max(1, min(C, g))

    def _partition(self):
        """RowPartition of the model's graph for this backend's process group (built once)."""
        if self.process_group is None:
            return None
        # cached on the graph: a new Laplace object (one per epoch in the reference's loop) makes a new
        # backend, the row slices of the CSR do not change
        cache = self.model.graph.meta.setdefault("_partitions", {})
        part = cache.get(id(self.process_group))
        if part is None or part.pg is not self.process_group:
            from .dist import RowPartition
            part = RowPartition.build(self.model.graph, self.process_group)
            cache[id(self.process_group)] = part
        part.sparse_halo_below = 0.5 if self.sparse_halo else 0.0
        return part

    def _chain(self, lay, logits, idx, Hs, Ws, Wp, c0, gc, gq, buf_a, buf_b, G, hdr=None, scratch=None):
        """One group of ``gc`` Hessian-sqrt columns pushed down all layers; a generator that yields
        after each layer so that two groups can be interleaved on two streams (their all-gathers
        then overlap the other group's SpMM).  ``gq >= gc`` is the group's width in the slabs: the
        columns beyond ``gc`` are all-zero right-hand sides (padding for the unit-compacted layout).

        ``buf_a`` / ``buf_b`` alternate as SpMM input / output.  Below the output layer the input slab
        delta = (gZ W) ⊙ relu' is about half zeros, in a pattern the columns of a node share: when every
        row is local (single GPU, or the column-parallel backward) it is compacted in place to the live
        units of each node (``ops.unit_pack``, which applies the mask itself) and the SpMM gathers only
        those (``ops.spmm_units``).  (A per-element compression of the same slabs was 2x slower than the dense
        gather — profiles/r1g_pack_lab.txt — and has left the tree.)  In the rows layout (``scratch`` given) the GEMM
        writes this rank's dense rows into ``scratch``, they are packed back to back into ``buf_a`` at the slots of
        the layer's ``_RaggedPlan``, and the all-gather moves the packed rows only."""
        L = len(Ws)
        C = logits.shape[1]
        dims = [w.shape[0] for w in Ws]                 # d_1 .. d_L (d_L = C)
        c_pad = (C + 3) // 4 * 4
        n_loc, n_in, slot0 = lay.n_local, lay.total_rows, lay.slot0
        P, Q = buf_a, buf_b                             # P: SpMM input, Q: SpMM output
        slab = P[: n_in * gq * c_pad].view(n_in, gq * c_pad)
        delta = slab[slot0:slot0 + n_loc]
        on_the_fly = lay.hess_stats is not None and ops.spmm_hess_supported(C, gq)
        if not on_the_fly:
            with ops.timed("hess_rhs", gc):
                delta.zero_()
                ops.hess_rhs(logits, idx, c0, gc, delta, c_pad, self.hess_sqrt)
        width, ld = C, c_pad
        units = None
        for l in range(L - 1, -1, -1):
            gz = Q[: n_loc * gq * ld].view(n_loc, gq * ld)
            if units is not None and lay.split_t is not None:
                sp = lay.split_t                    # hub rows in pieces: extra output rows behind the slab
                gz_all = Q[: (n_loc + sp.n_extra) * gq * ld].view(n_loc + sp.n_extra, gq * ld)
                ops.spmm_units(sp.csr, units, out=gz_all)
                sp.finish(gz_all, gq * ld)
                self._n_unit_spmm += 1
            elif units is not None:
                ops.spmm_units(lay.csr_t, units, out=gz)
                self._n_unit_spmm += 1
            elif l == L - 1 and on_the_fly:
                # output layer, right-hand sides rebuilt per edge from the nodes' softmax statistics: no slab
                ops.spmm_hess(lay.csr_t_top if lay.csr_t_top is not None else lay.csr_t, lay.hess_stats, C, c0, gc,
                              gq, out=gz)
            else:
                with ops.timed("allgather", gq * ld, 4.0 * n_in * gq * ld):
                    lay.gather(slab)
                # output layer: the slab is zero outside the batch's train rows -> no gather for those edges
                ops.spmm(lay.csr_t_top if (l == L - 1 and lay.csr_t_top is not None) else lay.csr_t, slab, out=gz)
            gz_rows = gz.view(n_loc * gq, ld)
            ops.syrk(gz_rows, n=width, alpha=1.0, beta=1.0, out=G[l], impl=self._impl(width))
            if self._layer_hook is not None:
                self._layer_hook(l, gz, gq, ld, width)
            if l > 0:
                d_prev = dims[l - 1]
                slab = P[: n_in * gq * d_prev].view(n_in, gq * d_prev)
                nxt = slab[slot0:slot0 + n_loc].view(n_loc * gq, d_prev)
                plan = lay.plans.get((l, gq % 4 != 0)) if (scratch is not None and lay.communicates) else None
                if plan is not None and not (plan.ok and self._can_unit(lay, gq, d_prev)):
                    plan = None
                if plan is not None:
                    nxt = scratch[: n_loc * gq * d_prev].view(n_loc * gq, d_prev)
                to_units = plan is not None or (hdr is not None and not lay.communicates and
                                                self._can_unit(lay, gq, d_prev))
                act = None if to_units else Hs[l]           # unit_pack drops the dead units: no mask needed
                if Wp[l] is not None:      # fused 3xTF32 tensor-core GEMM (+ relu' mask)
                    ops.gemm_mask(gz_rows, Wp[l], act, gq, out=nxt, m_rows=n_loc * gq)
                else:                      # shapes the fused kernel does not take: cuBLAS + mask kernel
                    with ops.timed("gemm_bwd", d_prev, 2.0 * n_loc * gq * width * d_prev):
                        torch.mm(gz_rows[:, :width], Ws[l], out=nxt)
                    if act is not None:
                        with ops.timed("relu_mask", d_prev, 2.0 * n_loc * gq * d_prev * 4):
                            ops.relu_mask_mul(nxt, act, gq)
                width, ld = d_prev, d_prev
                if plan is not None:
                    flat = P[: plan.floats(gq)]
                    ops.unit_pack_ragged(nxt.view(n_loc, gq * d_prev), Hs[l][:, :d_prev], gq, plan.row_first, flat, None)
                    with ops.timed("allgather", gq * d_prev, 4.0 * plan.floats(gq)):
                        lay.part.all_gather_slab(flat)
                    units = ops.UnitSlab(n_in, gq, d_prev, flat, plan.hdr, None, plan.live, ragged=True)
                else:
                    units = ops.unit_pack(slab, Hs[l], gq, hdr=hdr) if to_units else None
            yield

    def _units_possible(self, lay) -> bool:
        """Unit-compacted slabs need a graph without hub rows (the kernel gives one warp a whole row; hub rows cut
        into pieces count) and, in the rows layout, the whole-slab exchange (``unit_rows``: ragged rows travel; with
        a sparse halo the ranks exchange dense halo rows instead)."""
        mx = lay.csr_t.max_row_nnz
        local = not lay.communicates or (self.unit_rows and not lay.part.sparse_halo)   # rows layout: ragged rows travel
        return (self.unit_slabs and local and mx is not None and
                (mx <= self.unit_row_limit or lay.split_t is not None))

    def _can_unit(self, lay, g: int, h: int) -> bool:
        return (self._units_possible(lay) and g * h >= self.unit_min_width and
                ops.unit_slabs_supported(g, h))

    def _backward_columns(self, lay, logits, idx, Hs, Ws, cols, G):
        """Multi-RHS KFAC backward for the Hessian-sqrt columns ``cols = (first, count)`` on the
        rows ``lay`` describes: G[l] += sum_c gZ_{l,c}^T gZ_{l,c}  (kfac.py:653-661, 777-817)."""
        C = logits.shape[1]
        dims = [w.shape[0] for w in Ws]
        c_pad = (C + 3) // 4 * 4
        dmax = max([c_pad] + dims[:-1])
        c_first, c_count = cols
        dev = logits.device
        n_loc, n_in = lay.n_local, lay.total_rows
        # two column groups in flight when a group waits for collectives (rows layout).  (With every row local the
        # same interleaving — one group's SYRK / GEMM under the other's SpMM — was measured 10 % SLOWER than one
        # group at a time, 1,972 against 1,786 ms per products fit: the persistent tensor-core CTAs and the SpMM's
        # CTAs time-slice the SMs instead of sharing them; profiles/r2a_lab_switches.txt.)
        lanes = 2 if (self.overlap and lay.communicates and (dev.type == "cuda" or self._lanes_on_cpu)) else 1
        ragged = lay.communicates and self._units_possible(lay)     # rows layout with unit-compacted rows travelling
        # columns the HBM budget allows (ragged: a third, local slab per lane for the GEMM's dense rows)
        room = self._group_size(lanes * n_in, lanes * n_loc * (2 if ragged else 1), dmax, 1 << 30, dev)
        grp = min(room, C)
        grp = lay.agree_min(max(1, min(grp, (c_count + lanes - 1) // lanes)))
        # unit-compacted slabs want groups of 4, 8, 12 or 16 columns (any even count with unit_even_groups):
        # the last group of a pass is padded with all-zero right-hand sides (47 classes -> 16 + 16 + 15(+1))
        q = 2 if self.unit_even_groups else 4
        per_lane = (max(c_count, 1) + lanes - 1) // lanes         # = c_count with one group in flight
        cap = min(room // q * q, 16)
        # equal groups instead of full ones plus a narrow rest: 24 columns (a rank of the 2-GPU split) travel as
        # 12 + 12, not 16 + 8 — the unit SpMM runs at 0.88 of the copy rate at g = 12 against 0.75 at g = 8
        n_eq = (per_lane + cap - 1) // cap if cap > 0 else 1
        cand = min(cap, ((per_lane + n_eq - 1) // n_eq + q - 1) // q * q)
        pad4 = (self._units_possible(lay) and cand >= q and
                any(self._can_unit(lay, cand, h) for h in dims[:-1]))
        if pad4:
            grp = cand
        if c_count <= 0:
            return grp, 0
        groups = [(c0, min(grp, c_first + c_count - c0)) for c0 in range(c_first, c_first + c_count, grp)]
        width_of = (lambda gc: (gc + q - 1) // q * q) if pad4 else (lambda gc: gc)
        ragged = ragged and pad4
        hdr = (torch.empty(n_in, max(max(dims[:-1]) // 32, 1), 2, dtype=torch.int32, device=dev)
               if pad4 and not ragged else None)
        hdrs = [hdr, torch.empty_like(hdr) if (hdr is not None and lanes > 1) else None]
        if ragged:
            # one plan per hidden layer and slot layout (every rank takes the same branches: widths, limits and the
            # agreed capacity are the same everywhere), made on the main stream before the lanes start
            for gq in sorted({width_of(gc) for _, gc in groups}):
                for l in range(1, len(Ws)):
                    h = dims[l - 1]
                    key = (l, gq % 4 != 0)
                    if key in lay.plans or not self._can_unit(lay, gq, h):
                        continue
                    words = _slab(dev, 1200 + l, int(key[1]), n_in * (h // 32) * 2).view(torch.int32)
                    lay.plans[key] = _RaggedPlan(lay.part, Hs[l][:, :h], gq, words.view(n_in, h // 32, 2))
            ragged = any(p.ok for p in lay.plans.values())
        # W_l [d_l, d_{l-1}] as resident tensor-core operands, once per pass
        Wp = [None] + [ops.gemm_mask_prepare(Ws[l]) if self.fused_gemm and dev.type == "cuda" and
                       ops.gemm_mask_supported(Ws[l].shape[0], Ws[l].shape[1]) else None
                       for l in range(1, len(Ws))]
        lanes = min(lanes, len(groups))
        n_split = lay.split_t.n_extra if (pad4 and lay.split_t is not None) else 0
        # two slabs per lane, alternating as SpMM input / output
        row_floats = grp * dmax
        bufs = [(_slab(dev, i, 0, n_in * row_floats),
                 _slab(dev, i, 1, max(n_loc + n_split, 1) * row_floats),
                 _slab(dev, i, 2, max(n_loc, 1) * row_floats) if ragged else None)
                for i in range(lanes)]
        if lanes == 1:
            for c0, gc in groups:
                for _ in self._chain(lay, logits, idx, Hs, Ws, Wp, c0, gc, width_of(gc), bufs[0][0], bufs[0][1], G, hdr,
                                     bufs[0][2]):
                    pass
            return grp, len(groups)
        # two column groups in flight, each on its own stream with its own buffers and factor
        # accumulators: while one waits for its all-gather the other runs its SpMM / SYRK / GEMM
        on_device = dev.type == "cuda"           # the CPU double walks the same interleaving without streams
        main = torch.cuda.current_stream(dev) if on_device else None
        streams = _lane_streams(dev, lanes) if on_device else [None] * lanes
        G_lane = [G] + [[torch.zeros_like(t) for t in G] for _ in range(lanes - 1)]
        for st in streams:
            if st is not None:
                st.wait_stream(main)
        pending = list(groups)
        active = [None] * lanes
        while pending or any(a is not None for a in active):
            for i in range(lanes):
                with (torch.cuda.stream(streams[i]) if on_device else contextlib.nullcontext()):
                    if active[i] is None and pending:
                        c0, gc = pending.pop(0)
                        active[i] = self._chain(lay, logits, idx, Hs, Ws, Wp, c0, gc, width_of(gc), bufs[i][0],
                                                bufs[i][1], G_lane[i], hdrs[i], bufs[i][2])
                    if active[i] is not None:
                        try:
                            next(active[i])
                        except StopIteration:
                            active[i] = None
        for st in streams:
            if st is not None:
                main.wait_stream(st)
        for extra in G_lane[1:]:
            for t, e in zip(G, extra):
                t += e
        return grp, len(groups)

    # ------------------------------------------------------------------ kron
    def kron(self, x: torch.Tensor, y: torch.Tensor, N: int, **kwargs):
        """(loss, Kron) for the batch of train-node indices ``x`` with labels ``y``; ``N`` is the
        size of the whole training set (curvature.py:236-265).  With a ``process_group`` every rank
        passes the same (x, y) and receives the same (all-reduced) result.  Runs on the MODEL's device,
        whichever device is current in the calling thread."""
        from ._lib import on_device_of
        with on_device_of(self.model.X):
            return self._kron(x, y, N, **kwargs)

    def _kron(self, x: torch.Tensor, y: torch.Tensor, N: int, **kwargs):
        model = self.model
        g = model.graph
        Ws, bs = self._layers()
        L = len(Ws)
        M = int(y.shape[0])
        if x.shape[0] != M:
            raise ValueError("x (node indices) and y (labels) must have the same length")
        idx = x.to(torch.int64).contiguous()
        yy = y.to(torch.int64).contiguous()
        part = self._partition()
        Hs, logits = self._forward(Ws, bs, part)
        C = logits.shape[1]
        dev = logits.device
        if part is not None:                                   # this rank's train nodes, local row ids
            mine = (idx >= part.lo) & (idx < part.hi)
            idx_loc, y_loc = (idx[mine] - part.lo).contiguous(), yy[mine].contiguous()
        else:
            idx_loc, y_loc = idx, yy
        loss, _hits = ops.softmax_ce_sum(logits, idx_loc, y_loc)

        # input-side factors A_l over ALL graph nodes (kfac.py:870), then the M/N rescale
        # (curvlinops.py:46-53); with a partition: this rank's rows, summed by the all-reduce below
        A = []
        for l in range(L):
            if l == 0 and self.cache_input_factor:
                key = ("xtx", Hs[0].data_ptr(), tuple(Hs[0].shape), Hs[0]._version)
                if key not in self._cache:
                    self._cache.clear()
                    self._cache[key] = ops.syrk(Hs[0], impl=self._impl(Hs[0].shape[1]))
                a = self._cache[key] * (1.0 / M)
                if part is None:
                    # A_0 = raw / N is the cached Gram matrix times a scalar: the stand-in Kron.decompose can
                    # reuse ONE eigendecomposition of the raw matrix across fits (eigenvalues scale, vectors stay)
                    # — on the Cora shape the 1433 x 1433 eigh is 19 of the 44 ms of a fit
                    a._eig_of = (self._cache, key, (1.0 / M) * (M / N))
            else:
                a = ops.syrk(Hs[l], alpha=1.0 / M, impl=self._impl(Hs[l].shape[1]))
            a *= M / N
            A.append(a)

        # output-side factors G_l: multi-RHS backward in groups of Hessian-sqrt columns
        G = [torch.zeros(w.shape[0], w.shape[0], dtype=torch.float32, device=dev) for w in Ws]
        self._n_unit_spmm = 0
        whole = _Whole(g)
        mx_t = g.ahat_t.max_row_nnz
        if self.unit_slabs and self.unit_hub_split and mx_t is not None and mx_t > self.unit_row_limit:
            from .graph import split_hub_rows
            cache = g.meta.setdefault("_split_t", {})            # per graph, per limit: index work done once
            if self.unit_row_limit not in cache:
                cache[self.unit_row_limit] = split_hub_rows(g.ahat_t, self.unit_row_limit)
            whole.split_t = cache[self.unit_row_limit]
        if self.skip_zero_rows and (part is None or self.backward_parallel == "columns") and M < g.n:
            keep = torch.zeros(g.n, dtype=torch.uint8, device=dev)
            keep[idx] = 1
            whole.csr_t_top = ops.csr_with_masked_sources(g.ahat_t, keep)
        mx_row = g.ahat_t.max_row_nnz
        rows_ok = (part is not None and self.backward_parallel == "rows" and self.rows_hess_stats and
                   self.unit_rows and not part.sparse_halo)
        if (self.fused_hess_spmm and (part is None or self.backward_parallel == "columns" or rows_ok) and
                ops.spmm_hess_supported(C, 1) and mx_row is not None and mx_row <= self.unit_row_limit):
            self._want_hess_stats = True
        else:
            self._want_hess_stats = False
        if part is None:
            if self._want_hess_stats:
                cp = (C + 3) // 4 * 4
                whole.hess_stats = ops.hess_stats(logits, idx, self.hess_sqrt, C,
                                                  out=_slab(dev, 2000, 0, g.n * 5 * cp).view(g.n, 5 * cp))
            grp, n_groups = self._backward_columns(whole, logits, idx, Hs, Ws, (0, C), G)
        elif self.backward_parallel == "rows":
            lay = _Rows(part)
            if self._want_hess_stats:
                # output layer without a slab exchange: every rank writes the five softmax vectors of its train nodes
                # into its slot of a padded [N, 5 Cp] slab, ONE all-gather per fit (2.3 GB on the products shape
                # against 22 GB of dense right-hand sides), and lgnn_spmm_hess_f32 rebuilds the right-hand sides per
                # edge; edges from non-train sources are zeroed in this rank's slice of Â^T as on one device
                cp = (C + 3) // 4 * 4
                stats = _slab(dev, 2000, 0, part.total_rows * 5 * cp).view(part.total_rows, 5 * cp)
                stats.zero_()
                ops.hess_stats(logits, idx_loc, self.hess_sqrt, C, out=stats[part.slot0:part.slot0 + part.n_local])
                with ops.timed("allgather", 5 * cp, 4.0 * part.total_rows * 5 * cp):
                    part.all_gather_slab(stats)
                lay.hess_stats = stats
                if self.skip_zero_rows and M < g.n:
                    # train flags of all nodes in the padded layout; slots padded to 16 bytes for the all-gather
                    pad16 = (part.pad + 15) // 16 * 16
                    flags = torch.zeros(part.world, pad16, dtype=torch.uint8, device=dev)
                    flags[part.rank, idx_loc] = 1
                    part.all_gather_slab(flags)
                    keep = flags[:, :part.pad].reshape(-1)
                    lay.csr_t_top = ops.csr_with_masked_sources(part.ahat_t, keep)
            grp, n_groups = self._backward_columns(lay, logits, idx_loc, Hs, Ws, (0, C), G)
        else:                                                  # "columns": full graph, own columns
            from .dist import column_share
            # every rank needs H_l (relu' masks) and the logits of ALL nodes: in-place all-gather of the padded
            # slabs the forward wrote its rows into, then into natural node order — persistent buffers, no
            # allocation per fit
            # The logits are needed at once; the hidden activations only when the backward reaches their layer, so
            # their all-gathers (2.5 GB into every rank each on the products shape) are issued asynchronously, last
            # layer first, and waited for on first use (_DeferredActivations): they travel under the output-layer
            # SpMM / SYRK / GEMM instead of in front of them.
            def finish(l, w):
                slab = self._fwd_out[l]
                ldz = slab.shape[1]
                return part.compact(slab, ldz, out=_slab(dev, 1100 + l, 2, g.n * ldz).view(g.n, ldz))[:, :w]

            l_top = len(Hs) - 1                                # index of the logits' slab in _fwd_out
            with ops.timed("allgather", C, 4.0 * part.total_rows * self._fwd_out[l_top].shape[1]):
                part.all_gather_slab(self._fwd_out[l_top])
            full_logits = finish(l_top, C)
            full_H = _DeferredActivations(Hs[0], len(Hs))      # slot 0 (the features) is never read by the backward
            for l in range(len(Hs) - 1, 0, -1):                # H_l lives in _fwd_out[l - 1]
                w = Hs[l].shape[1]
                work = part.all_gather_slab(self._fwd_out[l - 1], async_op=dev.type == "cuda" and self.defer_gathers)

                def ready(l=l, w=w, work=work):
                    with ops.timed("allgather", w, 4.0 * part.total_rows * self._fwd_out[l - 1].shape[1]):
                        if work is not None:
                            work.wait()
                    return finish(l - 1, w)
                full_H.defer(l, ready)
            if self._want_hess_stats:
                cp = (C + 3) // 4 * 4
                whole.hess_stats = ops.hess_stats(full_logits, idx, self.hess_sqrt, C,
                                                  out=_slab(dev, 2000, 0, g.n * 5 * cp).view(g.n, 5 * cp))
            grp, n_groups = self._backward_columns(whole, full_logits, idx, full_H, Ws,
                                                   column_share(C, part.rank, part.world), G)
        if part is not None:
            with ops.timed("allreduce", 0, 4.0 * sum(t.numel() for t in G + A)):
                part.all_reduce_sum(G + A)
                part.all_reduce_sum([loss])
        self.last_stats = {"group": grp, "n_groups": n_groups, "M": M, "C": C,
                           "world": 1 if part is None else part.world,
                           "partition": part, "unit_slabs": self._n_unit_spmm}

        Kron = _kron_class()
        kfacs = []
        for l in range(L):
            kfacs.append([G[l], A[l]])
            if bs[l] is not None:
                dup = G[l].clone()
                dup._dup_of = G[l]          # lets the stand-in Kron.decompose skip the repeated eigh
                kfacs.append([dup])
        kron = Kron(kfacs)
        return (self.factor * loss).to(torch.float32), kron

    def _impl(self, n: int) -> str:
        return self.syrk_impl

    # ------------------------------------------------------------------ diag
    def diag(self, x: torch.Tensor, y: torch.Tensor, N: int | None = None, **kwargs):
        """(loss, diag GGN [P]) with the true Λ = diag(p) - pp^T (curvature.py:365-372, 412-432).

        ``diag_mode="exact"`` (default): exact, and therefore — like the reference, which materialises an
        (M, C, P) Jacobian — only for small graphs: one back-propagated column per (train node, class)
        pair, processed as multi-RHS tiles through the same SpMM kernel.
        ``diag_mode="node_factorised"``: an APPROXIMATION for graphs where that is infeasible:
        sum_c (gZ_c ∘ gZ_c)^T (H ∘ H), one multi-RHS backward like ``kron``; exact when no edge couples
        two nodes."""
        from ._lib import on_device_of
        from .diag import diag_ggn_exact, diag_ggn_node_factorised
        with on_device_of(self.model.X):
            if self.diag_mode == "node_factorised":
                return diag_ggn_node_factorised(self, x, y)
            return diag_ggn_exact(self, x, y)


def make_backend(base: type, name: str = "B200GGN") -> type:
    """Build the backend class on top of a given ``CurvatureInterface`` base (the reference's
    ``laplace.curvature.GGNInterface`` when integrating into that package)."""

    def __init__(self, model, likelihood, last_layer=False, subnetwork_indices=None,
                 dict_key_x="input_ids", dict_key_y="labels", stochastic=False, **b200_options):
        """``b200_options``: the keyword switches of ``_B200KFAC._b200_setup`` (hess_sqrt, process_group,
        backward_parallel, rhs_tile_bytes, ...); an unknown one raises TypeError there."""
        if stochastic:
            raise NotImplementedError("the MC Fisher is outside the hot path (TYPE2 GGN only)")
        try:
            base.__init__(self, model, likelihood, last_layer, subnetwork_indices, dict_key_x, dict_key_y)
        except TypeError:  # GGNInterface takes a `stochastic` argument in some versions
            base.__init__(self, model, likelihood, last_layer, subnetwork_indices, dict_key_x,
                          dict_key_y, stochastic)
        self.stochastic = False
        self._b200_setup(**b200_options)

    return type(name, (_B200KFAC, base), {"__init__": __init__, "__doc__": __doc__})


def _default_base() -> type:
    mod = sys.modules.get("laplace.curvature")
    if mod is not None and hasattr(mod, "GGNInterface"):
        return mod.GGNInterface
    return CurvatureInterfaceLite


B200GGN = make_backend(_default_base())
