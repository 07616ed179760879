"""Row-partitioned multi-GPU execution of the GCN KFAC-GGN pass (one process per GPU,
``torch.distributed`` / NCCL over NVLink).  The reference is single-device; this is
the B200 design of BASELINE.json's north star:

* the graph is split into ``world`` contiguous row blocks balanced by nnz (``lgnn_row_partition``,
  bit-exact against ``oracle.row_partition``); rank r owns rows [bounds[r], bounds[r+1]) of every
  activation / gradient slab and the matching CSR row slices of Â and Â^T;
* before each local SpMM the ranks all-gather the slab the SpMM reads (the halo of a uniformly
  random graph with average degree >= 15 is > 85 % of all rows, so the whole slab is exchanged;
  ``halo_fraction`` reports the exact figure).  Slabs live in a padded layout
  [world, pad, width] so the all-gather is in place and the local CSR slices carry column indices
  remapped to that layout (``lgnn_csr_slice_remap``);
* the small d x d Kronecker factors are summed with ONE all-reduce at the end of the pass, the loss
  with a second (double precision); the factor eigendecompositions / marglik then run replicated.

``backward_parallel`` selects how the C Hessian-sqrt columns of the backward are spread:

  "rows"     every rank processes all C columns on its row block; each layer step all-gathers the
             group's right-hand sides — below the output layer as ragged unit-compacted rows (the
             live relu units of each node, back to back: about half the N x (g*d) slab), at the output
             layer not at all (softmax statistics gathered once per fit) — hidden behind the SpMM of
             the other in-flight column group when ``overlap`` is on;
  "columns"  the forward stays row-partitioned (halo all-gather of Z_l, all-gather of H_l), then
             rank r back-propagates its own C/world columns on the full graph: no data-path
             collective in the backward at all.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import ops
from .ops import CSR


# Collectives of one communicator must not run concurrently.  torch launches a blocking collective (async_op=False)
# on the CALLER's current stream, so the two lanes of the overlapped rows layout — each on its own stream — could put
# two all-gathers of the same communicator on the device at once (seen once as a 600 s watchdog stall of the NCCL
# parity test on small shapes, where the lanes' collectives are microseconds apart).  A blocking collective issued
# here on ANOTHER stream than the previous one therefore first makes its stream wait for an event recorded behind
# that one: the host order, which is the same on every rank, becomes the device order.  Collectives that follow each
# other on one stream are ordered anyway and are left alone; asynchronous ones (the deferred all-gathers of the
# column-parallel backward) run on the group's own stream, and their callers wait for them before the next blocking
# collective.
_LAST_COLLECTIVE: dict = {}      # id(process group) -> (stream id, event recorded behind the last blocking collective)


def _after_previous(pg, t: torch.Tensor, async_op: bool = False) -> None:
    if not t.is_cuda or async_op:
        return
    prev = _LAST_COLLECTIVE.get(id(pg))
    cur = torch.cuda.current_stream(t.device)
    if prev is not None and prev[0] != cur.cuda_stream:
        cur.wait_event(prev[1])


def _issued(pg, t: torch.Tensor, work=None) -> None:
    if not t.is_cuda or work is not None:
        return
    cur = torch.cuda.current_stream(t.device)
    ev = torch.cuda.Event()
    ev.record(cur)
    _LAST_COLLECTIVE[id(pg)] = (cur.cuda_stream, ev)


@dataclass
class RowPartition:
    """State of one rank for a given (graph, process group)."""
    pg: object
    rank: int
    world: int
    bounds: list            # world+1 python ints
    pad: int                # rows per slot of the padded all-gather layout
    ahat: CSR               # rows [lo, hi) of Â,   columns in padded layout
    ahat_t: CSR             # rows [lo, hi) of Â^T, columns in padded layout
    graph: object = None    # the full graph (for the lazily evaluated halo statistic)
    _halo_fraction: float | None = None
    _halo_plan: dict | None = None
    # below this fraction of the other ranks' rows the SpMM inputs travel as halo rows (all-to-all) instead of
    # whole slabs (all-gather); a uniformly random graph sits at > 0.99, a partitioned mesh-like graph far below
    # (0 = never: the switch of B200GGN(sparse_halo=True) sets 0.5)
    sparse_halo_below: float = 0.0

    @property
    def halo_fraction(self) -> float:
        """Fraction of the other ranks' rows this rank's backward SpMM reads (lgnn_halo_mark; one host
        sync, evaluated on first use only)."""
        if self._halo_fraction is None:
            n_other = self.graph.n - (self.hi - self.lo)
            halo = ops.halo_columns(self.graph.ahat_t, self.lo, self.hi)
            self._halo_fraction = float(halo.numel()) / n_other if n_other > 0 else 0.0
        return self._halo_fraction

    @property
    def lo(self) -> int:
        return self.bounds[self.rank]

    @property
    def hi(self) -> int:
        return self.bounds[self.rank + 1]

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    @property
    def total_rows(self) -> int:
        return self.world * self.pad

    @property
    def slot0(self) -> int:
        return self.rank * self.pad

    @classmethod
    def build(cls, graph, pg) -> "RowPartition":
        rank, world = dist.get_rank(pg), dist.get_world_size(pg)
        bounds_t = graph.partition_bounds(world)
        bounds = [int(v) for v in bounds_t.tolist()]
        pad = max(max(bounds[r + 1] - bounds[r] for r in range(world)), 1)
        lo, hi = bounds[rank], bounds[rank + 1]
        a = ops.csr_slice_remap(graph.ahat, lo, hi, bounds_t, pad)
        at = a if graph.ahat_t is graph.ahat else ops.csr_slice_remap(graph.ahat_t, lo, hi, bounds_t, pad)
        return cls(pg, rank, world, bounds, pad, a, at, graph)

    # ---- collectives -------------------------------------------------------------------
    def all_gather_slab(self, slab: torch.Tensor, async_op: bool = False):
        """In-place all-gather of a padded slab [world*pad, width]: every rank has filled its own
        slot rows [rank*pad, rank*pad + n_local)."""
        flat = slab.view(self.world, -1)
        mine = flat[self.rank]
        if not slab.is_cuda:           # gloo (CPU tests): no in-place guarantee
            mine = mine.clone()
        _after_previous(self.pg, slab, async_op)
        work = dist.all_gather_into_tensor(flat.view(-1), mine.reshape(-1), group=self.pg, async_op=async_op)
        _issued(self.pg, slab, work if async_op else None)
        return work

    # ---- halo-only exchange ------------------------------------------------------------
    @property
    def sparse_halo(self) -> bool:
        return self.world > 1 and self.sparse_halo_below > 0.0 and self.halo_fraction < self.sparse_halo_below

    def halo_plan(self) -> dict:
        """Who sends which rows to whom, built once per (graph, group): this rank's halo = the columns its row
        slices of Â and Â^T reference outside [lo, hi) (lgnn_halo_mark), grouped by owner; the lists are
        exchanged so that every rank knows which of its own rows each peer needs."""
        if self._halo_plan is not None:
            return self._halo_plan
        g, dev = self.graph, self.ahat.rowptr.device
        need = ops.halo_columns(g.ahat, self.lo, self.hi)
        if g.ahat_t is not g.ahat:
            need = torch.unique(torch.cat([need, ops.halo_columns(g.ahat_t, self.lo, self.hi)]))
        need = need.to(torch.int64)
        bounds = torch.tensor(self.bounds, dtype=torch.int64, device=dev)
        owner = torch.searchsorted(bounds, need, right=True) - 1
        recv_counts = torch.bincount(owner, minlength=self.world)
        send_counts = torch.empty_like(recv_counts)
        dist.all_to_all_single(send_counts, recv_counts, group=self.pg)
        rc, sc = recv_counts.tolist(), send_counts.tolist()
        wanted = torch.empty(int(sum(sc)), dtype=torch.int64, device=dev)      # global ids peers want from me
        dist.all_to_all_single(wanted, need.contiguous(), output_split_sizes=sc, input_split_sizes=rc, group=self.pg)
        self._halo_plan = {
            "send_rows": (wanted - self.lo + self.slot0).contiguous(),          # my rows, positions in the padded slab
            "recv_rows": (owner * self.pad + (need - bounds[owner])).contiguous(),
            "send_counts": sc, "recv_counts": rc}
        return self._halo_plan

    def exchange_for_spmm(self, slab: torch.Tensor):
        """Make the rows this rank's SpMM reads present in the padded slab: the whole slab (all-gather) when the
        halo is dense, only the halo rows (one all-to-all of row lists fixed at partition time) when it is
        sparse.  Rows nobody reads stay whatever they were."""
        if not self.sparse_halo:
            return self.all_gather_slab(slab)
        plan = self.halo_plan()
        width = slab.shape[1]
        send = slab.index_select(0, plan["send_rows"])
        recv = torch.empty(int(sum(plan["recv_counts"])), width, dtype=slab.dtype, device=slab.device)
        _after_previous(self.pg, slab)
        dist.all_to_all_single(recv.view(-1), send.view(-1),
                               output_split_sizes=[c * width for c in plan["recv_counts"]],
                               input_split_sizes=[c * width for c in plan["send_counts"]], group=self.pg)
        _issued(self.pg, slab)
        slab.index_copy_(0, plan["recv_rows"], recv)
        return None

    def compact(self, slab: torch.Tensor, width: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """Padded [world*pad, width] -> natural node order [N, width] (into ``out`` when given)."""
        v = slab.view(self.world, self.pad, width)
        pieces = [v[r, : self.bounds[r + 1] - self.bounds[r]] for r in range(self.world)]
        return torch.cat(pieces, dim=0) if out is None else torch.cat(pieces, dim=0, out=out)

    def all_reduce_sum(self, tensors) -> None:
        """One all-reduce over the concatenation of same-dtype tensors (written back in place)."""
        flat = torch.cat([t.reshape(-1) for t in tensors])
        _after_previous(self.pg, flat)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
        _issued(self.pg, flat)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n


def column_share(C: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of Hessian-sqrt columns owned by ``rank``: sizes differ by at most one."""
    base, rem = divmod(C, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


# ------------------------------------------------------------------------------------------------
# sharded ingest: each rank moves 1 / world of the host inputs over ITS PCIe link
# ------------------------------------------------------------------------------------------------
def ingest_edge_index(host_edge_index: torch.Tensor, device, pg) -> torch.Tensor:
    """Host edge list [2, E] (int64, ideally pinned; identical on every rank) -> the same list on every rank's
    device, with each rank copying only its E / world columns over PCIe and the ranks exchanging the shards with
    ONE in-place all-gather over NVLink.  (Eight ranks each pulling the whole 2 GB list of the products shape
    through the host's memory system at once took 260 ms per fit in round 1; 1 / 8 each + 2 GB over NVSwitch is a
    few ms.)  Shards are padded to a common length by repeating their first edge: the CSR build collapses
    duplicates, so the graph is the one the unsharded list gives — bit for bit."""
    world, rank = dist.get_world_size(pg), dist.get_rank(pg)
    if host_edge_index.dtype != torch.int64 or host_edge_index.dim() != 2 or host_edge_index.shape[0] != 2:
        raise TypeError("edge_index must be an int64 tensor of shape [2, E]")
    E = int(host_edge_index.shape[1])
    if E == 0:
        return torch.empty(2, 0, dtype=torch.int64, device=device)
    S = (E + world - 1) // world
    lo, hi = min(rank * S, E), min((rank + 1) * S, E)
    buf = torch.empty(world, 2, S, dtype=torch.int64, device=device)
    mine = buf[rank]
    if hi > lo:
        for r in range(2):       # row by row: a [2, cols] slice of the pinned list is strided and would be staged on the host
            mine[r, : hi - lo].copy_(host_edge_index[r, lo:hi], non_blocking=True)
        if hi - lo < S:
            mine[:, hi - lo:] = mine[:, :1]
    else:                                    # more ranks than edges: a copy of the list's first edge
        mine[:] = host_edge_index[:, :1].to(device)
    flat = buf.view(-1)
    src = mine.reshape(-1)
    if device.type != "cuda":                # gloo (CPU tests): no in-place guarantee
        src = src.clone()
    dist.all_gather_into_tensor(flat, src, group=pg)
    return buf.permute(1, 0, 2).reshape(2, world * S)


def ingest_rows(host_x: torch.Tensor, lo: int, hi: int, device) -> torch.Tensor:
    """Rows [lo, hi) of a host matrix on the device (this rank's block of the node features: the row-partitioned
    forward reads nothing else of X).  Pass the result to ``SparseGCN(..., x_rows=(lo, hi))``."""
    return host_x[lo:hi].to(device, non_blocking=True)


def row_block(graph, pg) -> tuple[int, int]:
    """This rank's contiguous node block [lo, hi) of the nnz-balanced row partition (the one RowPartition.build
    derives for the same graph and group)."""
    b = graph.partition_bounds(dist.get_world_size(pg)).tolist()
    r = dist.get_rank(pg)
    return int(b[r]), int(b[r + 1])
