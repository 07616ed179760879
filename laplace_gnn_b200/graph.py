"""Sparse normalised adjacency Â = D^-1/2 A^T D^-1/2 (D = row sums of A incl. self loops) on the
device, replacing the reference's dense N x N ``adj`` parameter and its per-forward
``normalize_adj`` (gnn/models/utils.py:106-112, gnn/models/base_gnn.py:137).

``Graph`` is built once (integer kernels, bit-exact against the oracle) and then shared by the
forward (CSR of Â), the backward / KFAC passes (CSR of Â^T = pattern of A) and the row
partition used across GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import ops
from .ops import CSR


@dataclass
class Graph:
    n: int
    deg: torch.Tensor        # int64 [n] out-degree incl. self loop (row sums of A)
    dis: torch.Tensor        # fp32  [n] deg^-1/2
    ahat: CSR                # Â   : row i = in-neighbours of i
    ahat_t: CSR              # Â^T : row i = out-neighbours of i (pattern of A)
    symmetric_pattern: bool = False
    meta: dict = field(default_factory=dict)

    @property
    def nnz(self) -> int:
        return self.ahat.nnz

    @property
    def device(self):
        return self.dis.device

    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, num_nodes: int, symmetric: bool = False,
                        assume_undirected: bool = False) -> "Graph":
        """edge_index: int64 [2, E] on a CUDA device, A[src, dst] = 1 (gnn/utils.py:325-330).

        symmetric=True mirrors BaseGNN(symmetric=True) (A + A^T clamped, base_gnn.py:68-73).
        assume_undirected=True declares that the edge list already contains both directions, so
        Â^T = Â and one CSR serves both (halves graph memory); it is not verified.
        """
        from ._lib import on_device_of
        with on_device_of(edge_index):
            return cls._build(edge_index, num_nodes, symmetric, assume_undirected)

    @classmethod
    def _build(cls, edge_index, num_nodes, symmetric, assume_undirected) -> "Graph":
        a = ops.csr_from_edge_index(edge_index, num_nodes, symmetric)      # pattern of A  == Â^T
        deg, dis = ops.degree_norm(a)
        a.val = ops.edge_values(a, dis)
        if symmetric or assume_undirected:
            at = a
        else:
            at = ops.csr_transpose(a)                                         # pattern of A^T == Â
            at.val = ops.edge_values(at, dis)
        # longest rows (one host read at graph build): SpMM skips its hub-row passes on graphs without hubs
        a.max_row_nnz = int(deg.max()) if num_nodes > 0 else 0
        if at is not a:
            at.max_row_nnz = int((at.rowptr[1:] - at.rowptr[:-1]).max()) if num_nodes > 0 else 0
        return cls(int(num_nodes), deg, dis, at, a, symmetric or assume_undirected,
                   {"symmetrised": bool(symmetric)})

    # ---- Â x / Â^T x with hub rows of power-law graphs cut into pieces -------------------------------------
    HUB_ROW_LIMIT = 4096

    def hub_split(self, transpose: bool = False):
        """``SplitCSR`` of Â (or Â^T) when a row exceeds HUB_ROW_LIMIT non-zeros, else None; built once per graph."""
        csr = self.ahat_t if transpose else self.ahat
        if csr.max_row_nnz is None or csr.max_row_nnz <= self.HUB_ROW_LIMIT:
            return None
        cache = self.meta.setdefault("_hub_split", {})
        key = "t" if (transpose and self.ahat_t is not self.ahat) else "n"
        if key not in cache:
            cache[key] = split_hub_rows(csr, self.HUB_ROW_LIMIT)
        return cache[key]

    def extra_rows(self, transpose: bool = False) -> int:
        """Rows to add to an output buffer handed to ``propagate`` (the pieces of the hub rows)."""
        sp = self.hub_split(transpose)
        return 0 if sp is None else sp.n_extra

    def propagate(self, x: torch.Tensor, transpose: bool = False, relu: bool = False,
                  out: torch.Tensor | None = None) -> torch.Tensor:
        """Â x (or Â^T x), optionally relu'd: the narrow SpMM of the forward / training step.  The warp-per-row
        kernel gives one warp a whole row; on a power-law graph a hub row of 10^5 non-zeros then runs for the whole
        launch (0.51 of the copy rate at d = 256 on R-MAT 2^22, round 1).  Hub rows are therefore cut into pieces of
        HUB_ROW_LIMIT non-zeros (``split_hub_rows``: extra output rows behind the matrix, summed onto their owners
        afterwards).  ``out``: optional buffer with ``n + extra_rows()`` rows; the result is its first n rows."""
        csr = self.ahat_t if transpose else self.ahat
        sp = self.hub_split(transpose)
        if sp is None:
            return ops.spmm(csr, x, relu=relu, out=out)
        d = int(x.shape[1])
        rows = self.n + sp.n_extra
        if out is None or out.shape[0] < rows:
            out = torch.empty(rows, d, dtype=torch.float32, device=x.device)
        ops.spmm(sp.csr, x, out=out)
        y = sp.finish(out, d)
        if relu:
            torch.relu_(y[:, :d])
        return y

    def partition_bounds(self, nparts: int) -> torch.Tensor:
        """nnz-balanced contiguous row blocks of Â (int64 [nparts+1], device)."""
        return ops.row_partition(self.ahat.rowptr, nparts)


# ------------------------------------------------------------------------------------------------
# hub rows of power-law graphs, for the kernels that give one warp group a whole row
# ------------------------------------------------------------------------------------------------
@dataclass
class SplitCSR:
    """A CSR whose rows longer than ``limit`` non-zeros are cut into pieces of at most ``limit``: the first
    piece stays in the row, the others become extra rows appended behind the matrix (rows n_rows ..
    n_rows + n_extra - 1, hub by hub, piece by piece).  ``finish`` adds the extra rows of an SpMM result
    onto their owners.  (col, val) of the extra pieces are moved to the end; everything else keeps its place
    and its order."""
    csr: CSR                 # [n_rows + n_extra, n_cols]
    n_rows: int              # rows of the original matrix
    n_extra: int
    hub_rows: torch.Tensor   # int64 [n_hub] rows that were cut, ascending
    gather: CSR              # [n_hub, n_extra] ones: row k lists the extra rows of hub k, in piece order

    def finish(self, y: torch.Tensor, d: int) -> torch.Tensor:
        """y: [>= n_rows + n_extra, ld] holding the SpMM over ``csr``; returns y[:n_rows] with the pieces summed
        (each hub's extra rows in piece order, then onto the first piece)."""
        extra = y[self.n_rows:self.n_rows + self.n_extra]
        sums = ops.spmm(self.gather, extra, d=d)
        y[self.hub_rows, :d] = y[self.hub_rows, :d] + sums
        return y[:self.n_rows]


def split_hub_rows(a: CSR, limit: int = 4096) -> SplitCSR | None:
    """``SplitCSR`` of ``a``, or None when no row is longer than ``limit``.  Index arithmetic only (torch, on
    the matrix's device), done once per graph."""
    if limit < 1:
        raise ValueError("split_hub_rows: limit must be positive")
    rp = a.rowptr
    lens = rp[1:] - rp[:-1]
    extra_per_row = torch.clamp((lens + (limit - 1)) // limit - 1, min=0)
    hub_rows = torch.nonzero(extra_per_row > 0).reshape(-1)
    if hub_rows.numel() == 0:
        return None
    dev = rp.device
    n, nnz = a.n_rows, a.nnz
    n_extra = int(extra_per_row.sum())
    # entries past the first `limit` of a hub row move behind the matrix; only the hub rows are visited
    moved = torch.zeros(nnz, dtype=torch.bool, device=dev)
    new_lens = lens.clone()
    extra_lens = []
    begins, ends = rp[hub_rows].tolist(), rp[hub_rows + 1].tolist()       # one host read for all hubs
    new_lens[hub_rows] = limit
    for b, e in zip(begins, ends):                               # the hubs only, not the graph
        moved[b + limit:e] = True
        rest = e - b - limit
        extra_lens += [limit] * (rest // limit) + ([rest % limit] if rest % limit else [])
    all_lens = torch.cat([new_lens, torch.tensor(extra_lens, dtype=torch.int64, device=dev)])
    rowptr = torch.zeros(n + n_extra + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(all_lens, 0)
    col = torch.cat([a.col[~moved], a.col[moved]])
    val = None if a.val is None else torch.cat([a.val[~moved], a.val[moved]])
    csr = CSR(n + n_extra, a.n_cols, rowptr, col, val, min(limit, int(lens.max())))
    g_rowptr = torch.zeros(hub_rows.numel() + 1, dtype=torch.int64, device=dev)
    g_rowptr[1:] = torch.cumsum(extra_per_row[hub_rows], 0)
    gather = CSR(int(hub_rows.numel()), n_extra, g_rowptr, torch.arange(n_extra, dtype=torch.int32, device=dev),
                 torch.ones(n_extra, dtype=torch.float32, device=dev), int(extra_per_row.max()))
    return SplitCSR(csr, n, n_extra, hub_rows, gather)


# ------------------------------------------------------------------------------------------------
# graph ingest: on-disk cache of the built CSR, k-nearest-neighbour graphs
# ------------------------------------------------------------------------------------------------


def save_graph(graph: Graph, path: str) -> None:
    """Cache a built graph (integer CSR, degrees, values) so that it is not rebuilt per run."""
    def pack(c: CSR):
        return {"n_rows": c.n_rows, "n_cols": c.n_cols, "rowptr": c.rowptr.cpu(), "col": c.col.cpu(),
                "val": None if c.val is None else c.val.cpu(), "max_row_nnz": c.max_row_nnz}
    same = graph.ahat_t is graph.ahat
    torch.save({"n": graph.n, "deg": graph.deg.cpu(), "dis": graph.dis.cpu(), "ahat": pack(graph.ahat),
                "ahat_t": None if same else pack(graph.ahat_t), "symmetric_pattern": graph.symmetric_pattern,
                "meta": {k: v for k, v in graph.meta.items() if not k.startswith("_")}}, path)


def load_graph(path: str, device) -> Graph:
    # the payload save_graph writes is dicts / tensors / ints / bools / None only: no pickled code is accepted
    z = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(z, dict) or not {"n", "deg", "dis", "ahat", "ahat_t", "symmetric_pattern", "meta"} <= set(z):
        raise ValueError(f"{path}: not a graph cache written by save_graph")
    for name, dt in (("deg", torch.int64), ("dis", torch.float32)):
        if not isinstance(z[name], torch.Tensor) or z[name].dtype != dt or z[name].numel() != z["n"]:
            raise ValueError(f"{path}: field {name!r} has the wrong type or length")
    for name in ("ahat", "ahat_t"):
        d = z[name]
        if d is None and name == "ahat_t":
            continue
        if (not isinstance(d, dict) or d["rowptr"].dtype != torch.int64 or d["col"].dtype != torch.int32 or
                d["rowptr"].numel() != d["n_rows"] + 1 or int(d["rowptr"][-1]) != d["col"].numel() or
                (d["val"] is not None and (d["val"].dtype != torch.float32 or d["val"].numel() != d["col"].numel()))):
            raise ValueError(f"{path}: field {name!r} is not a consistent CSR")

    def unpack(d):
        return CSR(d["n_rows"], d["n_cols"], d["rowptr"].to(device), d["col"].to(device),
                   None if d["val"] is None else d["val"].to(device), d["max_row_nnz"])
    a = unpack(z["ahat"])
    at = a if z["ahat_t"] is None else unpack(z["ahat_t"])
    return Graph(z["n"], z["deg"].to(device), z["dis"].to(device), a, at, z["symmetric_pattern"], z["meta"])


def knn_edge_index(X: torch.Tensor, k: int = 3, chunk: int = 4096) -> torch.Tensor:
    """Directed k-nearest-neighbour edges (neighbour -> node, Euclidean, no self loops) like the
    ``knn_graph(X, k, loop=False, cosine=False)`` call inside ``get_knn_graph`` (gnn/utils.py:355-369).
    Build the reference's kNN graph (symmetrised, self loops set) with
    ``Graph.from_edge_index(knn_edge_index(X, k), n, symmetric=True)``.  Distances are evaluated in
    row chunks (a library GEMM + top-k per chunk), never as an N x N matrix."""
    n = X.shape[0]
    if not 1 <= k < n:
        raise ValueError("knn_edge_index needs 1 <= k < number of nodes")
    Xf = X.float()
    sq = (Xf * Xf).sum(1)
    src, dst = [], []
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        d2 = sq[s:e, None] + sq[None, :] - 2.0 * (Xf[s:e] @ Xf.t())
        d2[torch.arange(e - s, device=X.device), torch.arange(s, e, device=X.device)] = float("inf")   # loop=False
        nb = torch.topk(d2, k, dim=1, largest=False).indices                     # [chunk, k]
        src.append(nb.reshape(-1))
        dst.append(torch.arange(s, e, device=X.device).repeat_interleave(k))
    return torch.stack([torch.cat(src), torch.cat(dst)]).to(torch.int64)
