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
        return cls(int(num_nodes), deg, dis, at, a, symmetric or assume_undirected)

    def partition_bounds(self, nparts: int) -> torch.Tensor:
        """nnz-balanced contiguous row blocks of Â (int64 [nparts+1], device)."""
        return ops.row_partition(self.ahat.rowptr, nparts)
