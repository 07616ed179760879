"""Sparse GCN mirroring the reference's model interface (gnn/models/layers.py:32-46 GCNConv,
gnn/models/base_gnn.py BaseGNN, gnn/models/models.py:14-34 GCN) with the dense ``adj @ x``
replaced by the hand-written CSR SpMM.

Same names as the reference so its call sites read the same:
``model.convs[l].lin`` is the ``nn.Linear`` whose weight/bias the Laplace approximation is
taken over, ``model(idx)`` returns the logits of the nodes ``idx``, parameters are named
``convs.{l}.lin.weight / bias`` (the order laplace's parameter vector uses).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from ._lib import on_device_of
from .graph import Graph


class GCNConvFunction(torch.autograd.Function):
    """y = Â z.  backward: dz = Â^T dy (what autograd derives for the reference's dense
    ``adj @ z``, layers.py:46); the graph gets no gradient (adjacency fixed: update_adj=False)."""

    @staticmethod
    def forward(ctx, z: torch.Tensor, graph: Graph) -> torch.Tensor:
        ctx.graph = graph
        with on_device_of(z):
            return graph.propagate(z.contiguous())

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        with on_device_of(grad_out):
            return ctx.graph.propagate(grad_out.contiguous(), transpose=True), None


class SparseGCNConv(nn.Module):
    """Drop-in for the reference GCNConv: ``forward(graph, x) = Â @ lin(x)`` (bias before
    aggregation, layers.py:45-46)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.lin = nn.Linear(in_channels, out_channels, bias=bias)

    def reset_parameters(self):
        self.lin.reset_parameters()

    def forward(self, graph: Graph, x: torch.Tensor) -> torch.Tensor:
        return GCNConvFunction.apply(self.lin(x), graph)


class SparseGCN(nn.Module):
    """Mirror of ``GCN(in_channels, hidden_channels, out_channels, num_layers, X, init_adj, ...)``
    with ``init_adj`` replaced by a ``Graph`` (or an int64 edge_index [2, E], from which the graph
    is built with self loops exactly like ``GCN.__init__`` + ``normalize_adj``).

    Supported on the hot path: act="relu", norm=None, res=False (the GCN configuration of
    BASELINE.json); dropout is active in train mode only, like the reference.
    """

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int,
                 X: torch.Tensor, graph, dropout_p: float = 0.5, act: Optional[str] = "relu",
                 symmetric: bool = False, bias: bool = True, x_rows: Optional[tuple] = None, **kwargs):
        super().__init__()
        if act != "relu":
            raise NotImplementedError("SparseGCN supports act='relu' only (the GCN hot path)")
        if kwargs.get("norm") not in (None, "none") or kwargs.get("res", False):
            raise NotImplementedError("norm / res are outside the GCN hot path")
        if not isinstance(graph, Graph):
            graph = Graph.from_edge_index(graph, X.shape[0], symmetric=symmetric)
        self.graph = graph
        self.X = X
        # multi-GPU Laplace fits read only this rank's node block of X (dist.row_block): ``x_rows=(lo, hi)`` says
        # that X holds just those rows (dist.ingest_rows).  Such a model serves B200GGN with a process group; its
        # own forward() (training step) needs all rows and raises.
        self.x_rows = None if x_rows is None else (int(x_rows[0]), int(x_rows[1]))
        if self.x_rows is not None and X.shape[0] != self.x_rows[1] - self.x_rows[0]:
            raise ValueError(f"x_rows={self.x_rows} but X has {X.shape[0]} rows")
        self.in_channels = in_channels
        self.hidden_channels = hidden_channels
        self.out_channels = out_channels
        self.num_layers = num_layers
        self.dropout = nn.Dropout(p=dropout_p)
        self.act = nn.ReLU()
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.convs = nn.ModuleList(
            [SparseGCNConv(dims[i], dims[i + 1], bias=bias) for i in range(num_layers)])

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, x_indices: torch.Tensor) -> torch.Tensor:
        if self.x_rows is not None:
            raise RuntimeError("this SparseGCN holds a row block of X (x_rows): only the row-partitioned Laplace fit "
                               "(B200GGN with a process group) can run on it")
        x = self.X
        for i in range(self.num_layers - 1):
            x = self.convs[i](self.graph, x)
            x = self.act(x)
            x = self.dropout(x)
        x = self.convs[-1](self.graph, x)
        return x[x_indices]
