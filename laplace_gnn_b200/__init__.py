"""laplace_gnn_b200 — B200-native GCN forward/backward + KFAC-GGN Kronecker-factor accumulation
behind the reference's curvature-backend plugin API (anitasyang/Laplace-GNN hot path).

Public surface:
    Graph, SparseGCN, SparseGCNConv, GCNConvFunction   model side (mirrors gnn/models)
    B200GGN, make_backend                              curvature backend (laplace/curvature contract)
    Laplace, KronLaplace, DiagLaplace, Kron            stand-ins when the `laplace` package is absent
    marglik_edge_grad, log_marginal_likelihood_of_edges  d marglik / dA on sparse entries (structure learning)
    ops                                                tensor-level wrappers over the C-ABI (include/lgnn.h)
"""
from . import _lib, ops  # noqa: F401
from .graph import Graph, knn_edge_index, load_graph, save_graph  # noqa: F401
from .gcn import GCNConvFunction, SparseGCN, SparseGCNConv  # noqa: F401
from .curvature import B200GGN, make_backend, release_workspace, workspace_bytes  # noqa: F401
from .data import TensorBatchLoader  # noqa: F401
from . import datasets  # noqa: F401
from .kron import DiagLaplace, Kron, KronDecomposed, KronLaplace, Laplace  # noqa: F401
from .training import MarglikTrainingResult, marglik_training  # noqa: F401
from .structure import EdgeGradient, EdgeScores, log_marginal_likelihood_of_edges, marglik_edge_grad  # noqa: F401

__all__ = ["Graph", "knn_edge_index", "save_graph", "load_graph", "SparseGCN", "SparseGCNConv", "GCNConvFunction", "B200GGN", "make_backend", "release_workspace", "workspace_bytes",
           "TensorBatchLoader", "datasets", "marglik_edge_grad", "log_marginal_likelihood_of_edges", "EdgeGradient", "EdgeScores", "marglik_training", "MarglikTrainingResult", "Laplace", "KronLaplace", "DiagLaplace", "Kron", "KronDecomposed", "ops"]
