"""TEST INFRASTRUCTURE — dev container only.  Golden of the reference's kNN graph recipe.

Runs the reference's own `gnn.utils.get_knn_graph(X, k, return_edge_index=True)` (gnn/utils.py:355-369: kNN edges ->
`edge_index_to_adj` -> `adj + adj.T` -> `fill_diagonal_(1)` -> `adj_to_edge_index`).  Its first step,
`torch_geometric.nn.knn_graph` = `torch_cluster.knn_graph` (third party, not vendored, not installed: pyproject lists
torch_geometric without a pin), is supplied here by its published definition in float64 — for every node the k
nearest other nodes in Euclidean distance, edges neighbour -> node (`flow="source_to_target"`), `loop=False`,
`cosine=False` — and `torch_geometric.utils.to_scipy_sparse_matrix` likewise; everything else is the reference's code.  Features are drawn so that no two distances tie.
    python oracle/make_golden_knn.py   ->  tests/golden/knn_*.npz"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def knn_graph_published(x, k, batch=None, loop=False, flow="source_to_target", cosine=False, num_workers=1):
    assert batch is None and not loop and not cosine and flow == "source_to_target"
    xd = x.double()
    d = ((xd[:, None, :] - xd[None, :, :]) ** 2).sum(-1)
    d.fill_diagonal_(float("inf"))
    nb = torch.topk(d, k, dim=1, largest=False).indices
    n = x.shape[0]
    return torch.stack([nb.reshape(-1), torch.arange(n).repeat_interleave(k)])


def to_scipy_sparse_matrix_published(edge_index, edge_attr=None, num_nodes=None):
    """torch_geometric.utils.to_scipy_sparse_matrix (third party): COO with unit (or given) weights, duplicates
    summed by scipy on conversion."""
    import scipy.sparse as sp
    r, c = edge_index.cpu().numpy()
    w = np.ones(r.shape[0]) if edge_attr is None else edge_attr.cpu().numpy()
    return sp.coo_matrix((w, (r, c)), shape=(num_nodes, num_nodes))


def main():
    ref_loader.load()
    import gnn.utils as U
    U.knn_graph = knn_graph_published
    U.to_scipy_sparse_matrix = to_scipy_sparse_matrix_published
    for name, (n, f, k) in {"knn_small": (300, 9, 3), "knn_k7": (500, 16, 7)}.items():
        rng = np.random.Generator(np.random.PCG64(n))
        x = rng.standard_normal((n, f)).astype(np.float32)
        adj, edge_index = U.get_knn_graph(torch.from_numpy(x), k=k, return_edge_index=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), x=x, k=np.int64(k),
                            adj_bits=np.packbits(adj.numpy().astype(np.uint8), axis=1),
                            edge_index=edge_index.numpy().astype(np.int32))
        print(f"[golden] {name}: n={n} k={k} nnz={int(adj.sum())}")


if __name__ == "__main__":
    main()
