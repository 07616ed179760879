"""Graph ingest from on-disk datasets: the raw files the reference's ``load_data``
reads through torch_geometric (gnn/utils.py:203-322) parsed directly — torch_geometric / ogb are not
dependencies of this package — plus the reference's own 60/20/20 split.

  load_planetoid_raw(raw_dir, name)   ind.<name>.{x,tx,allx,y,ty,ally,graph,test.index} (Cora / Citeseer /
                                      Pubmed as distributed by Planetoid; what ``Planetoid(root, name)[0]``
                                      parses, gnn/utils.py:205-206)
  load_ogb_raw(raw_dir)               edge.csv(.gz), node-feat.csv(.gz), node-label.csv(.gz) of an OGB
                                      node-property dataset (ogbn-arxiv / ogbn-products, BASELINE configs 3-4)
  load_npz_graph(path)                a generic {edge_index, x, y} archive
  reference_split(n, n_rand_splits)   gnn/utils.py:283-321: ShuffleSplit(train_size=0.8, random_state=0), then
                                      ShuffleSplit(train_size=0.6, random_state=0) inside it
  to_device_model_inputs(data, ...)   edge list -> CSR graph on the device (the integer kernels), X, labels

These run on the host (file parsing, numpy); everything downstream of the edge list is the CUDA path.
"""
from __future__ import annotations

import gzip
import os
import pickle
from dataclasses import dataclass, field

import numpy as np
import torch


@dataclass
class NodeDataset:
    """What the reference uses of a PyG ``Data`` object: x, y, edge_index (+ its split)."""
    x: torch.Tensor                 # float32 [n, F]
    y: torch.Tensor                 # int64 [n]
    edge_index: torch.Tensor        # int64 [2, E], A[src, dst] = 1
    name: str = ""
    train_indices: torch.Tensor | None = None       # int64 [n_train, n_rand_splits] (reference layout)
    val_indices: torch.Tensor | None = None
    test_indices: torch.Tensor | None = None
    meta: dict = field(default_factory=dict)

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_classes(self) -> int:
        return int(self.y.max()) + 1 if self.y.numel() else 0


def _coalesce_undirected(src: np.ndarray, dst: np.ndarray, n: int) -> np.ndarray:
    """Drop self loops, add the reverse of every edge, sort by (src, dst), drop duplicates — what
    PyG's planetoid reader does with the adjacency dict (remove_self_loops + to_undirected/coalesce)."""
    keep = src != dst
    src, dst = src[keep], dst[keep]
    s = np.concatenate([src, dst]).astype(np.int64)
    d = np.concatenate([dst, src]).astype(np.int64)
    key = np.unique(s * n + d)
    return np.stack([key // n, key % n])


def load_planetoid_raw(raw_dir: str, name: str) -> NodeDataset:
    """Parse the Planetoid distribution files in ``raw_dir`` (``ind.cora.x`` ...).  Node order, the
    re-insertion of Citeseer's isolated test nodes as zero rows and the label of a node follow the
    published format (Yang et al. 2016, as read by torch_geometric.io.read_planetoid_data)."""
    name = name.lower()

    def read(ext):
        path = os.path.join(raw_dir, f"ind.{name}.{ext}")
        if ext == "test.index":
            with open(path) as f:
                return np.array([int(line) for line in f.read().split()], dtype=np.int64)
        with open(path, "rb") as f:
            obj = pickle.load(f, encoding="latin1")
        if ext == "graph":
            return obj
        return np.asarray(obj.todense() if hasattr(obj, "todense") else obj, dtype=np.float32)

    tx, allx = read("tx"), read("allx")
    ty, ally = read("ty"), read("ally")
    graph, test_index = read("graph"), read("test.index")
    sorted_test = np.sort(test_index)
    if name == "citeseer":
        # test indices with gaps: the missing nodes are isolated and get zero features / label 0
        span = int(test_index.max() - test_index.min() + 1)
        tx_ext = np.zeros((span, tx.shape[1]), np.float32)
        ty_ext = np.zeros((span, ty.shape[1]), np.float32)
        tx_ext[sorted_test - test_index.min()] = tx
        ty_ext[sorted_test - test_index.min()] = ty
        tx, ty = tx_ext, ty_ext
    x = np.concatenate([allx, tx], axis=0)
    y = np.concatenate([ally, ty], axis=0).argmax(axis=1).astype(np.int64)
    x[test_index] = x[sorted_test]
    y[test_index] = y[sorted_test]
    n = x.shape[0]
    src = np.fromiter((k for k, v in graph.items() for _ in v), dtype=np.int64)
    dst = np.fromiter((j for v in graph.values() for j in v), dtype=np.int64)
    ei = _coalesce_undirected(src, dst, n)
    return NodeDataset(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(ei), name,
                       meta={"n_train_planetoid": int(read("y").shape[0]), "format": "planetoid"})


def _read_csv(path: str, dtype):
    for p in (path, path + ".gz"):
        if os.path.exists(p):
            opener = gzip.open if p.endswith(".gz") else open
            with opener(p, "rt") as f:
                arr = np.loadtxt(f, delimiter=",", dtype=dtype, ndmin=2)
            return arr
    raise FileNotFoundError(path + "[.gz]")


def load_ogb_raw(raw_dir: str, name: str = "", undirected: bool = True) -> NodeDataset:
    """``raw/`` directory of an OGB node-property-prediction dataset: ``edge.csv`` (src,dst per line, each
    edge once), ``node-feat.csv``, ``node-label.csv`` (optionally gzipped).  ``undirected=True`` adds the
    reverse edges — ogbn-products is undirected, and GCN use of ogbn-arxiv symmetrises its citations
    (nnz ≈ 2U + N)."""
    edges = _read_csv(os.path.join(raw_dir, "edge.csv"), np.int64)
    x = _read_csv(os.path.join(raw_dir, "node-feat.csv"), np.float32)
    y = _read_csv(os.path.join(raw_dir, "node-label.csv"), np.float64)
    y = np.nan_to_num(y[:, 0], nan=-1).astype(np.int64)
    if edges.shape[1] != 2:
        raise ValueError("edge.csv must hold two columns (source, target)")
    ei = edges.T
    if undirected:
        ei = np.concatenate([ei, ei[::-1]], axis=1)
    return NodeDataset(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(y),
                       torch.from_numpy(np.ascontiguousarray(ei)), name, meta={"format": "ogb-raw"})


def load_npz_graph(path: str) -> NodeDataset:
    z = np.load(path)
    return NodeDataset(torch.from_numpy(z["x"].astype(np.float32)), torch.from_numpy(z["y"].astype(np.int64)),
                       torch.from_numpy(z["edge_index"].astype(np.int64)), os.path.basename(path),
                       meta={"format": "npz"})


def reference_split(n: int, n_rand_splits: int = 1):
    """The reference's 60/20/20 split (gnn/utils.py:283-321), same sklearn calls and seeds, returned in its
    layout: int64 [count, n_rand_splits] index matrices (train, val, test)."""
    from sklearn.model_selection import ShuffleSplit
    dummy = np.zeros((n, 1))
    tr, va, te = [], [], []
    outer = ShuffleSplit(n_splits=n_rand_splits, train_size=0.6 + 0.2, random_state=0)
    for train_and_val, test in outer.split(dummy):
        a, b = next(ShuffleSplit(n_splits=1, train_size=0.6, random_state=0).split(dummy[train_and_val]))
        tr.append(train_and_val[a]); va.append(train_and_val[b]); te.append(test)
    stack = lambda parts: torch.tensor(np.array(parts)).t().contiguous()
    return stack(tr), stack(va), stack(te)


def with_reference_split(data: NodeDataset, n_rand_splits: int = 1) -> NodeDataset:
    data.train_indices, data.val_indices, data.test_indices = reference_split(data.num_nodes, n_rand_splits)
    return data


def to_device_model_inputs(data: NodeDataset, device, symmetric: bool = False, split: int = 0):
    """(Graph, X, y, train_idx, val_idx, test_idx) on ``device``: the edge list goes through the CSR build /
    normalisation kernels (Graph.from_edge_index), the rest is copied."""
    from .graph import Graph
    graph = Graph.from_edge_index(data.edge_index.to(device), data.num_nodes, symmetric=symmetric)
    pick = lambda t: None if t is None else t[:, split].to(device)
    return (graph, data.x.to(device), data.y.to(device), pick(data.train_indices), pick(data.val_indices),
            pick(data.test_indices))
