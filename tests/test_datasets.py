"""Host-side dataset readers on fabricated files in the published formats: the
Planetoid ``ind.*`` pickles (incl. Citeseer's isolated test nodes), the OGB raw csv layout, the
reference's 60/20/20 split (same sklearn calls as gnn/utils.py:283-321)."""
import gzip
import os
import pickle

import numpy as np
import scipy.sparse as sp
import torch

from laplace_gnn_b200 import datasets as D


def _write_planetoid(tmp, name, n_train, n_all, test_ids, F, C, graph, rng):
    """x/y: labelled train nodes; allx/ally: the first n_all nodes; tx/ty: the test nodes in the order of
    test.index (an arbitrary permutation)."""
    feats = (rng.random((n_all + (max(test_ids) - min(test_ids) + 1), F)) < 0.2).astype(np.float32)
    labels = rng.integers(0, C, feats.shape[0])
    onehot = np.eye(C, dtype=np.int32)[labels]
    allx, ally = feats[:n_all], onehot[:n_all]
    tx, ty = feats[test_ids], onehot[test_ids]
    objs = {"x": sp.csr_matrix(allx[:n_train]), "y": ally[:n_train], "allx": sp.csr_matrix(allx), "ally": ally,
            "tx": sp.csr_matrix(tx), "ty": ty, "graph": graph}
    for ext, obj in objs.items():
        with open(os.path.join(tmp, f"ind.{name}.{ext}"), "wb") as f:
            pickle.dump(obj, f)
    with open(os.path.join(tmp, f"ind.{name}.test.index"), "w") as f:
        f.write("\n".join(str(i) for i in test_ids) + "\n")
    return feats, labels


def test_planetoid_raw_reader(tmp_path):
    rng = np.random.Generator(np.random.PCG64(0))
    n_all, test_ids = 30, [37, 30, 33, 31, 39, 32, 35, 34, 38, 36]            # contiguous range, shuffled
    graph = {i: [int(j) for j in rng.integers(0, 40, 3)] for i in range(40)}
    graph[5].append(5)                                                         # a self loop (dropped)
    feats, labels = _write_planetoid(str(tmp_path), "cora", 10, n_all, test_ids, 12, 4, graph, rng)
    d = D.load_planetoid_raw(str(tmp_path), "Cora")
    assert d.num_nodes == 40 and d.x.shape == (40, 12) and d.num_classes <= 4
    assert np.array_equal(d.x.numpy(), feats[:40]) and np.array_equal(d.y.numpy(), labels[:40])
    ei = d.edge_index.numpy()
    dense = np.zeros((40, 40), bool)
    for k, v in graph.items():
        for j in v:
            if j != k:
                dense[k, j] = dense[j, k] = True
    got = np.zeros((40, 40), bool); got[ei[0], ei[1]] = True
    assert np.array_equal(got, dense) and ei.shape[1] == dense.sum()           # undirected, coalesced, no loops
    assert np.all(np.diff(ei[0] * 40 + ei[1]) > 0)                             # sorted like PyG's coalesce


def test_planetoid_citeseer_isolated_test_nodes(tmp_path):
    rng = np.random.Generator(np.random.PCG64(1))
    n_all, test_ids = 20, [27, 20, 23, 25, 22]                                 # 21, 24, 26 are missing: isolated
    graph = {i: [int((i + 1) % 20)] for i in range(20)}
    feats, labels = _write_planetoid(str(tmp_path), "citeseer", 6, n_all, test_ids, 8, 3, graph, rng)
    d = D.load_planetoid_raw(str(tmp_path), "citeseer")
    assert d.num_nodes == 28
    for i in test_ids:
        assert np.array_equal(d.x[i].numpy(), feats[i]) and int(d.y[i]) == labels[i]
    for i in (21, 24, 26):
        assert float(d.x[i].abs().sum()) == 0.0 and int(d.y[i]) == 0


def test_ogb_raw_reader_and_device_free_fields(tmp_path):
    rng = np.random.Generator(np.random.PCG64(2))
    n, E, F = 50, 120, 6
    edges = rng.integers(0, n, (E, 2))
    x = rng.standard_normal((n, F)).astype(np.float32)
    y = rng.integers(0, 5, n)
    with gzip.open(tmp_path / "edge.csv.gz", "wt") as f:
        np.savetxt(f, edges, fmt="%d", delimiter=",")
    np.savetxt(tmp_path / "node-feat.csv", x, delimiter=",", fmt="%.8e")
    with gzip.open(tmp_path / "node-label.csv.gz", "wt") as f:
        np.savetxt(f, y[:, None], fmt="%d", delimiter=",")
    d = D.load_ogb_raw(str(tmp_path), "toy")
    assert d.edge_index.shape == (2, 2 * E) and np.array_equal(d.edge_index[:, :E].numpy(), edges.T)
    assert np.array_equal(d.edge_index[:, E:].numpy(), edges.T[::-1])
    assert np.allclose(d.x.numpy(), x, rtol=1e-6) and np.array_equal(d.y.numpy(), y)
    d1 = D.load_ogb_raw(str(tmp_path), undirected=False)
    assert d1.edge_index.shape == (2, E)
    np.savez(tmp_path / "g.npz", edge_index=edges.T, x=x, y=y)
    d2 = D.load_npz_graph(str(tmp_path / "g.npz"))
    assert np.array_equal(d2.edge_index.numpy(), edges.T) and d2.num_nodes == n


def test_reference_split_is_the_references():
    """Same calls as gnn/utils.py:283-321 (restated here independently), 60/20/20, disjoint, deterministic."""
    from sklearn.model_selection import ShuffleSplit
    n = 137
    tr, va, te = D.reference_split(n, n_rand_splits=2)
    assert tr.shape[1] == va.shape[1] == te.shape[1] == 2
    data_x = np.zeros((n, 3))
    rs = ShuffleSplit(n_splits=2, train_size=0.8, random_state=0)
    for i, (tv, test) in enumerate(rs.split(data_x)):
        a, b = next(ShuffleSplit(n_splits=1, train_size=0.6, random_state=0).split(data_x[tv]))
        assert np.array_equal(tr[:, i].numpy(), tv[a]) and np.array_equal(va[:, i].numpy(), tv[b])
        assert np.array_equal(te[:, i].numpy(), test)
        assert len(set(tv[a]) | set(tv[b]) | set(test)) == len(tv[a]) + len(tv[b]) + len(test)
    tr2, _, _ = D.reference_split(n, 2)
    assert torch.equal(tr, tr2)


def test_to_device_model_inputs_with_cpu_double(fake_ops, tmp_path):
    rng = np.random.Generator(np.random.PCG64(3))
    edges = rng.integers(0, 30, (2, 80))
    np.savez(tmp_path / "g.npz", edge_index=edges, x=rng.standard_normal((30, 4)), y=rng.integers(0, 3, 30))
    d = D.with_reference_split(D.load_npz_graph(str(tmp_path / "g.npz")))
    graph, X, y, tr, va, te = D.to_device_model_inputs(d, "cpu")
    assert graph.n == 30 and X.shape == (30, 4) and tr.numel() + va.numel() + te.numel() == 30
    from oracle import gcn_kfac_oracle as O
    R = O.build_graph(edges.astype(np.int64), 30)
    assert np.array_equal(graph.ahat.col.numpy(), R.col) and np.array_equal(graph.ahat.val.numpy(), R.val)
