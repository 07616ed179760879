"""SURVEY 8(f) row 1 pinned: `laplace_gnn_b200.training.marglik_training` against the per-epoch trajectory of the
reference's OWN `gnn.marglik_training.marglik_optimization` (gnn/marglik_training.py:41-329; goldens made by
oracle/make_golden_epochs.py with dropout off): train loss, -log marginal likelihood and validation loss of every
epoch, the two selected epochs and the final weights.  CPU double here, the same on the device with `-m gpu`."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

CASES = ["epochs_small_2l", "epochs_cora_shape_2l"]


def _run(name, device):
    import laplace_gnn_b200 as L
    from laplace_gnn_b200.training import marglik_training
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, F = int(z["n"]), int(z["F"])
    x = z["x"] if z["x"].size else np.unpackbits(z["x_bits"], axis=1)[:, :F].astype(np.float32)
    L_ = sum(1 for k in z.files if k.startswith("W") and not k.startswith("Wend"))
    Ws, bs = [z[f"W{l}"] for l in range(L_)], [z[f"b{l}"] for l in range(L_)]
    graph = L.Graph.from_edge_index(torch.from_numpy(z["edge_index"].astype(np.int64)).to(device), n)
    model = L.SparseGCN(F, Ws[0].shape[0], Ws[-1].shape[0], L_, torch.from_numpy(x).to(device), graph,
                        dropout_p=0.0).to(device)
    with torch.no_grad():
        for l, conv in enumerate(model.convs):
            conv.lin.weight.copy_(torch.from_numpy(Ws[l]))
            conv.lin.bias.copy_(torch.from_numpy(bs[l]))
    t = lambda k: torch.from_numpy(z[k].astype(np.int64)).to(device)
    res = marglik_training(model, t("train_idx"), t("train_y"), t("val_idx"), t("val_y"), n_epochs=int(z["epochs"]),
                           lr=float(z["lr"]), weight_decay=float(z["weight_decay"]))
    rel = lambda a, b: np.abs(np.asarray(a) - b).max() / np.abs(b).max()
    assert rel(res.losses, z["losses"]) <= 1e-4, (res.losses, z["losses"])
    assert rel(res.val_losses, z["val_losses"]) <= 1e-4
    assert rel(res.neg_margliks, z["neg_margliks"]) <= 1e-3, (res.neg_margliks, z["neg_margliks"])   # north-star marglik tolerance
    assert res.best_marglik_epoch == int(z["best_marglik_epoch"]) and res.best_valloss_epoch == int(z["best_valloss_epoch"])
    for l, conv in enumerate(model.convs):
        w = conv.lin.weight.detach().cpu().numpy()
        assert np.abs(w - z[f"Wend{l}"]).max() <= 1e-4 * np.abs(z[f"Wend{l}"]).max()


@pytest.mark.parametrize("name", CASES)
def test_epoch_loop_matches_the_reference_trajectory(fake_ops, name):
    _run(name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_epoch_loop_matches_the_reference_trajectory_on_the_device(name):
    _run(name, "cuda:0")


# ---------------------------------------------------------------------------------- kNN graph recipe (8f row 4)
def _knn(name, device):
    import laplace_gnn_b200 as L
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    x, k = torch.from_numpy(z["x"]).to(device), int(z["k"])
    n = x.shape[0]
    want = np.unpackbits(z["adj_bits"], axis=1)[:, :n]                        # the reference's dense 0/1 adjacency
    g = L.Graph.from_edge_index(L.knn_edge_index(x, k, chunk=128), n, symmetric=True)
    rowptr, col = g.ahat_t.rowptr.cpu().numpy(), g.ahat_t.col.cpu().numpy()  # pattern of A
    got = np.zeros_like(want)
    got[np.repeat(np.arange(n), np.diff(rowptr)), col] = 1
    assert np.array_equal(got, want)
    assert np.array_equal(g.deg.cpu().numpy(), want.sum(1))
    # the reference's edge list of the same graph (adj_to_edge_index) builds the same CSR
    g2 = L.Graph.from_edge_index(torch.from_numpy(z["edge_index"].astype(np.int64)).to(device), n)
    assert torch.equal(g2.ahat.rowptr, g.ahat.rowptr) and torch.equal(g2.ahat.col, g.ahat.col)
    assert torch.equal(g2.ahat.val, g.ahat.val)


@pytest.mark.parametrize("name", ["knn_small", "knn_k7"])
def test_knn_graph_matches_the_reference_recipe(fake_ops, name):
    """`knn_edge_index` + `Graph.from_edge_index(symmetric=True)` against the adjacency the reference's own
    `get_knn_graph` returns (gnn/utils.py:355-369; golden by oracle/make_golden_knn.py)."""
    _knn(name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["knn_small", "knn_k7"])
def test_knn_graph_matches_the_reference_recipe_on_the_device(name):
    _knn(name, "cuda:0")
