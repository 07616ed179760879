"""d marglik / dA on the edges (and on candidate entries) — the hand-derived
adjoint of laplace_gnn_b200/structure.py against
  (1) the reference's own dense ``model.adj.grad`` (STEGCN + ``(-marglik).backward()``,
      tests/golden/adjgrad_*.npz from oracle/make_golden_adjgrad.py), and
  (2) the float64 autograd oracle (oracle/adj_grad_oracle.py), itself held to (1).
Tolerance: <= 1e-4 of max |grad| (the Kronecker factors' tolerance in BASELINE.json; observed ~1e-7 with
the CPU double, fp32)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, Golden
from helpers import build_model
from oracle import adj_grad_oracle as AG

ADJGRAD = ["tiny_undirected_2l", "tiny_directed_3l", "tiny_directed_dups_2l", "small_multibatch_2l"]
TOL = 1e-4


def _ref_grad(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"adjgrad_{name}.npz"))
    return -z["neg_marglik_adj_grad"].astype(np.float64), -float(z["neg_marglik"])


@pytest.mark.parametrize("name", ADJGRAD)
def test_oracle_matches_reference_adj_grad(name):
    g = Golden(name)
    ref, ref_ml = _ref_grad(name)
    ml, grad = AG.marglik_adj_grad(AG.dense_adj01(g.edge_index, g.n), g.x, g.Ws, g.bs, g.idx, g.y,
                                   batch_size=g.batch_size)
    assert abs(ml - ref_ml) <= 1e-5 * abs(ref_ml)
    assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.abs(np.diag(grad)).max() == 0.0


def _check_package(name, device, mode="reference", prior=1.0):
    import laplace_gnn_b200 as L
    from laplace_gnn_b200.structure import marglik_edge_grad
    g = Golden(name)
    model = build_model(g, device)
    idx, y = torch.from_numpy(g.idx).to(device), torch.from_numpy(g.y).to(device)
    A = AG.dense_adj01(g.edge_index, g.n)
    ml, dense = AG.marglik_adj_grad(A, g.x, g.Ws, g.bs, g.idx, g.y, prior_prec=prior, mode=mode,
                                    batch_size=g.batch_size)
    # candidates: every non-edge of the first rows, plus a diagonal entry (must come back as 0)
    cm, ck = np.nonzero(A[:12] + np.eye(g.n)[:12] == 0)
    cand = torch.from_numpy(np.stack([np.append(cm, 3), np.append(ck, 3)])).to(device)
    res = marglik_edge_grad(model, idx, y, prior_precision=prior, hess_sqrt=mode, candidates=cand, group=2,
                            batch_size=None if g.batch_size == len(g.idx) else g.batch_size)
    scale = np.abs(dense).max()
    rows, cols = res.rows.cpu().numpy(), res.cols.cpu().numpy()
    assert abs(float(res.marglik) - ml) <= 1e-4 * abs(ml)
    assert np.abs(res.grad_edges.cpu().numpy() - dense[rows, cols]).max() <= TOL * scale
    gc = res.grad_candidates.cpu().numpy()
    assert np.abs(gc[:-1] - dense[cm, ck]).max() <= TOL * scale and gc[-1] == 0.0
    # every edge of A is there once, self loops included (with gradient 0)
    A1 = A.copy(); np.fill_diagonal(A1, 1)
    assert rows.size == int(A1.sum()) and np.all(A1[rows, cols] == 1)
    assert np.all(res.grad_edges.cpu().numpy()[rows == cols] == 0)
    return res, dense


@pytest.mark.parametrize("name", ADJGRAD)
def test_edge_grad_matches_reference_cpu_double(name, fake_ops):
    res, dense = _check_package(name, "cpu")
    ref, ref_ml = _ref_grad(name)
    rows, cols = res.rows.numpy(), res.cols.numpy()
    assert np.abs(res.grad_edges.numpy() - ref[rows, cols]).max() <= TOL * np.abs(ref).max()
    assert abs(float(res.marglik) - ref_ml) <= 1e-3 * abs(ref_ml)


def test_edge_grad_ggn_mode_and_prior_cpu_double(fake_ops):
    _check_package("tiny_directed_3l", "cpu", mode="ggn", prior=0.3)


def test_edge_param_backward_fills_grad_like_the_reference(fake_ops):
    """``(-marglik).backward()`` on a sparse edge parameter, the reference's usage pattern
    (gnn/marglik_training.py:207-215)."""
    from laplace_gnn_b200.structure import log_marginal_likelihood_of_edges
    g = Golden("tiny_undirected_2l")
    model = build_model(g)
    idx, y = torch.from_numpy(g.idx), torch.from_numpy(g.y)
    w = torch.ones(model.graph.ahat_t.nnz, requires_grad=True)
    neg = -log_marginal_likelihood_of_edges(model, idx, y, w)
    neg.backward()
    ref, ref_ml = _ref_grad("tiny_undirected_2l")
    at = model.graph.ahat_t
    rows = np.repeat(np.arange(g.n), np.diff(at.rowptr.numpy()))
    assert abs(float(neg) + ref_ml) <= 1e-3 * abs(ref_ml)
    assert np.abs(w.grad.numpy() + ref[rows, at.col.numpy()]).max() <= TOL * np.abs(ref).max()
    with pytest.raises(ValueError):
        log_marginal_likelihood_of_edges(model, idx, y, torch.ones(3, requires_grad=True))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ADJGRAD)
def test_edge_grad_matches_reference_gpu(name):
    res, dense = _check_package(name, "cuda:0")
    ref, ref_ml = _ref_grad(name)
    rows, cols = res.rows.cpu().numpy(), res.cols.cpu().numpy()
    assert np.abs(res.grad_edges.cpu().numpy() - ref[rows, cols]).max() <= TOL * np.abs(ref).max()


@pytest.mark.gpu
def test_edge_grad_medium_graph_gpu():
    """2,000 nodes, 3 layers, h = 64, 10 classes (tcgen05 SYRK, several column groups) against the
    float64 autograd oracle on the dense adjacency."""
    import laplace_gnn_b200 as L
    from laplace_gnn_b200.structure import marglik_edge_grad
    from oracle import gcn_kfac_oracle as O
    n, U, F, h, C, layers = 2000, 6000, 24, 64, 10, 3
    ei = O.synthetic_edges(n, U, seed=5, directed=True)
    gen = torch.Generator().manual_seed(5)
    X = torch.randn(n, F, generator=gen)
    torch.manual_seed(5)
    model = L.SparseGCN(F, h, C, layers, X.to("cuda:0"), L.Graph.from_edge_index(torch.from_numpy(ei).to("cuda:0"), n)).to("cuda:0")
    idx = torch.randperm(n, generator=gen)[: int(0.6 * n)].sort().values
    y = torch.randint(0, C, (idx.numel(),), generator=gen)
    Ws = [c.lin.weight.detach().cpu().numpy() for c in model.convs]
    bs = [c.lin.bias.detach().cpu().numpy() for c in model.convs]
    A = AG.dense_adj01(ei, n)
    ml, dense = AG.marglik_adj_grad(A, X.numpy(), Ws, bs, idx.numpy(), y.numpy())
    cand = torch.randint(0, n, (2, 5000), generator=gen)
    res = marglik_edge_grad(model, idx.to("cuda:0"), y.to("cuda:0"), candidates=cand.to("cuda:0"), group=4)
    rows, cols = res.rows.cpu().numpy(), res.cols.cpu().numpy()
    scale = np.abs(dense).max()
    assert abs(float(res.marglik) - ml) <= 1e-4 * abs(ml)
    assert np.abs(res.grad_edges.cpu().numpy() - dense[rows, cols]).max() <= 1e-3 * scale
    cm, ck = cand[0].numpy(), cand[1].numpy()
    off = (A[cm, ck] == 0) & (cm != ck)                       # candidates are non-edges by contract
    assert np.abs(res.grad_candidates.cpu().numpy()[off] - dense[cm[off], ck[off]]).max() <= 1e-3 * scale
    res1 = marglik_edge_grad(model, idx.to("cuda:0"), y.to("cuda:0"), group=10)
    assert float((res1.grad_edges - res.grad_edges).abs().max()) <= 1e-4 * scale


@pytest.mark.gpu
def test_edge_grad_ggn_mode_gpu():
    _check_package("tiny_directed_dups_2l", "cuda:0", mode="ggn", prior=2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [1, 5, 48, 100, 512, 516, 3072])
def test_sddmm_matches_torch(d):
    from laplace_gnn_b200 import ops
    gen = torch.Generator(device="cuda:0").manual_seed(d)
    n, e = 3000, 20_000
    u = torch.randn(n, d + (4 if d % 4 == 0 else 3), device="cuda:0", generator=gen)
    v = torch.randn(n, d, device="cuda:0", generator=gen)
    rows = torch.randint(0, n, (e,), device="cuda:0", generator=gen, dtype=torch.int32)
    cols = torch.randint(0, n, (e,), device="cuda:0", generator=gen, dtype=torch.int32)
    ref = (u[rows.long(), :d].double() * v[cols.long(), :d].double()).sum(1)
    out = ops.sddmm(rows, cols, u, v, d)
    tol = 1e-5 * float(ref.abs().max()) + 1e-6
    assert float((out.double() - ref).abs().max()) <= tol
    ops.sddmm(rows, cols, u, v, d, out=out, accumulate=True)
    assert float((out.double() - 2 * ref).abs().max()) <= 2 * tol
    assert ops.sddmm(rows[:0], cols[:0], u, v, d).numel() == 0


def test_edge_scores_step_reproduces_the_reference_adj_update(fake_ops):
    """Tracking ALL off-diagonal entries, one SGD step on the straight-through scores must give the
    binarised adjacency the reference gets from ``adj_optimizer.step()`` on its dense parameter
    (gnn/marglik_training.py:213-220): adj - lr * adj.grad, thresholded at 0.5."""
    import laplace_gnn_b200 as L
    name, lr = "tiny_directed_dups_2l", 0.4
    g = Golden(name)
    model = build_model(g)
    idx, y = torch.from_numpy(g.idx), torch.from_numpy(g.y)
    n = g.n
    A = AG.dense_adj01(g.edge_index, n)
    allpairs = torch.from_numpy(np.stack(np.nonzero(np.ones((n, n)) - np.eye(n))))
    es = L.EdgeScores(torch.from_numpy(g.edge_index), n, candidates=allpairs)
    assert es.score.numel() == n * n - n and int(es.active.sum()) == int((A * (1 - np.eye(n))).sum())
    opt = torch.optim.SGD([es.score], lr=lr)
    ml = es.neg_marglik_step(model, idx, y, opt)
    z = np.load(os.path.join(GOLDEN_DIR, f"adjgrad_{name}.npz"))
    assert abs(float(ml) + float(z["neg_marglik"])) <= 1e-3 * abs(float(z["neg_marglik"]))
    new_adj = A - lr * z["neg_marglik_adj_grad"].astype(np.float64)
    want = (new_adj > 0.5) & ~np.eye(n, dtype=bool)
    m, k = es.entries.numpy()
    # entries whose updated score is within rounding of the threshold are not decided by fp32
    clear = np.abs(new_adj[m, k] - 0.5) > 1e-4
    assert np.array_equal(es.active.numpy()[clear], want[m, k][clear])
    assert np.abs(es.score.detach().numpy() - new_adj[m, k]).max() <= 1e-4
    assert (es.active.numpy() != (A[m, k] > 0)).sum() > 0                     # the step did flip entries
    # the model now runs on the rebuilt graph: its pattern is the re-binarised matrix + the diagonal
    at = model.graph.ahat_t
    rows = np.repeat(np.arange(n), np.diff(at.rowptr.numpy()))
    got = np.zeros((n, n), bool); got[rows, at.col.numpy()] = True
    assert np.array_equal(got[m, k], es.active.numpy()) and got.diagonal().all()


def test_structure_learning_inside_the_epoch_loop(fake_ops):
    import laplace_gnn_b200 as L
    g = Golden("tiny_undirected_2l")
    model = build_model(g)
    idx, y = torch.from_numpy(g.idx), torch.from_numpy(g.y)
    rng = np.random.Generator(np.random.PCG64(0))
    cand = torch.from_numpy(rng.integers(0, g.n, (2, 200)))
    es = L.EdgeScores(torch.from_numpy(g.edge_index), g.n, candidates=cand)
    before = int(es.active.sum())
    res = L.marglik_training(model, idx, y, idx[:8], y[:8], n_epochs=4, edge_scores=es, lr_adj=0.3,
                             n_hypersteps=2, n_epochs_burnin=2, marglik_frequency=2)
    assert [e for e, _ in res.n_edges] == [2, 4] and len(res.neg_margliks) == 4
    assert res.n_edges[-1][1] == int(es.active.sum()) and int(es.active.sum()) != before
    assert model.graph.ahat_t.nnz == int(es.active.sum()) + g.n
