"""Predictive path on the Kron posterior against goldens produced by the
reference (oracle/make_golden_predictive.py): KronDecomposed.bmm with exponent -1/2, posterior
samples for fixed draws, and the MC predictive with those samples."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, Golden, max_rel_err
from helpers import build_model, loader_for
from oracle import gcn_kfac_oracle as O

NAMES = ["tiny_undirected_2l", "tiny_directed_3l", "tiny_symmetrised_2l"]


def _pred(name):
    return np.load(os.path.join(GOLDEN_DIR, f"pred_{name}.npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_bmm_samples_and_predictive_match_reference(name):
    g, z = Golden(name), _pred(name)
    kfacs = [[torch.from_numpy(h) for h in blk] for blk in g.kfacs]
    bmm = O.kron_bmm(kfacs, torch.from_numpy(z["eps"]), -0.5, 1.0)
    assert max_rel_err(bmm.numpy(), z["bmm"]) <= 1e-4
    smp = O.posterior_samples(z["mean"], kfacs, z["eps"])
    assert max_rel_err(smp.numpy(), z["samples"]) <= 1e-4
    G = O.build_graph(g.edge_index, g.n, g.symmetric)
    shapes = [w.shape for w in g.Ws]
    py = O.mc_predictive(G, g.x, z["samples"], shapes, z["eval_idx"])
    assert max_rel_err(py.numpy(), z["py"]) <= 1e-5


def _check_package(name, device):
    import laplace_gnn_b200 as L
    g, z = Golden(name), _pred(name)
    model = build_model(g, device)
    la = L.Laplace(model, "classification", backend=L.B200GGN)
    la.fit(loader_for(g, device))
    eps = torch.from_numpy(z["eps"]).to(device)
    bmm = la.posterior_precision.bmm(eps, exponent=-0.5)
    assert max_rel_err(bmm.cpu().numpy(), z["bmm"]) <= 2e-4
    assert max_rel_err(la.mean.cpu().numpy(), z["mean"]) <= 1e-6
    samples = torch.from_numpy(z["samples"]).to(device)
    eval_idx = torch.from_numpy(z["eval_idx"]).to(device)
    py = la(eval_idx, pred_type="nn", link_approx="mc", samples=samples)
    assert max_rel_err(py.cpu().numpy(), z["py"]) <= 1e-5
    from laplace_gnn_b200.predictive import mc_predictive
    py1 = mc_predictive(model, samples, eval_idx, tile_bytes=1)        # one sample per tile
    assert max_rel_err(py1.cpu().numpy(), z["py"]) <= 1e-5
    drawn = la(eval_idx, n_samples=64, generator=torch.Generator(device=device).manual_seed(0))
    assert torch.allclose(drawn.sum(1), torch.ones_like(drawn.sum(1)), atol=1e-5)
    with pytest.raises(NotImplementedError):
        la(eval_idx, pred_type="glm")


@pytest.mark.parametrize("name", NAMES)
def test_package_predictive_matches_reference_cpu_double(name, fake_ops):
    _check_package(name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_package_predictive_matches_reference_gpu(name):
    _check_package(name, "cuda:0")
