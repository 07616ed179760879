"""TEST INFRASTRUCTURE — golden vectors for the predictive path on the Kron posterior,
produced by the UNMODIFIED reference on CPU (dev container only):

    python oracle/make_golden_predictive.py

For the small fixtures already under tests/golden/ it rebuilds the reference's dense GCN with the
stored weights, fits the reference KronLaplace, and stores
  eps      fixed standard-normal draws [S, P]
  bmm      la.posterior_precision.bmm(eps, exponent=-0.5)          (laplace/utils/matrix.py:396-475)
  samples  la.mean + bmm — what KronLaplace.sample returns for those draws (baselaplace.py:1646-1655)
  py       la(eval_idx, pred_type="nn", link_approx="mc", n_samples=S) with sample() pinned to `samples`
           (baselaplace.py:1183-1199; the call gnn/marglik_training.py:341-353 makes)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.make_golden import CASES, GOLDEN_DIR, dense_adj_from_edges  # noqa: E402

NAMES = ["tiny_undirected_2l", "tiny_directed_3l", "tiny_symmetrised_2l"]
S = 6


def main():
    R = ref_loader.load()
    from torch.utils.data import DataLoader, TensorDataset
    for name in NAMES:
        cfg = CASES[name]
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        n, F = int(z["n"]), int(z["F"])
        ei = z["edge_index"].astype(np.int64)
        X = torch.from_numpy(z["x"])
        L = sum(1 for k in z.files if k.startswith("W"))
        C, h = z[f"W{L - 1}"].shape[0], z["W0"].shape[0]
        model = R.GCN(F, h, C, L, X, dense_adj_from_edges(ei, n), dropout_p=0.5,
                      symmetric=cfg.get("symmetric", False))
        with torch.no_grad():
            for l, conv in enumerate(model.convs):
                conv.lin.weight.copy_(torch.from_numpy(z[f"W{l}"]))
                conv.lin.bias.copy_(torch.from_numpy(z[f"b{l}"]))
        model.eval()
        idx_t, y = torch.from_numpy(z["idx"].astype(np.int64)), torch.from_numpy(z["y"].astype(np.int64))
        la = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron")
        la.fit(DataLoader(TensorDataset(idx_t, y), batch_size=int(z["batch_size"]), shuffle=False))
        assert abs(float(la.log_marginal_likelihood()) - float(z["marglik"])) <= 1e-4 * abs(float(z["marglik"]))
        rng = np.random.Generator(np.random.PCG64(1234))
        eps = rng.standard_normal((S, la.n_params)).astype(np.float32)
        with torch.no_grad():
            bmm = la.posterior_precision.bmm(torch.from_numpy(eps), exponent=-0.5).detach()
            samples = (la.mean.reshape(1, -1) + bmm.reshape(S, -1)).detach()
            la.sample = lambda n_samples=100, generator=None: samples[:n_samples]
            eval_idx = np.setdiff1d(np.arange(n), z["idx"]).astype(np.int64)
            py = la(torch.from_numpy(eval_idx), pred_type="nn", link_approx="mc", n_samples=S).detach()
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"pred_{name}.npz"), eps=eps, bmm=bmm.numpy().reshape(S, -1),
                            samples=samples.numpy(), mean=la.mean.detach().numpy(), eval_idx=eval_idx,
                            py=py.numpy(), prior_precision=np.float64(1.0))
        print(f"[golden-pred] {name}: P={la.n_params} py{tuple(py.shape)} rowsum={float(py.sum(1).mean()):.6f}")


if __name__ == "__main__":
    main()
