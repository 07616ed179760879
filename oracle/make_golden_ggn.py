"""TEST INFRASTRUCTURE — golden vectors that pin ``hess_sqrt="ggn"`` (dev container only):

    python oracle/make_golden_ggn.py

The fork's vendored curvlinops differs from UPSTREAM curvlinops 2.0 on this path by one expression: the
per-sample Hessian square root is built from ``out`` instead of ``out.detach()`` (curvlinops/kfac.py:631-642; the
upstream statement is still there as a comment, :631-636).  That is SURVEY trap T1: the fork back-propagates
through the square root as well and its factors are not the GGN's.  The alternate backends the north star names
(asdl, backpack) are neither vendored nor installed, so the textbook mode cannot be pinned on them; it CAN be
pinned on upstream curvlinops' arithmetic by running the reference's own classes with that one call restored at
run time: ``curvlinops.kfac.loss_hessian_matrix_sqrt`` is wrapped so that it receives ``out.detach()``.  No
reference file is copied or edited; everything else — hooks, packing, rescaling, Kron.decompose, the marglik —
is the reference's code.

For the small fixtures already under tests/golden/ this stores the factors, loss and log marginal likelihood of
that run as tests/golden/ggn_<name>.npz (inputs and weights are those of <name>.npz).  A self-check first runs the
UNPATCHED reference and requires it to reproduce <name>.npz, so the only difference is the wrapped call.
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

NAMES = ["tiny_undirected_2l", "tiny_directed_3l", "tiny_directed_dups_2l", "tiny_symmetrised_2l",
         "small_multibatch_2l"]


def fit(R, model, idx_t, y, bs):
    from torch.utils.data import DataLoader, TensorDataset
    la = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron")
    la.fit(DataLoader(TensorDataset(idx_t, y), batch_size=bs, shuffle=False))
    return la, float(la.log_marginal_likelihood())


def main():
    R = ref_loader.load()
    import curvlinops.kfac as K
    original = K.loss_hessian_matrix_sqrt

    def upstream(output_one_datum, target_one_datum, loss_func):      # upstream curvlinops: .detach() on the input
        return original(output_one_datum.detach(), target_one_datum, loss_func)

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
        bs = int(z["batch_size"])
        # self-check: the unpatched reference reproduces the committed golden
        la0, ml0 = fit(R, model, idx_t, y, bs)
        assert abs(ml0 - float(z["marglik"])) <= 1e-6 * abs(float(z["marglik"])), (name, ml0, float(z["marglik"]))
        K.loss_hessian_matrix_sqrt = upstream
        try:
            la, ml = fit(R, model, idx_t, y, bs)
        finally:
            K.loss_hessian_matrix_sqrt = original
        out = {"loss": np.float64(float(la.loss)), "marglik": np.float64(ml),
               "n_blocks": np.int64(len(la.H_facs.kfacs)), "fork_marglik": np.float64(ml0)}
        for b, blk in enumerate(la.H_facs.kfacs):
            for j, Hi in enumerate(blk):
                out[f"kfac_{b}_{j}"] = Hi.detach().numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"ggn_{name}.npz"), **out)
        print(f"[golden ggn] {name}: marglik {ml:.6f} (fork: {ml0:.6f})")


if __name__ == "__main__":
    main()
