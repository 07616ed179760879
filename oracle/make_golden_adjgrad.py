"""TEST INFRASTRUCTURE — generates tests/golden/adjgrad_*.npz by running the UNMODIFIED reference
(/root/reference via oracle/ref_loader.py) on CPU.  Dev container only.

    python oracle/make_golden_adjgrad.py

For each small case of oracle/make_golden.py: the reference's STEGCN (dense ``adj`` parameter,
straight-through binarisation; gnn/models/models.py:65-118) with the golden fixture's weights,
``la = Laplace(model, "classification", "all", "kron"); la.fit(loader);
(-la.log_marginal_likelihood()).backward()`` exactly as gnn/marglik_training.py:197-215 does, and
the resulting dense ``model.adj.grad`` is stored next to the inputs' fixture name.
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
from oracle.make_golden import CASES, GOLDEN_DIR, dense_adj_from_edges, make_inputs  # noqa: E402

ADJGRAD_CASES = ["tiny_undirected_2l", "tiny_directed_3l", "tiny_directed_dups_2l", "small_multibatch_2l"]


def main():
    R = ref_loader.load()
    from gnn.models.models import STEGCN
    from torch.utils.data import DataLoader, TensorDataset
    for name in ADJGRAD_CASES:
        cfg = CASES[name]
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        ei, x, y_all, idx = make_inputs(name, cfg)
        assert np.array_equal(ei, z["edge_index"]) and np.array_equal(idx, z["idx"])
        n, F, C, h, L = cfg["n"], cfg["F"], cfg["C"], cfg["h"], cfg["L"]
        adj = dense_adj_from_edges(ei, n)
        torch.manual_seed(0)
        model = STEGCN(F, h, C, L, torch.from_numpy(x), adj.clone(), dropout_p=0.5)
        with torch.no_grad():
            for l, conv in enumerate(model.convs):
                conv.lin.weight.copy_(torch.from_numpy(z[f"W{l}"]))
                conv.lin.bias.copy_(torch.from_numpy(z[f"b{l}"]))
        model.eval()
        idx_t, y = torch.from_numpy(idx), torch.from_numpy(y_all[idx])
        loader = DataLoader(TensorDataset(idx_t, y), batch_size=cfg.get("batch_size", len(idx)), shuffle=False)
        la = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron")
        la.fit(loader)
        neg = -la.log_marginal_likelihood()
        neg.backward()
        g = model.adj.grad.detach().numpy().copy()
        assert abs(float(neg) + float(z["marglik"])) <= 1e-4 * abs(float(z["marglik"])), (float(neg), float(z["marglik"]))
        np.savez_compressed(os.path.join(GOLDEN_DIR, "adjgrad_" + name + ".npz"),
                            neg_marglik_adj_grad=g.astype(np.float32), neg_marglik=np.float64(float(neg)))
        print(f"[golden] adjgrad_{name}: -marglik={float(neg):.6f} |grad|max={np.abs(g).max():.4f} "
              f"diag max={np.abs(np.diag(g)).max():.1e}")


if __name__ == "__main__":
    main()
