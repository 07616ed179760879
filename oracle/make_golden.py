"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference via oracle/ref_loader.py) on CPU.  Dev container only.

    python oracle/make_golden.py [--only tiny_undirected_2l ...]

Each file stores the inputs (edge list, features, labels, train indices, weights)
and the reference's outputs for one call chain
    la = Laplace(model, "classification", "all", "kron"); la.fit(loader); la.log_marginal_likelihood()
(gnn/marglik_training.py:261-269): loss, every Kronecker factor in Kron.kfacs
order, the log marginal likelihood, plus Â (dense, small cases) and the logits.

Tiers: O1 = reference dense ``gnn.models.GCN``; O2 = reference
``GCNConv`` + laplace + vendored curvlinops with Â given as a torch sparse CSR tensor
(needed for the Pubmed shape: the dense model does 2 N^3 matmuls per forward).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gcn_kfac_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: dict(n, undirected pairs U, F, C, h, L, directed, symmetric flag, tier, feature kind)
    "tiny_undirected_2l": dict(n=40, U=70, F=10, C=4, h=8, L=2, directed=False, tier="O1", feat="normal"),
    "tiny_directed_3l": dict(n=50, U=160, F=12, C=5, h=8, L=3, directed=True, tier="O1", feat="normal"),
    "tiny_directed_dups_2l": dict(n=30, U=90, F=6, C=3, h=5, L=2, directed=True, tier="O1", feat="normal",
                                  dups=True, isolated=True),
    "tiny_symmetrised_2l": dict(n=36, U=60, F=7, C=3, h=6, L=2, directed=True, symmetric=True, tier="O1",
                                feat="normal"),
    "small_multibatch_2l": dict(n=300, U=900, F=16, C=4, h=8, L=2, directed=False, tier="O1", feat="normal",
                                batch_size=64),
    "cora_shape": dict(n=2708, U=5278, F=1433, C=7, h=16, L=2, directed=False, tier="O1", feat="bow"),
    "pubmed_shape": dict(n=19717, U=44324, F=500, C=3, h=64, L=2, directed=False, tier="O2", feat="bow"),
    # the arxiv / products kernel shapes at a size the reference finishes in seconds: 3 layers, h = 256 (fused
    # tcgen05 GEMM, tcgen05 SYRK n = 256, unit-compacted slabs), C = 40 (column groups 16 + 16 + 8)
    "arxiv_mini_3l": dict(n=1200, U=8000, F=128, C=40, h=256, L=3, directed=False, tier="O2", feat="normal"),
    # the products shape in small: 47 classes (logits pitch padded to 48, column groups 16 + 16 + 15(+1)), F = 100
    "products_mini_3l": dict(n=1000, U=12000, F=100, C=47, h=256, L=3, directed=False, tier="O2", feat="normal"),
}


def make_inputs(name, cfg, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    n = cfg["n"]
    ei = O.synthetic_edges(n, cfg["U"], seed=seed, directed=cfg.get("directed", False))
    if cfg.get("isolated"):
        # make the last three nodes isolated (deg = self loop only)
        keep = (ei[0] < n - 3) & (ei[1] < n - 3)
        ei = ei[:, keep]
    if cfg.get("dups"):
        ei = np.concatenate([ei, ei[:, :25], ei[:, :7]], axis=1)  # duplicates (and triplicates)
    if cfg["feat"] == "bow":
        x = (rng.random((n, cfg["F"])) < 0.0127).astype(np.float32)
    else:
        x = rng.standard_normal((n, cfg["F"])).astype(np.float32)
    y_all = rng.integers(0, cfg["C"], n).astype(np.int64)
    perm = rng.permutation(n)
    idx = np.sort(perm[: int(0.6 * n)]).astype(np.int64)
    return ei, x, y_all, idx


class SparseRefGCN(torch.nn.Module):
    """O2: the reference's own GCNConv layers, Â supplied as a sparse CSR tensor.
    Mirrors BaseGNN.forward in eval mode (base_gnn.py:136-161)."""

    def __init__(self, R, dims, X, ahat_sparse):
        super().__init__()
        self.X = X
        self.ahat = ahat_sparse
        self.convs = torch.nn.ModuleList(
            [R.GCNConv(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])

    def forward(self, idx):
        x = self.X
        for conv in self.convs[:-1]:
            x = torch.relu(conv(self.ahat, x))
        return self.convs[-1](self.ahat, x)[idx]


def dense_adj_from_edges(ei, n):
    """Exactly the driver's recipe (gnn/utils.py:325-330, marglik_training.py:403-405)."""
    import scipy.sparse as sp
    a = sp.coo_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(n, n)).toarray()
    adj = torch.tensor(a, dtype=torch.int64).float()
    adj[adj > 1] = 1
    return adj


def run_case(name, cfg, R):
    torch.manual_seed(0)
    ei, x, y_all, idx = make_inputs(name, cfg)
    n, F, C, h, L = cfg["n"], cfg["F"], cfg["C"], cfg["h"], cfg["L"]
    X = torch.from_numpy(x)
    y = torch.from_numpy(y_all[idx])
    idx_t = torch.from_numpy(idx)
    dims = [F] + [h] * (L - 1) + [C]
    out = {}
    if cfg["tier"] == "O1":
        adj = dense_adj_from_edges(ei, n)
        model = R.GCN(F, h, C, L, X, adj, dropout_p=0.5, symmetric=cfg.get("symmetric", False))
        with torch.no_grad():
            out["ahat_dense"] = model.forward_adj().numpy()
    else:
        g = O.build_graph(ei, n, cfg.get("symmetric", False))
        ahat = torch.sparse_csr_tensor(torch.from_numpy(g.rowptr), torch.from_numpy(g.col.astype(np.int64)),
                                       torch.from_numpy(g.val), size=(n, n))
        model = SparseRefGCN(R, dims, X, ahat)
    # a few Adam steps so that the softmax is not degenerate
    opt = torch.optim.Adam([p for k, p in model.named_parameters() if "adj" not in k], lr=0.01)
    model.eval()  # no dropout: keeps the golden independent of torch's dropout RNG stream
    for _ in range(10):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(idx_t), y)
        loss.backward()
        opt.step()
    weights = [c.lin.weight.detach().numpy().copy() for c in model.convs]
    biases = [c.lin.bias.detach().numpy().copy() for c in model.convs]

    from torch.utils.data import DataLoader, TensorDataset
    bs = cfg.get("batch_size", len(idx))
    loader = DataLoader(TensorDataset(idx_t, y), batch_size=bs, shuffle=False)
    t0 = time.time()
    la = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron")
    la.fit(loader)
    ml = la.log_marginal_likelihood()
    dt = time.time() - t0
    with torch.no_grad():
        out["logits"] = model(idx_t).numpy()
    out.update(
        edge_index=ei.astype(np.int32), n=np.int64(n), symmetric=np.bool_(cfg.get("symmetric", False)),
        x_bits=np.packbits(x.astype(np.uint8), axis=1) if cfg["feat"] == "bow" else np.zeros(0, np.uint8),
        x=x if cfg["feat"] != "bow" else np.zeros(0, np.float32),
        F=np.int64(F), y=y_all[idx], idx=idx, batch_size=np.int64(bs),
        loss=np.float64(float(la.loss)), marglik=np.float64(float(ml)),
        ref_fit_seconds=np.float64(dt), tier=np.str_(cfg["tier"]),
    )
    for l in range(L):
        out[f"W{l}"] = weights[l]
        out[f"b{l}"] = biases[l]
    k = 0
    for b, Fb in enumerate(la.H_facs.kfacs):
        for j, Hi in enumerate(Fb):
            out[f"kfac_{b}_{j}"] = Hi.detach().numpy()
            k += 1
    out["n_blocks"] = np.int64(len(la.H_facs.kfacs))
    # exact diagonal GGN through the reference's DiagLaplace (tiny cases only: n_t*C*P memory)
    if n <= 64:
        lad = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="diag")
        lad.fit(DataLoader(TensorDataset(idx_t, y), batch_size=len(idx), shuffle=False))
        out["diag_H"] = lad.H.detach().numpy()
        out["diag_marglik"] = np.float64(float(lad.log_marginal_likelihood()))
    print(f"[golden] {name}: loss={float(la.loss):.6f} marglik={float(ml):.6f} fit+ml={dt:.2f}s")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    R = ref_loader.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in CASES.items():
        if args.only and name not in args.only:
            continue
        out = run_case(name, cfg, R)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)


if __name__ == "__main__":
    main()
