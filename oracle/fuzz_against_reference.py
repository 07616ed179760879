"""TEST INFRASTRUCTURE — differential fuzz of the oracle against the UNMODIFIED reference (dev container only):

    python oracle/fuzz_against_reference.py [--cases 40] [--seed 0]

Random small configurations — directed / undirected / symmetrised graphs, duplicate edges, explicit self loops in
the edge list, isolated nodes, nodes with only in- or only out-edges, 1 to 4 layers, 2 to 9 classes, hidden widths
from 1, batch sizes that split the train set unevenly, repeated train nodes — are run through the reference's dense
``GCN`` + ``Laplace(..., "all", "kron")`` (tier O1) and through ``oracle.fit_and_marglik``; factors, loss and marglik
must agree to fp32 rounding.  Nothing is stored: the committed goldens stay the pin, this widens the net around them.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gcn_kfac_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import dense_adj_from_edges  # noqa: E402


def random_case(rng, wide=False):
    n = int(rng.integers(6, 40))
    directed = bool(rng.integers(0, 2))
    symmetric = bool(directed and rng.integers(0, 3) == 0)
    e = int(rng.integers(0, 4 * n))
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)          # self loops and duplicates included
    if not directed:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    if rng.integers(0, 2) and e > 0:                                  # a few isolated nodes
        iso = rng.permutation(n)[: max(1, n // 8)]
        keep = ~(np.isin(src, iso) | np.isin(dst, iso))
        src, dst = src[keep], dst[keep]
    ei = np.stack([src, dst]).astype(np.int64)
    L, C, h, F = int(rng.integers(1, 5)), int(rng.integers(2, 10)), int(rng.integers(1, 9)), int(rng.integers(1, 7))
    if wide:                                        # hidden widths the unit-compacted slabs take (multiples of 32)
        h, C = int(rng.choice([32, 64, 96])), int(rng.integers(2, 20))
    m = int(rng.integers(1, n + 1))
    idx = np.sort(rng.permutation(n)[:m]).astype(np.int64)
    if rng.integers(0, 4) == 0 and m > 1:
        idx = np.concatenate([idx, idx[:2]])                          # repeated train nodes
    y = rng.integers(0, C, idx.shape[0]).astype(np.int64)
    bs = int(idx.shape[0]) if rng.integers(0, 2) else int(rng.integers(1, idx.shape[0] + 1))
    x = rng.standard_normal((n, F)).astype(np.float32)
    return dict(n=n, ei=ei, symmetric=symmetric, L=L, C=C, h=h, F=F, idx=idx, y=y, bs=bs, x=x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--adjgrad", action="store_true",
                    help="also compare d marglik / d adj: the reference's STEGCN + (-marglik).backward() against "
                         "oracle/adj_grad_oracle.py (and the package's marglik_edge_grad with --package)")
    ap.add_argument("--ggn", action="store_true",
                    help="compare hess_sqrt='ggn' against the reference run with upstream curvlinops' detach restored "
                         "(the wrapper of oracle/make_golden_ggn.py) instead of the fork's mode")
    ap.add_argument("--wide", action="store_true", help="hidden widths 32 / 64 / 96 and up to 19 classes")
    ap.add_argument("--predictive", action="store_true",
                    help="also compare posterior_precision.bmm(eps, -1/2), pinned samples and the MC predictive "
                         "(la(idx, pred_type='nn', link_approx='mc')) with the oracle (and the package with --package)")
    ap.add_argument("--diag", action="store_true",
                    help="also compare the exact diagonal GGN (reference DiagLaplace) against oracle.diag_ggn")
    ap.add_argument("--package", action="store_true",
                    help="also run the package's host logic (B200GGN on the CPU test double, tests/fake_ops.py)")
    args = ap.parse_args()
    if args.package:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        from laplace_gnn_b200.kron import Laplace as StandIn
        for nm in F.ALL:
            setattr(ops, nm, getattr(F, nm))
    R = ref_loader.load()
    from torch.utils.data import DataLoader, TensorDataset
    rng = np.random.default_rng(args.seed)
    worst = {"factor": 0.0, "marglik": 0.0, "loss": 0.0}
    mode = "reference"
    if args.ggn:
        import curvlinops.kfac as K
        original = K.loss_hessian_matrix_sqrt
        K.loss_hessian_matrix_sqrt = lambda out, tgt, lf: original(out.detach(), tgt, lf)
        mode = "ggn"
    for case in range(args.cases):
        c = random_case(rng, args.wide)
        torch.manual_seed(case)
        model = R.GCN(c["F"], c["h"], c["C"], c["L"], torch.from_numpy(c["x"]), dense_adj_from_edges(c["ei"], c["n"]),
                      dropout_p=0.5, symmetric=c["symmetric"])
        model.eval()
        Ws = [conv.lin.weight.detach().numpy().copy() for conv in model.convs]
        bs_ = [conv.lin.bias.detach().numpy().copy() for conv in model.convs]
        idx_t, y_t = torch.from_numpy(c["idx"]), torch.from_numpy(c["y"])
        la = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron")
        la.fit(DataLoader(TensorDataset(idx_t, y_t), batch_size=c["bs"], shuffle=False))
        ml = float(la.log_marginal_likelihood())
        G = O.build_graph(c["ei"], c["n"], c["symmetric"])
        bsz = None if c["bs"] == len(c["idx"]) else c["bs"]
        loss, kfacs, ml_o = O.fit_and_marglik(G, c["x"], Ws, bs_, c["idx"], c["y"], 1.0, mode, torch.float32, bsz)
        tag = (f"case {case}: n={c['n']} E={c['ei'].shape[1]} sym={c['symmetric']} L={c['L']} C={c['C']} h={c['h']} "
               f"F={c['F']} M={len(c['idx'])} bs={c['bs']}")
        assert len(kfacs) == len(la.H_facs.kfacs), tag
        for blk, ref_blk in zip(kfacs, la.H_facs.kfacs):
            for a, b in zip(blk, ref_blk):
                b = b.detach().numpy()
                err = float(np.abs(a.numpy() - b).max() / max(np.abs(b).max(), 1e-30))
                worst["factor"] = max(worst["factor"], err)
                if err > 5e-6:
                    print(f"   note: factor {tuple(b.shape)} differs by {err:.2e} of its max {np.abs(b).max():.3e}")
                assert err <= 5e-5, (tag, "factor", err)
        e_l = abs(float(loss) - float(la.loss)) / max(abs(float(la.loss)), 1e-30)
        e_m = abs(float(ml_o) - ml) / max(abs(ml), 1e-30)
        worst["loss"], worst["marglik"] = max(worst["loss"], e_l), max(worst["marglik"], e_m)
        assert e_l <= 1e-5 and e_m <= 1e-5, (tag, e_l, e_m)
        if args.package:
            graph = L.Graph.from_edge_index(torch.from_numpy(c["ei"]), c["n"], symmetric=c["symmetric"])
            pm = L.SparseGCN(c["F"], c["h"], c["C"], c["L"], torch.from_numpy(c["x"]), graph)
            with torch.no_grad():
                for l, conv in enumerate(pm.convs):
                    conv.lin.weight.copy_(torch.from_numpy(Ws[l]))
                    conv.lin.bias.copy_(torch.from_numpy(bs_[l]))
            kw = {"hess_sqrt": mode, "unit_min_width": 0, "rhs_tile_bytes": int(rng.integers(1, 4)) * 2 * c["n"] * 32 * 4}
            pl = StandIn(pm, "classification", backend=L.B200GGN, backend_kwargs=kw)
            pl.fit(DataLoader(TensorDataset(idx_t, y_t), batch_size=c["bs"], shuffle=False))
            for blk, ref_blk in zip(pl.H_facs.kfacs, la.H_facs.kfacs):
                for a, b in zip(blk, ref_blk):
                    b = b.detach().numpy()
                    assert float(np.abs(a.numpy() - b).max() / max(np.abs(b).max(), 1e-30)) <= 5e-5, (tag, "package factor")
            e_p = abs(float(pl.log_marginal_likelihood()) - ml) / max(abs(ml), 1e-30)
            assert e_p <= 1e-5, (tag, "package marglik", e_p)
        if args.predictive:
            S = 5
            eps = torch.randn(S, la.n_params, generator=torch.Generator().manual_seed(case))
            ref_bmm = la.posterior_precision.bmm(eps, exponent=-0.5).detach()
            ref_samples = (la.mean.reshape(1, -1) + ref_bmm).detach()
            eval_idx = torch.from_numpy(np.sort(rng.permutation(c["n"])[: max(1, c["n"] // 3)]).astype(np.int64))
            orig_sample = la.sample
            la.sample = lambda n_samples=S, generator=None: ref_samples          # pin the draws
            with torch.no_grad():
                ref_py = la(eval_idx, pred_type="nn", link_approx="mc", n_samples=S).detach().numpy()
            la.sample = orig_sample
            kf = [[h.detach() for h in blk] for blk in la.H_facs.kfacs]
            o_bmm = O.kron_bmm(kf, eps, -0.5, 1.0)
            e_b = float((o_bmm - ref_bmm).abs().max() / ref_bmm.abs().max())
            shapes = [w.shape for w in Ws]
            o_py = O.mc_predictive(G, c["x"], ref_samples.numpy(), shapes, eval_idx.numpy())
            e_y = float(np.abs(o_py.numpy() - ref_py).max() / np.abs(ref_py).max())
            worst["bmm"], worst["mc_predictive"] = max(worst.get("bmm", 0.0), e_b), max(worst.get("mc_predictive", 0.0), e_y)
            assert e_b <= 2e-4 and e_y <= 2e-5, (tag, "predictive", e_b, e_y)
            if args.package:
                p_bmm = pl.posterior_precision.bmm(eps, exponent=-0.5)
                e_pb = float((p_bmm - ref_bmm).abs().max() / ref_bmm.abs().max())
                p_py = pl(eval_idx, pred_type="nn", link_approx="mc", samples=ref_samples)
                e_py = float(np.abs(p_py.numpy() - ref_py).max() / np.abs(ref_py).max())
                worst["bmm_package"] = max(worst.get("bmm_package", 0.0), e_pb)
                worst["mc_predictive_package"] = max(worst.get("mc_predictive_package", 0.0), e_py)
                assert e_pb <= 5e-4 and e_py <= 2e-5, (tag, "predictive (package)", e_pb, e_py)
        if args.diag and c["n"] <= 24:
            ld = R.Laplace(model, "classification", subset_of_weights="all", hessian_structure="diag")
            ld.fit(DataLoader(TensorDataset(idx_t, y_t), batch_size=len(c["idx"]), shuffle=False))
            _, dg = O.diag_ggn(G, c["x"], Ws, bs_, c["idx"], c["y"])
            ref_d = ld.H.detach().numpy()
            e_d = float(np.abs(dg.numpy() - ref_d).max() / max(np.abs(ref_d).max(), 1e-30))
            worst["diag"] = max(worst.get("diag", 0.0), e_d)
            assert e_d <= 5e-5, (tag, "diag GGN", e_d)
        if args.adjgrad and not c["symmetric"] and not args.ggn:
            from gnn.models.models import STEGCN
            from oracle import adj_grad_oracle as AG
            ste = STEGCN(c["F"], c["h"], c["C"], c["L"], torch.from_numpy(c["x"]),
                         dense_adj_from_edges(c["ei"], c["n"]).clone(), dropout_p=0.5)
            with torch.no_grad():
                for l, conv in enumerate(ste.convs):
                    conv.lin.weight.copy_(torch.from_numpy(Ws[l]))
                    conv.lin.bias.copy_(torch.from_numpy(bs_[l]))
            ste.eval()
            ls = R.Laplace(ste, "classification", subset_of_weights="all", hessian_structure="kron")
            ls.fit(DataLoader(TensorDataset(idx_t, y_t), batch_size=c["bs"], shuffle=False))
            (-ls.log_marginal_likelihood()).backward()
            ref_grad = -ste.adj.grad.detach().numpy().astype(np.float64)
            A01 = AG.dense_adj01(c["ei"], c["n"])
            ml64, grad = AG.marglik_adj_grad(A01, c["x"], Ws, bs_, c["idx"], c["y"], batch_size=c["bs"])
            scale = max(np.abs(ref_grad).max(), 1e-30)
            e_g = np.abs(grad - ref_grad).max() / scale
            worst["adjgrad"] = max(worst.get("adjgrad", 0.0), float(e_g))
            assert e_g <= 2e-4, (tag, "adj grad (oracle)", e_g)
            if args.package:
                from laplace_gnn_b200.structure import marglik_edge_grad
                res = marglik_edge_grad(pm, idx_t, y_t, group=2,
                                        batch_size=None if c["bs"] == len(c["idx"]) else c["bs"])
                rr, cc = res.rows.numpy(), res.cols.numpy()
                e_pg = np.abs(res.grad_edges.numpy() - ref_grad[rr, cc]).max() / scale
                worst["adjgrad_package"] = max(worst.get("adjgrad_package", 0.0), float(e_pg))
                assert e_pg <= 5e-4, (tag, "adj grad (package)", e_pg)
        print(tag, f"ok (marglik {ml:.5f})", flush=True)
    print("worst relative differences:", worst)


if __name__ == "__main__":
    main()
