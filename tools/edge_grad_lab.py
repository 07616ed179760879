"""Timing of d marglik / dA (laplace_gnn_b200/structure.py) against one fit on synthetic shapes (GPU box).
usage: edge_grad_lab.py [pubmed|arxiv|products-1/8]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops
from laplace_gnn_b200.structure import marglik_edge_grad
SHAPES = {"pubmed": (19_717, 44_324, 500, 3, 64, 2), "arxiv": (169_343, 1_166_243, 128, 40, 256, 3),
          "products-1/8": (306_128, 7_732_392, 100, 47, 256, 3)}
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["pubmed", "arxiv"]:
    n, u, f, c, h, l = SHAPES[name]
    gen = torch.Generator(device=dev).manual_seed(0)
    src = torch.randint(0, n, (u,), device=dev, generator=gen); dst = torch.randint(0, n, (u,), device=dev, generator=gen)
    graph = L.Graph.from_edge_index(torch.stack([torch.cat([src, dst]), torch.cat([dst, src])]), n, assume_undirected=True)
    X = torch.randn(n, f, device=dev, generator=gen)
    idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
    y = torch.randint(0, c, (idx.numel(),), device=dev, generator=gen)
    torch.manual_seed(0)
    model = L.SparseGCN(f, h, c, l, X, graph).to(dev)
    def fit():
        la = L.Laplace(model, "classification", backend=L.B200GGN)
        la.fit(L.TensorBatchLoader(idx, y)); return la.log_marginal_likelihood()
    def timed(fn, reps=2):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): out = fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps, out
    t_fit, ml = timed(fit)
    ops.profile_begin()
    t_grad, res = timed(lambda: marglik_edge_grad(model, idx, y), reps=1)
    kinds = {}
    for rec in list(ops.PROFILE):
        kinds[rec["kind"]] = kinds.get(rec["kind"], 0.0) + rec["start"].elapsed_time(rec["end"])
    ops.PROFILE = None
    print(f"{name}: n={n} nnz={graph.nnz} fit+marglik {1e3 * t_fit:8.1f} ms | marglik + d/dA on {graph.nnz} entries {1e3 * t_grad:8.1f} ms "
          f"({t_grad / t_fit:4.1f} fits)  marglik {float(ml):.2f} / {float(res.marglik):.2f}  |grad|max {float(res.grad_edges.abs().max()):.3e}", flush=True)
    print("    ms by kernel kind (last call):", {k: round(v, 1) for k, v in sorted(kinds.items())}, flush=True)
