"""BASELINE config 5: R-MAT power-law scale sweep (a,b,c,d = .57,.19,.19,.05), kron GGN fit, SpMM GB/s
and SYRK useful TFLOP/s against the measured peaks.  Run on the GPU box:
    python tools/rmat_sweep.py [--scales 20,22 --degrees 16,64]"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--scales", default="20,22")
ap.add_argument("--degrees", default="16,64")
ap.add_argument("--features", type=int, default=128)
ap.add_argument("--classes", type=int, default=16)
ap.add_argument("--hidden", type=int, default=256)
ap.add_argument("--hub-split", action="store_true",
                help="also time the fit with unit_hub_split=True (unit-compacted slabs on graphs with hub rows)")
a = ap.parse_args()
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def rmat_edges(scale, n_edges, gen):
    src = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    dst = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    for _ in range(scale):
        r = torch.rand(n_edges, device=dev, generator=gen)
        src = (src << 1) | (r >= 0.76).long()
        dst = (dst << 1) | (((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).long()
    return torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"# peak {peak} GB/s (measured copy); F={a.features} C={a.classes} h={a.hidden} L=3; hub rows split above 16384 nnz")
print(f"{'scale':>5s} {'deg':>4s} {'nodes':>10s} {'nnz':>11s} {'maxdeg':>8s} | {'spmm d=3072':>22s} | {'spmm d=256':>20s} | {'kron fit':>22s}")
for scale in [int(v) for v in a.scales.split(",")]:
    for deg in [int(v) for v in a.degrees.split(",")]:
        n = 1 << scale
        gen = torch.Generator(device=dev).manual_seed(scale * 100 + deg)
        ei = rmat_edges(scale, n * deg // 2, gen)
        g = L.Graph.from_edge_index(ei, n, assume_undirected=True)
        del ei
        maxdeg = int(g.deg.max())
        row = f"{scale:5d} {deg:4d} {n:10d} {g.nnz:11d} {maxdeg:8d} |"
        for d in (3072, 256):
            x = torch.randn(n, d, device=dev); y = torch.empty(n, d, device=dev)
            ms = timed(lambda: ops.spmm(g.ahat, x, out=y))
            gbs = ops.spmm_algorithmic_bytes(n, g.nnz, d) / ms / 1e6
            row += f" {ms:8.2f} ms {gbs:6.0f} GB/s {gbs/peak:4.2f} |"
            del x, y
        X = torch.randn(n, a.features, device=dev, generator=gen)
        idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
        yl = torch.randint(0, a.classes, (idx.numel(),), device=dev, generator=gen)
        torch.manual_seed(0)
        model = L.SparseGCN(a.features, a.hidden, a.classes, 3, X, g).to(dev)
        def fit():
            la = L.Laplace(model, "classification", backend=L.B200GGN)
            la.fit(L.TensorBatchLoader(idx, yl))
            return la.log_marginal_likelihood()
        ms = timed(fit, reps=2)
        row += f" {ms:8.1f} ms {n/ms*1e3/1e6:6.2f} Mnodes/s"
        if a.hub_split:
            ml_dense = float(fit())
            def fit_split():
                la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs={"unit_hub_split": True})
                la.fit(L.TensorBatchLoader(idx, yl))
                return la.log_marginal_likelihood(), la.backend.last_stats["unit_slabs"]
            ms2 = timed(fit_split, reps=2)
            ml_split, n_units = fit_split()
            row += (f" | hub split {ms2:8.1f} ms {n/ms2*1e3/1e6:6.2f} Mnodes/s, unit SpMMs {n_units}, "
                    f"marglik rel diff {abs(float(ml_split) - ml_dense) / abs(ml_dense):.1e}")
        print(row, flush=True)
        del model, X, g
        torch.cuda.empty_cache()
