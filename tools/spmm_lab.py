"""SpMM kernel microbenchmark on a products-shaped random graph (run on the GPU box):
    python tools/spmm_lab.py [--n 2449029 --pairs 61859140 --d 3072,960,256 --impl ldg,bulk]
Prints achieved algorithmic GB/s per (impl, d) against MEASURED_PEAKS.json."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_449_029)
ap.add_argument("--pairs", type=int, default=61_859_140)
ap.add_argument("--d", default="3072,4096,960,256")
ap.add_argument("--impl", default="ldg,bulk")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, a.n, (a.pairs,), device=dev, generator=gen)
dst = torch.randint(0, a.n, (a.pairs,), device=dev, generator=gen)
ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
g = L.Graph.from_edge_index(ei, a.n, assume_undirected=True)
del ei, src, dst
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(f"n={a.n} nnz={g.nnz} peak={peak} GB/s", flush=True)
for d in [int(v) for v in a.d.split(",")]:
    x = torch.randn(a.n, d, device=dev)
    y = torch.empty(a.n, d, device=dev)
    for impl in a.impl.split(","):
        if impl == "bulk" and d < 512:
            continue
        ops.spmm(g.ahat, x, out=y, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ops.spmm(g.ahat, x, out=y, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        b = ops.spmm_algorithmic_bytes(a.n, g.nnz, d)
        print(f"d={d:6d} impl={impl:5s} {ms:9.3f} ms  {b/ms/1e6:8.1f} GB/s  frac={b/ms/1e6/peak:.3f}", flush=True)
    del x, y
