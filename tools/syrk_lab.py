"""SYRK microbenchmark (run on the GPU box): accuracy vs fp64 and useful TFLOP/s of the tcgen05 3xTF32
kernel and the CUDA-core kernel at hot-path shapes.  LGNN_SYRK_SEG_STEPS selects the TMEM segment
length.    python tools/syrk_lab.py [--k 8000000 --n 256,48,100]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laplace_gnn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=8_000_000)
ap.add_argument("--n", default="256,48,100,128")
ap.add_argument("--impl", default="tcgen05,simt")
ap.add_argument("--dist", default="relu,randn")
a = ap.parse_args()
dev = torch.device("cuda:0")
for n in [int(v) for v in a.n.split(",")]:
    ld = (n + 3) // 4 * 4
    for dist in a.dist.split(","):
        x = torch.randn(a.k, ld, device=dev)
        if dist == "relu":
            x = torch.relu(x)
        xs = x[:, :n]
        ref = torch.zeros(n, n, device=dev, dtype=torch.float64)
        for s in range(0, a.k, 1 << 20):
            blk = xs[s:s + (1 << 20)].double()
            ref += blk.T @ blk
        for impl in a.impl.split(","):
            c = ops.syrk(xs, impl=impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.syrk(xs, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            err = float((c.double() - ref).abs().max() / ref.abs().max())
            fl = a.k * n * (n + 1)
            print(f"n={n:4d} {dist:5s} impl={impl:8s} seg={os.environ.get('LGNN_SYRK_SEG_STEPS','dflt'):>5s} "
                  f"{ms:8.3f} ms  {fl/ms/1e9:8.1f} useful TFLOP/s  {a.k*ld*4/ms/1e6:7.0f} GB/s  rel_err={err:.2e}", flush=True)
        del x, xs
