"""Unit-compacted SpMM microbenchmark on the products-shaped graph (run on the GPU box).
usage: units_lab.py [g ...]   (default 12 16)"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops
dev = torch.device("cuda:0")
n, pairs = 2_449_029, 61_859_140
scale = float(os.environ.get("LAB_SCALE", "1"))
n, pairs = int(n * scale), int(pairs * scale)
gen = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, n, (pairs,), device=dev, generator=gen); dst = torch.randint(0, n, (pairs,), device=dev, generator=gen)
g = L.Graph.from_edge_index(torch.stack([torch.cat([src, dst]), torch.cat([dst, src])]), n, assume_undirected=True)
del src, dst
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
h = 256
variants = {12: (1, 8, 9, 10, 11, 12, 13, 14, 15), 16: (1, 8, 9, 10, 12, 13, 14), 8: (1, 8, 12), 4: (1, 8, 12),
            6: (8, 9, 12, 24), 10: (8, 9, 12), 14: (8, 9, 12), 2: (8, 12)}   # g % 4 == 2: spmm_units_even.cu   # 1: simple kernel, 8+: staged
print(f"n={n} nnz={g.nnz} peak={peak} GB/s", flush=True)
for gg in [int(a) for a in sys.argv[1:]] or [12, 16]:
    d = gg * h
    act = torch.randn(n, h, device=dev)                       # ~50 % live units, like relu of a centred layer
    x = torch.randn(n, d, device=dev)
    x.view(n, gg, h).mul_((act > 0)[:, None, :])
    y = torch.empty(n, d, device=dev)
    t_dense = timed(lambda: ops.spmm(g.ahat, x, out=y))
    ref = y[:200_000].clone()
    dense_bytes = ops.spmm_algorithmic_bytes(n, g.nnz, d)
    hdr = torch.empty(n, h // 32, 2, dtype=torch.int32, device=dev)
    x2 = x.clone()
    t_pack = timed(lambda: ops.unit_pack(x2.copy_(x), act, gg, hdr=hdr)) - timed(lambda: x2.copy_(x))
    del x2
    us = ops.unit_pack(x, act, gg, hdr=hdr)
    live = int(us.live_units()[g.ahat.col.long()].sum())
    by = g.nnz * (8 + 8 * (h // 32)) + (n + 1) * 8 + live * gg * 4 + n * d * 4
    print(f"g={gg} d={d}: dense {t_dense:8.2f} ms ({dense_bytes / t_dense / 1e6:6.0f} GB/s, {dense_bytes / t_dense / 1e6 / peak:4.2f})"
          f" | unit_pack {t_pack:6.2f} ms | live fraction {live * gg * 4 / (g.nnz * d * 4):.3f}", flush=True)
    for v in variants.get(gg, (0,)):
        t = timed(lambda: ops.spmm_units(g.ahat, us, out=y, variant=v))
        same = bool(torch.equal(y[:200_000], ref))
        print(f"    variant {v}: {t:8.2f} ms  {by / t / 1e6:6.0f} GB/s of live bytes ({by / t / 1e6 / peak:4.2f} of peak)"
              f"  dense-equivalent {dense_bytes / t / 1e6:6.0f} GB/s  speed-up {t_dense / t:4.2f}x (incl. pack {t_dense / (t + t_pack):4.2f}x)  bit-equal={same}", flush=True)
    del x, y, act, us, hdr, ref
