"""Zero-compressed SpMM microbenchmark on the products-shaped graph (run on the GPU box)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops
dev = torch.device("cuda:0")
n, pairs = 2_449_029, 61_859_140
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
for d in (3840, 4096, 3072):
    x = torch.randn(n, d, device=dev)
    x *= (torch.rand(n, d, device=dev) < 0.5)
    y = torch.empty(n, d, device=dev)
    t_dense = timed(lambda: ops.spmm(g.ahat, x, out=y))
    buf = torch.empty(n * ops.pack_rows_pitch(d), dtype=torch.uint8, device=dev)
    t_pack = timed(lambda: ops.pack_rows(x, d, out=buf))
    pr = ops.pack_rows(x, d, out=buf)
    t_packed = timed(lambda: ops.spmm_packed(g.ahat, pr, out=y))
    live = int(pr.len.to(torch.int64)[g.ahat.col.long()].sum())
    by = g.nnz * 12 + live + n * d * 4
    print(f"d={d}: dense {t_dense:8.2f} ms | pack {t_pack:6.2f} ms | packed spmm {t_packed:8.2f} ms  "
          f"({by / t_packed / 1e6:6.0f} GB/s of live bytes, {by / t_packed / 1e6 / peak:4.2f} of peak)  speed-up incl. pack {t_dense / (t_packed + t_pack):4.2f}x", flush=True)
    del x, y, buf, pr
