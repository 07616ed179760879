"""Where does a products-shaped step go besides our kernels?  Runs two fits under torch.profiler (kineto)
on the GPU box and prints (a) GPU time per kernel name, (b) the largest idle gaps of the device inside the
step with the kernel that follows them.  usage: step_timeline.py [scale]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from torch.profiler import profile, ProfilerActivity
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda:0")
n, u, f, c, h, l = int(2_449_029 * scale), int(61_859_140 * scale), 100, 47, 256, 3
gen = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, n, (u,), device=dev, generator=gen); dst = torch.randint(0, n, (u,), device=dev, generator=gen)
graph = L.Graph.from_edge_index(torch.stack([torch.cat([src, dst]), torch.cat([dst, src])]), n, assume_undirected=True)
del src, dst
X = torch.randn(n, f, device=dev, generator=gen)
idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
y = torch.randint(0, c, (idx.numel(),), device=dev, generator=gen)
torch.manual_seed(0)
model = L.SparseGCN(f, h, c, l, X, graph).to(dev)
loader = L.TensorBatchLoader(idx, y)
def step():
    la = L.Laplace(model, "classification", backend=L.B200GGN)
    la.fit(loader)
    return la.log_marginal_likelihood()
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
print(f"device span {1e-3 * (t1 - t0):.1f} ms, {len(evs)} device activities")
agg = {}
for e in evs:
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
busy = sum(a[1] for a in agg.values())
print(f"busy {1e-3 * busy:.1f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{1e-3 * a[1]:10.2f} ms {a[0]:5d}  {k}")
gaps, end = [], evs[0].time_range.end
for e in evs[1:]:
    if e.time_range.start > end:
        gaps.append((e.time_range.start - end, e.name[:60], 1e-3 * (e.time_range.start - t0)))
    end = max(end, e.time_range.end)
print(f"idle {1e-3 * sum(g[0] for g in gaps):.1f} ms in {len(gaps)} gaps; largest:")
for g in sorted(gaps, reverse=True)[:15]:
    print(f"{1e-3 * g[0]:10.2f} ms before {g[1]}  (at {g[2]:.0f} ms)")
