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

# ---- which stage paces the tcgen05 SYRK?  The lab copy of the kernel (csrc/syrk_tcgen05_lab.cu) with single stages
# switched off.  Results of ablated launches are wrong by construction; only their times are read.
import ctypes as C
from laplace_gnn_b200 import _lib
lib = _lib.load()
i64, vp = C.c_int64, C.c_void_p
lib.lgnn_syrk_lab_f32.restype = C.c_int
lib.lgnn_syrk_lab_f32.argtypes = [vp, i64, i64, i64, vp, i64, vp, vp]
lib.lgnn_syrk_lab_workspace_bytes.restype = C.c_size_t
lib.lgnn_syrk_lab_workspace_bytes.argtypes = [i64, i64]
lib.lgnn_syrk_lab_set_ablate.restype = C.c_int
lib.lgnn_syrk_lab_set_ablate.argtypes = [C.c_int]
n = 256
x = torch.randn(a.k, n, device=dev)
c = torch.empty(n, n, device=dev)
ws = torch.empty(lib.lgnn_syrk_lab_workspace_bytes(a.k, n), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
print(f"ablation, n={n} k={a.k}:")
for bits, what in [(0, "everything on (lab copy)"), (2, "only hi.hi issued (1 of 3 MMAs)"), (1, "no MMA issued"),
                   (4, "no hi / lo stores by the transform warps"), (8, "no red.global flush"),
                   (1 | 4, "no MMA, no transform stores"), (1 | 4 | 8, "TMA + barriers + TMEM drain only")]:
    lib.lgnn_syrk_lab_set_ablate(bits)
    def run():
        rc = lib.lgnn_syrk_lab_f32(x.data_ptr(), x.stride(0), a.k, n, c.data_ptr(), c.stride(0), ws.data_ptr(), st)
        assert rc == 0, lib.lgnn_last_error()
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"   {what:42s} {ms:8.3f} ms   {a.k * n * (n + 1) / ms / 1e9:7.1f} useful TFLOP/s-equivalent", flush=True)
lib.lgnn_syrk_lab_set_ablate(0)
