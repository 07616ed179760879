"""Fused (A W) ⊙ mask kernel microbenchmark (run on the GPU box).
    python tools/gemm_lab.py [--m 20000000]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laplace_gnn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=19_999_992)  # multiple of group 12
a = ap.parse_args()
dev = torch.device("cuda:0")
for k, n, group, masked in [(256, 256, 12, True), (256, 256, 12, False), (47, 256, 12, True), (256, 128, 12, True), (64, 64, 12, True)]:
    ld = (k + 3) // 4 * 4
    x = torch.randn(a.m, ld, device=dev)[:, :k]
    w = torch.randn(k, n, device=dev) / k ** 0.5
    act = torch.randn((a.m + group - 1) // group, n, device=dev) if masked else None
    out = torch.empty(a.m, n, device=dev)
    wp = ops.gemm_mask_prepare(w)
    ops.gemm_mask(x, wp, act, group, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.gemm_mask(x, wp, act, group, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = 2.0 * a.m * k * n
    by = a.m * (ld + n) * 4 + (act.numel() * 4 if masked else 0)
    # two-step baseline: cuBLAS fp32 + mask kernel
    e0.record()
    ref = torch.mm(x, w, out=out)
    if masked:
        ops.relu_mask_mul(out, act, group)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1)
    print(f"k={k:3d} n={n:3d} mask={int(masked)}  fused {ms:8.3f} ms  {fl/ms/1e9:7.1f} useful TFLOP/s  {by/ms/1e6:7.0f} GB/s | "
          f"cuBLAS fp32 + mask {ms2:8.3f} ms", flush=True)
    del x, out, act

# ---- which stage paces a tile?  The lab copy of the kernel (csrc/gemm_mask_lab.cu) with single stages switched off.
# Results of ablated launches are wrong by construction; only their times are read.
import ctypes as C
from laplace_gnn_b200 import _lib
lib = _lib.load()
i64, vp = C.c_int64, C.c_void_p
lib.lgnn_gemm_mask_lab_f32.restype = C.c_int
lib.lgnn_gemm_mask_lab_f32.argtypes = [vp, i64, i64, i64, vp, vp, i64, vp, i64, C.c_int32, vp, i64, vp]
lib.lgnn_gemm_mask_lab_set_ablate.restype = C.c_int
lib.lgnn_gemm_mask_lab_set_ablate.argtypes = [C.c_int]
k, n, group = 256, 256, 16
m = a.m // group * group
x = torch.randn(m, k, device=dev)
w = torch.randn(k, n, device=dev) / k ** 0.5
out = torch.empty(m, n, device=dev)
wp = ops.gemm_mask_prepare(w)
st = torch.cuda.current_stream().cuda_stream
print(f"ablation, k={k} n={n} m={m}, no mask:")
for bits, what in [(0, "everything on (lab copy)"), (2, "only hi.hi issued (1 of 3 MMAs)"), (1, "no MMA issued"),
                   (4, "no tcgen05.st of the A operand"), (8, "no global stores"), (1 | 4, "no MMA, no tcgen05.st"),
                   (1 | 4 | 8, "TMA + barriers + TMEM drain only")]:
    lib.lgnn_gemm_mask_lab_set_ablate(bits)
    def run():
        rc = lib.lgnn_gemm_mask_lab_f32(x.data_ptr(), x.stride(0), m, k, wp.wt_hi.data_ptr(), wp.wt_lo.data_ptr(), n,
                                        None, 0, group, out.data_ptr(), out.stride(0), st)
        assert rc == 0, lib.lgnn_last_error()
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    tiles = (m + 127) // 128
    print(f"   {what:42s} {ms:8.3f} ms   {ms * 1e3 / (tiles / 37):6.2f} us per tile and cluster", flush=True)
lib.lgnn_gemm_mask_lab_set_ablate(0)
