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
