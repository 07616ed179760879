"""Bring-up probe for the tcgen05 SYRK kernel: runs a few shapes in a subprocess each (a device
trap poisons the CUDA context) and prints the error against fp64.  Usage: python tools/syrk_bringup.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r"""
import sys, torch
sys.path.insert(0, %r)
from laplace_gnn_b200 import ops
k, n = %d, %d
x = torch.randn(k, n, device='cuda') + 0.5
ref = x.double().T @ x.double()
c = ops.syrk(x, impl='tcgen05')
torch.cuda.synchronize()
err = float((c.double() - ref).abs().max() / ref.abs().max())
print('RESULT k=%%d n=%%d rel_err=%%.3e' %% (k, n, err))
if err > 1e-4:
    d = (c.double() - ref).abs() / ref.abs().max()
    print(' worst idx', int(d.argmax()) // n, int(d.argmax()) %% n, ' c[0,:4]', c[0, :4].tolist(), ' ref[0,:4]', ref[0, :4].tolist())
"""
for swap in ("0", "1"):
    for k, n in [(64, 16), (1000, 64), (5000, 256), (100000, 47)]:
        env = dict(os.environ, LGNN_SYRK_SWAP_LBO_SBO=swap)
        try:
            out = subprocess.run([sys.executable, "-c", CASE % (ROOT, k, n)], env=env, capture_output=True,
                                 text=True, timeout=120)
            tail = (out.stdout.strip().splitlines() or ["<no stdout>"])
            print(f"[swap={swap}] rc={out.returncode}", " | ".join(tail[-2:]), flush=True)
            if out.returncode != 0:
                print("   stderr:", out.stderr.strip().splitlines()[-1:] , flush=True)
        except subprocess.TimeoutExpired:
            print(f"[swap={swap}] k={k} n={n} TIMEOUT", flush=True)
