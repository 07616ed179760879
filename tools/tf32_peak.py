"""Measured TF32 tensor-core GEMM rate of this B200 (cuBLAS through torch.matmul, allow_tf32), the way
MEASURED_PEAKS.json measures bf16: 8192^3, best of 10 (burst) and back to back for 3 s (sustained).  The SYRK / fused
GEMM roofline fractions of bench.py use half the measured bf16 figure until this file exists:
    python tools/tf32_peak.py > gpurun_out/tf32_peak.json     (then copy to profiles/tf32_peak.json)"""
import json, time
import torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = float("inf")
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps, t0 = 0, time.time()
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(20):
        torch.matmul(a, b, out=c)
    reps += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
flops = 2.0 * n ** 3
print(json.dumps({"tf32_tflops": flops / best / 1e9, "tf32_tflops_sustained": flops * reps / e0.elapsed_time(e1) / 1e9,
                  "how": "torch.matmul fp32 inputs, allow_tf32, 8192^3; best of 10 and 3 s back to back",
                  "gpu": torch.cuda.get_device_name(0)}))
