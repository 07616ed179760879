"""How long do the small factor eigendecompositions take on the device (torch.linalg.eigh variants)?"""
import torch, time
dev = torch.device("cuda:0")
def psd(n):
    a = torch.randn(n, 4 * n, device=dev)
    return a @ a.T / (4 * n)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for n in (47, 100, 256, 500, 1433):
    m = psd(n)
    m64 = m.double()
    print(f"n={n:5d}  fp32 {timed(lambda: torch.linalg.eigh(m, UPLO='U')):8.2f} ms   fp64 {timed(lambda: torch.linalg.eigh(m64, UPLO='U')):8.2f} ms", flush=True)
m4 = torch.stack([psd(256) for _ in range(4)])
print(f"batched 4x256 fp32 {timed(lambda: torch.linalg.eigh(m4, UPLO='U')):8.2f} ms   fp64 {timed(lambda: torch.linalg.eigh(m4.double(), UPLO='U')):8.2f} ms")
try:
    torch.backends.cuda.preferred_linalg_library("magma")
    m = psd(256)
    print(f"magma 256 fp32 {timed(lambda: torch.linalg.eigh(m, UPLO='U')):8.2f} ms")
except Exception as e:
    print("magma unavailable:", e)
