"""BASELINE config 5: R-MAT power-law scale sweep (a, b, c, d = .57, .19, .19, .05): kron GGN fit against the
node-factorised diagonal GGN, SpMM GB/s and SYRK useful TFLOP/s per kernel group against the measured peaks.
Run on the GPU box:
    python tools/rmat_sweep.py [--scales 20,22,24 --degrees 16,64] > profiles/r2_rmat_sweep.txt"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laplace_gnn_b200 as L
from laplace_gnn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--scales", default="20,22")
ap.add_argument("--degrees", default="16,64")
ap.add_argument("--features", type=int, default=128)
ap.add_argument("--classes", type=int, default=16)
ap.add_argument("--hidden", type=int, default=256)
ap.add_argument("--no-diag", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6650.0))
try:
    tf32 = float(json.load(open(os.path.join(ROOT, "profiles", "tf32_peak.json")))["tf32_tflops_sustained"])
except Exception:
    tf32 = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2


def rmat_edges(scale, n_edges, gen):
    src = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    dst = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    for _ in range(scale):
        r = torch.rand(n_edges, device=dev, generator=gen)
        src = (src << 1) | (r >= 0.76).long()
        dst = (dst << 1) | (((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).long()
    return torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def profiled(fn):
    """One call with per-launch records -> {kind: (ms, algorithmic bytes or useful flops)}."""
    fn(); torch.cuda.synchronize()
    ops.profile_begin()
    fn()
    torch.cuda.synchronize()
    recs = ops.PROFILE
    out = {}
    for r in recs:
        r["ms"] = r["start"].elapsed_time(r["end"])
    for r in ops.profile_end():
        k = out.setdefault(r["kind"], [0.0, 0.0])
        k[0] += r["ms"]; k[1] += r["bytes"]
    return out


print(f"# R-MAT sweep, one B200.  HBM copy peak {peak:.0f} GB/s, TF32 GEMM peak {tf32:.0f} TFLOP/s (sustained, measured); "
      f"F={a.features} C={a.classes} h={a.hidden} L=3; hub rows (> 4096 non-zeros) cut into pieces for the unit SpMM and the narrow forward SpMM")
for scale in [int(v) for v in a.scales.split(",")]:
    for deg in [int(v) for v in a.degrees.split(",")]:
        n = 1 << scale
        gen = torch.Generator(device=dev).manual_seed(scale * 100 + deg)
        ei = rmat_edges(scale, n * deg // 2, gen)
        g = L.Graph.from_edge_index(ei, n, assume_undirected=True)
        del ei
        maxdeg = int(g.deg.max())
        print(f"scale {scale} avg degree {deg}: nodes {n} nnz {g.nnz} max row {maxdeg} pieces of hub rows {g.extra_rows()}", flush=True)
        x = torch.randn(n, 256, device=dev)
        ms_plain = timed(lambda: ops.spmm(g.ahat, x))
        buf = torch.empty(n + g.extra_rows(), 256, device=dev)
        ms_split = timed(lambda: g.propagate(x, out=buf))
        by = ops.spmm_algorithmic_bytes(n, g.nnz, 256)
        print(f"   forward SpMM d=256: warp per row {ms_plain:8.2f} ms {by / ms_plain / 1e6:6.0f} GB/s ({by / ms_plain / 1e6 / peak:4.2f})"
              f" | hub rows in pieces {ms_split:8.2f} ms {by / ms_split / 1e6:6.0f} GB/s ({by / ms_split / 1e6 / peak:4.2f})", flush=True)
        del x, buf
        X = torch.randn(n, a.features, device=dev, generator=gen)
        idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
        yl = torch.randint(0, a.classes, (idx.numel(),), device=dev, generator=gen)
        torch.manual_seed(0)
        model = L.SparseGCN(a.features, a.hidden, a.classes, 3, X, g).to(dev)

        def fit(structure="kron", **kw):
            la = L.Laplace(model, "classification", hessian_structure=structure, backend=L.B200GGN, backend_kwargs=kw)
            la.fit(L.TensorBatchLoader(idx, yl))
            return la, la.log_marginal_likelihood()

        res = {}
        for tag, kw in (("dense slabs", {"unit_slabs": False}), ("unit slabs + hub split", {"unit_hub_split": True})):
            try:
                ms = timed(lambda: fit(**kw), reps=2)
                la, ml = fit(**kw)
                kinds = profiled(lambda: fit(**kw))
                res[tag] = float(ml)
                sp_ms = sum(v[0] for k, v in kinds.items() if k.startswith("spmm"))
                sp_by = sum(v[1] for k, v in kinds.items() if k.startswith("spmm"))
                sy = kinds.get("syrk", [0.0, 0.0]); gm = kinds.get("gemm_mask", [0.0, 0.0])
                print(f"   kron, {tag:24s} {ms:9.1f} ms {n / ms * 1e3 / 1e6:6.2f} Mnodes/s  marglik {float(ml):.1f}  group {la.backend.last_stats['group']} x {la.backend.last_stats['n_groups']}"
                      f" | SpMM {sp_ms:8.1f} ms {sp_by / max(sp_ms, 1e-9) / 1e6:6.0f} GB/s ({sp_by / max(sp_ms, 1e-9) / 1e6 / peak:4.2f} of the copy peak)"
                      f" | SYRK {sy[0]:7.1f} ms {sy[1] / max(sy[0], 1e-9) / 1e9:6.1f} useful TFLOP/s ({sy[1] / max(sy[0], 1e-9) / 1e9 / tf32:4.2f} of TF32, x2.25-3 issued)"
                      f" | fused GEMM {gm[0]:7.1f} ms {gm[1] / max(gm[0], 1e-9) / 1e9:6.1f} TFLOP/s"
                      f" | by kind {{{', '.join(f'{k}: {v[0]:.1f}' for k, v in sorted(kinds.items()))}}}", flush=True)
                del la
            except torch.OutOfMemoryError as e:
                print(f"   kron, {tag}: out of memory ({str(e)[:80]})", flush=True)
                from laplace_gnn_b200 import curvature
                curvature.release_workspace(); torch.cuda.empty_cache()
        if len(res) == 2:
            v = list(res.values())
            print(f"   marglik of the two layouts: rel diff {abs(v[0] - v[1]) / abs(v[0]):.1e}", flush=True)
        if not a.no_diag:
            try:
                ms = timed(lambda: fit("diag", diag_mode="node_factorised", unit_hub_split=True), reps=2)
                _, ml = fit("diag", diag_mode="node_factorised", unit_hub_split=True)
                print(f"   diag (node-factorised)          {ms:9.1f} ms {n / ms * 1e3 / 1e6:6.2f} Mnodes/s  marglik {float(ml):.1f}", flush=True)
            except torch.OutOfMemoryError as e:
                print(f"   diag: out of memory ({str(e)[:80]})", flush=True)
        del model, X, g
        from laplace_gnn_b200 import curvature
        curvature.release_workspace()
        torch.cuda.empty_cache()
