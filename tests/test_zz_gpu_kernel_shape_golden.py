"""GPU: the whole fit on the shapes the headline kernels take (3 layers, h = 256, C = 40: fused tcgen05 GEMM +
relu' mask, tcgen05 SYRK at n = 256, unit-compacted slabs in column groups 16 + 16 + 8) against the REFERENCE's own
factors, loss and log marginal likelihood (tests/golden/arxiv_mini_3l.npz and products_mini_3l.npz — 47 classes, logits
pitch padded to 48, groups 16 + 16 + 15(+1) — oracle/make_golden.py tier O2).
North-star tolerances: factors <= 1e-4 rel, marglik <= 1e-3 rel.  (File name: sorts after the per-kernel tests.)"""
import pytest
import torch

from conftest import GOLDEN_KERNEL_SHAPES, GOLDEN_SMALL, GgnGolden, Golden, max_rel_err
from helpers import build_model, check_against_golden, loader_for

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.mark.parametrize("kw", [
    {},                                                        # the default path of bench.py
    {"unit_slabs": False},                                     # dense slabs, fused GEMM with the mask epilogue
    {"fused_gemm": False, "syrk_impl": "simt"},                # cuBLAS + mask kernel, CUDA-core SYRK
    {"hess_sqrt": "reference", "rhs_tile_bytes": 2 * 1200 * 256 * 4 * 12},   # column groups of 12
], ids=["default", "dense-slabs", "two-step-simt", "groups-of-12"])
@pytest.mark.parametrize("name", GOLDEN_KERNEL_SHAPES)
def test_fit_matches_the_reference_at_kernel_shapes(name, kw):
    import laplace_gnn_b200 as L
    g = Golden(name)
    if "rhs_tile_bytes" in kw:
        kw = dict(kw, rhs_tile_bytes=2 * g.n * 256 * 4 * 12)
    model = build_model(g, DEV)
    la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs=kw)
    la.fit(loader_for(g, DEV))
    ml = la.log_marginal_likelihood()
    check_against_golden(g, la.loss, la.H_facs.kfacs, ml)
    st = la.backend.last_stats
    assert (st["unit_slabs"] > 0) == kw.get("unit_slabs", True)
    if "rhs_tile_bytes" in kw:
        # a budget of 12 columns, equal groups: 40 = 4 x 10, 47 = 12 + 12 + 12 + 11(+1)
        assert st["group"] == (10 if g.C == 40 else 12) and st["n_groups"] == 4


@pytest.mark.parametrize("name", GOLDEN_KERNEL_SHAPES)
def test_logits_match_the_reference_at_kernel_shapes(name):
    g = Golden(name)
    model = build_model(g, DEV)
    model.eval()
    with torch.no_grad():
        out = model(torch.from_numpy(g.idx).to(DEV))
    # three GEMM + SpMM layers deep: the per-kernel 1e-5 (tests/test_gpu_parity.py) compounds
    assert max_rel_err(out.cpu().numpy(), g.z["logits"]) <= 5e-5


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_ggn_mode_matches_upstream_curvlinops_goldens(name):
    """hess_sqrt="ggn" on the device against the reference's classes run with upstream curvlinops' detach restored
    (oracle/make_golden_ggn.py): factors <= 1e-4 rel, marglik <= 1e-3 rel."""
    import laplace_gnn_b200 as L
    g, gg = Golden(name), GgnGolden(name)
    la = L.Laplace(build_model(g, DEV), "classification", backend=L.B200GGN, backend_kwargs={"hess_sqrt": "ggn"})
    la.fit(loader_for(g, DEV))
    ml = la.log_marginal_likelihood()
    for blk, ref_blk in zip(la.H_facs.kfacs, gg.kfacs):
        for h, ref in zip(blk, ref_blk):
            assert max_rel_err(h.cpu().numpy(), ref) <= 1e-4
    assert abs(float(la.loss) - gg.loss) <= 1e-4 * abs(gg.loss)
    assert abs(float(ml) - gg.marglik) <= 1e-3 * abs(gg.marglik)
