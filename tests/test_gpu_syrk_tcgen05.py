"""GPU: the tcgen05 / TMA 3xTF32 SYRK kernel against fp64 and against the CUDA-core kernel.
Tolerance: Kronecker factors <= 1e-4 relative (north star); 3xTF32 lands near 1e-6."""
import numpy as np
import pytest
import torch

from conftest import Golden, max_rel_err
from helpers import build_model, check_against_golden, loader_for

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from laplace_gnn_b200 import ops
    return ops


@pytest.mark.parametrize("k,n", [(16, 16), (1, 8), (100, 16), (2708, 16), (5000, 47), (19717, 64),
                                  (4096, 128), (3000, 130), (10_000, 256), (1_000_000, 256),
                                  (333_333, 47), (17, 255)])
def test_syrk_tcgen05_matches_fp64(k, n):
    ops = _ops()
    gen = torch.Generator(device=DEV).manual_seed(k + n)
    ld = (n + 3) // 4 * 4                                             # TMA: row pitch multiple of 16 bytes
    x = (torch.randn(k, ld, device=DEV, generator=gen) * 3.0 + 0.5)[:, :n]   # non-zero mean: sums do not cancel
    ref = x.double().T @ x.double()
    c = ops.syrk(x, impl="tcgen05")
    err = max_rel_err(c.cpu().numpy(), ref.cpu().numpy())
    assert err <= 2e-5, err                                           # as accurate as fp32 FMA accumulation, far inside 1e-4
    assert torch.equal(c, c.T)
    assert torch.equal(ops.syrk(x, impl="tcgen05"), c)                # deterministic reduction
    simt = ops.syrk(x, impl="simt")
    assert max_rel_err(c.cpu().numpy(), simt.cpu().numpy()) <= 4e-5   # each is ~1e-5 from fp64


def test_syrk_tcgen05_alpha_beta_and_leading_dimension():
    ops = _ops()
    x = torch.randn(5000, 64, device=DEV)
    ref = (x[:, :47].double().T @ x[:, :47].double())
    c0 = torch.ones(47, 47, device=DEV)
    c = ops.syrk(x, n=47, alpha=0.25, beta=2.0, out=c0.clone(), impl="tcgen05")   # ldx=64 > n=47
    assert max_rel_err(c.cpu().numpy(), (0.25 * ref + 2.0).cpu().numpy()) <= 2e-6
    c = ops.syrk(x, n=47, k_rows=1234, impl="tcgen05")
    ref2 = x[:1234, :47].double().T @ x[:1234, :47].double()
    assert max_rel_err(c.cpu().numpy(), ref2.cpu().numpy()) <= 2e-6


def test_syrk_tcgen05_is_3xtf32_not_plain_tf32():
    """A plain (1x) TF32 product would be off by ~1e-3 on this input; 3xTF32 is fp32-faithful.
    All-positive data is also the worst case for the tensor core's round-toward-zero accumulation:
    the segmented accumulation (192 accumulates per TMEM chain, RN adds across segments) keeps the
    drift near 1e-5 (the level of plain fp32 FMA accumulation) where one unbroken chain measured
    6e-4 at 54 k rows per CTA and would reach 1e-2 at products scale."""
    ops = _ops()
    x = (1.0 + torch.rand(200_000, 32, device=DEV) * 1e-3)            # mantissa bits below tf32 matter
    ref = x.double().T @ x.double()
    assert max_rel_err(ops.syrk(x, impl="tcgen05").cpu().numpy(), ref.cpu().numpy()) <= 2e-5
    x = torch.rand(3_000_000, 256, device=DEV)                        # long chain per CTA, n = 256
    ref = x.double().T @ x.double()
    assert max_rel_err(ops.syrk(x, impl="tcgen05").cpu().numpy(), ref.cpu().numpy()) <= 2e-5


def test_syrk_tcgen05_rejects_unaligned_pitch_and_auto_falls_back():
    ops = _ops()
    from laplace_gnn_b200._lib import LgnnError
    x = torch.randn(500, 47, device=DEV)                              # pitch 47 floats: not TMA-addressable
    with pytest.raises(LgnnError):
        ops.syrk(x, impl="tcgen05")
    ref = x.double().T @ x.double()
    assert max_rel_err(ops.syrk(x, impl="auto").cpu().numpy(), ref.cpu().numpy()) <= 1e-5


def test_kron_fit_with_tcgen05_matches_reference_goldens(golden):
    import laplace_gnn_b200 as L
    g = golden
    model = build_model(g, DEV)
    la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs={"syrk_impl": "auto"})
    la.fit(loader_for(g, DEV))
    check_against_golden(g, la.loss, la.H_facs.kfacs, la.log_marginal_likelihood())
