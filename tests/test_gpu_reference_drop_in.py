"""GPU: the UNMODIFIED reference `laplace` package (its own `ParametricLaplace.fit`, `KronLaplace._curv_closure`
at laplace/baselaplace.py:1565-1571, `Kron` / `KronDecomposed`, `log_marginal_likelihood`,
`optimize_prior_precision`) drives `B200GGN` on the device through `backend=` and reproduces the goldens its own
`CurvlinopsGGN` produced on the CPU — the drop-in claim of SURVEY §8(b), executed on a B200.

The reference tree is /root/reference in the dev container and its staged copy under the git-ignored
baseline/_ref/ on the GPU box (oracle/stage_reference.py, run by __graft_entry__.build()); without either the test
skips."""
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _reference_available():
    from oracle import ref_loader
    return ref_loader.available()


@pytest.mark.skipif(not _reference_available(), reason="no reference tree (neither /root/reference nor baseline/_ref)")
def test_reference_kronlaplace_drives_b200ggn_on_the_device():
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ref_loader
R = ref_loader.load()
import laplace, laplace_gnn_b200 as L
assert laplace.__file__.startswith(ref_loader.REFERENCE_ROOT)
from laplace_gnn_b200 import _lib
from conftest import Golden
from helpers import build_model, loader_for, check_against_golden
assert issubclass(L.B200GGN, R.GGNInterface)          # built on the reference's own interface class
dev = 'cuda:0'
for name in ['tiny_directed_3l', 'tiny_symmetrised_2l', 'small_multibatch_2l', 'cora_shape', 'pubmed_shape',
             'arxiv_mini_3l', 'products_mini_3l']:
    g = Golden(name)
    model = build_model(g, dev)
    n0 = _lib.launch_count()
    la = R.Laplace(model, 'classification', subset_of_weights='all', hessian_structure='kron', backend=L.B200GGN)
    assert type(la).__module__.startswith('laplace.')  # the reference's KronLaplace, not the stand-in
    la.fit(loader_for(g, dev))
    assert isinstance(la.H_facs, R.Kron) and la.H_facs.kfacs[0][0].is_cuda
    assert _lib.launch_count() > n0                     # the factors came from liblgnn kernels
    ml = la.log_marginal_likelihood()
    check_against_golden(g, la.loss, la.H_facs.kfacs, ml)
    if g.n <= 3000:
        la.optimize_prior_precision(n_steps=5)          # detached factors: the reference's tuning loop works
        assert torch.isfinite(la.log_marginal_likelihood())
print('DROPIN_GPU_OK')
""" % (ROOT, ROOT)
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, timeout=900)
    assert "DROPIN_GPU_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
