import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_SMALL = ["tiny_undirected_2l", "tiny_directed_3l", "tiny_directed_dups_2l",
                "tiny_symmetrised_2l", "small_multibatch_2l"]
GOLDEN_ALL = GOLDEN_SMALL + ["cora_shape", "pubmed_shape"]
# the headline kernels' shapes (3 layers, h = 256, C = 40) at a size the reference finishes in seconds
GOLDEN_KERNEL_SHAPES = ["arxiv_mini_3l", "products_mini_3l"]


# kernels that have not yet run on a B200 keep their GPU tests behind LGNN_LAB=1, so that the default
# `-m gpu` run covers exactly what the package uses by default
LAB = os.environ.get("LGNN_LAB") == "1"
lab_only = pytest.mark.skipif(not LAB, reason="lab kernel, not on the default path: set LGNN_LAB=1 to run")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """One tests/golden/*.npz fixture (inputs + the reference's outputs, oracle/make_golden.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        self.n = int(z["n"])
        self.F = int(z["F"])
        self.symmetric = bool(z["symmetric"])
        self.edge_index = z["edge_index"].astype(np.int64)
        self.x = z["x"] if z["x"].size else np.unpackbits(z["x_bits"], axis=1)[:, : self.F].astype(np.float32)
        self.idx = z["idx"].astype(np.int64)
        self.y = z["y"].astype(np.int64)
        self.L = sum(1 for k in z.files if k.startswith("W"))
        self.Ws = [z[f"W{l}"] for l in range(self.L)]
        self.bs = [z[f"b{l}"] for l in range(self.L)]
        self.batch_size = int(z["batch_size"])
        self.loss = float(z["loss"])
        self.marglik = float(z["marglik"])
        self.kfacs = []
        for b in range(int(z["n_blocks"])):
            blk, j = [], 0
            while f"kfac_{b}_{j}" in z.files:
                blk.append(z[f"kfac_{b}_{j}"])
                j += 1
            self.kfacs.append(blk)
        self.C = self.Ws[-1].shape[0]
        self.h = self.Ws[0].shape[0]


class GgnGolden:
    """tests/golden/ggn_<name>.npz: the reference's classes with upstream curvlinops' ``.detach()`` on the
    Hessian-sqrt input restored at run time (oracle/make_golden_ggn.py) — what pins ``hess_sqrt="ggn"``.
    Inputs and weights are those of <name>.npz."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, f"ggn_{name}.npz"))
        self.name = name
        self.loss, self.marglik = float(z["loss"]), float(z["marglik"])
        self.fork_marglik = float(z["fork_marglik"])
        self.kfacs = []
        for b in range(int(z["n_blocks"])):
            blk, j = [], 0
            while f"kfac_{b}_{j}" in z.files:
                blk.append(z[f"kfac_{b}_{j}"])
                j += 1
            self.kfacs.append(blk)


@pytest.fixture(params=GOLDEN_SMALL)
def golden_small(request):
    return Golden(request.param)


@pytest.fixture(params=GOLDEN_ALL)
def golden(request):
    return Golden(request.param)


@pytest.fixture
def fake_ops(monkeypatch):
    """Replace laplace_gnn_b200.ops kernels by the oracle-backed CPU test double."""
    import fake_ops as F
    import laplace_gnn_b200.ops as ops
    for name in F.ALL:
        monkeypatch.setattr(ops, name, getattr(F, name))
    return F


def max_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
