"""TEST INFRASTRUCTURE — loads the *unmodified* reference from /root/reference (dev container) or from
its staged copy under the git-ignored baseline/_ref/ (GPU box; oracle/stage_reference.py).

It is used by ``oracle/make_golden.py`` to generate the committed golden
fixtures under ``tests/golden/`` and by CPU tests that pin the numpy oracle to
the reference's own classes.  Nothing in the product package imports this.

The reference cannot be imported as-is because optional third-party packages
are missing here (torchmetrics, opt_einsum, backpack, asdl, einconv,
torch_geometric, ipdb, matplotlib, GPUtil).  None of them is on the GCN+KFAC
hot path, so we fabricate empty stand-in modules for those
roots with a ``sys.meta_path`` finder.  Four stand-ins need real behaviour:

* ``opt_einsum.contract``        -> ``torch.einsum``
* ``torchmetrics.Metric``        -> ``torch.nn.Module``
* ``torch_geometric.nn.resolver.activation_resolver("relu")`` -> ``nn.ReLU()``
* ``ipdb.set_trace``             -> no-op  (reference: gnn/models/base_gnn.py:108-109)
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """/root/reference in the dev container; on the GPU box the copy oracle/stage_reference.py put under the
    git-ignored baseline/_ref/ (it travels with the gpurun snapshot)."""
    env = os.environ.get("LGNN_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "laplace")) and os.path.isdir(os.path.join(cand, "curvlinops")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_STUB_ROOTS = (
    "torchmetrics",
    "opt_einsum",
    "backpack",
    "asdl",
    "einconv",
    "torch_geometric",
    "ipdb",
    "matplotlib",
    "GPUtil",
    "seaborn",
    "networkx",
)


class _Dummy:
    """Attribute sink: any attribute / call / subscript yields another dummy."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, k):
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        # classes, so that `class X(stub.Base)` and isinstance checks work
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        import torch

        name = module.__name__
        if name == "opt_einsum":
            module.contract = torch.einsum
        elif name == "torchmetrics":
            module.Metric = torch.nn.Module
        elif name == "torch_geometric.nn.resolver":

            def activation_resolver(act="relu", **kw):
                if callable(act) and not isinstance(act, str):
                    return act
                table = {"relu": torch.nn.ReLU, "tanh": torch.nn.Tanh,
                         "elu": torch.nn.ELU, None: torch.nn.Identity}
                return table[act](**kw)

            module.activation_resolver = activation_resolver
        elif name == "ipdb":
            module.set_trace = lambda *a, **k: None


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "laplace"))


def install() -> None:
    """Make ``import laplace, curvlinops, gnn.models`` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.meta_path.append(_StubFinder())  # appended: real packages win if present
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load():
    """Return a namespace with the reference classes on the hot path."""
    install()
    import laplace  # noqa: F401
    from laplace import Laplace
    from laplace.curvature import CurvlinopsGGN, GGNInterface
    from laplace.utils import Kron, KronDecomposed
    from gnn.models.layers import GCNConv
    from gnn.models.models import GCN
    from gnn.models.utils import normalize_adj

    return types.SimpleNamespace(
        Laplace=Laplace, CurvlinopsGGN=CurvlinopsGGN, GGNInterface=GGNInterface,
        Kron=Kron, KronDecomposed=KronDecomposed, GCNConv=GCNConv, GCN=GCN,
        normalize_adj=normalize_adj,
    )
