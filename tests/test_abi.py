"""CPU: the C-ABI library loads and exports every symbol include/lgnn.h declares.
No compute entry point is called here (there is no GPU in this environment)."""
import ctypes
import os
import re

from conftest import ROOT
from laplace_gnn_b200 import _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lgnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgnn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in lgnn.h but not exported by liblgnn.so"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in _lib.PROTOTYPES"
    assert sorted(_lib.PROTOTYPES) == names


def test_abi_version_and_host_only_queries():
    lib = _lib.load()
    assert lib.lgnn_abi_version() == 1
    # pure host arithmetic, no device needed
    assert lib.lgnn_csr_build_workspace_bytes(1000, 5000, 0) > 5000 * 4
    assert lib.lgnn_csr_build_workspace_bytes(1000, 5000, 1) > lib.lgnn_csr_build_workspace_bytes(1000, 5000, 0)
    assert lib.lgnn_syrk_workspace_bytes(10000, 64, _lib.SYRK_SIMT) > 0
    assert lib.lgnn_csr_transpose_workspace_bytes(10, 10, 30) > 0


def test_bad_arguments_are_rejected_before_any_launch():
    lib = _lib.load()
    rc = lib.lgnn_spmm_f32(-1, 0, None, None, None, None, 4, None, 4, 4, 0, None)
    assert rc == -1 and b"spmm" in lib.lgnn_last_error()
    rc = lib.lgnn_hess_rhs_f32(None, 4, 4, None, 1, 0, 4, 4, 7, None, None)
    assert rc == -1
    rc = lib.lgnn_syrk_f32(None, 8, 10, 8, 1.0, 0.0, None, 8, None, 0, 0, None)
    assert rc == -1
    # unit-compacted slabs: shape support is a host-side predicate; bad shapes / pitches never launch
    assert lib.lgnn_unit_slabs_supported(12, 256) == 1 and lib.lgnn_unit_slabs_supported(16, 1024) == 1
    assert lib.lgnn_unit_slabs_supported(11, 256) == 0 and lib.lgnn_unit_slabs_supported(12, 100) == 0
    # g % 4 == 2: the even-group kernels (spmm_units_even.cu) behind the same entry points
    assert all(lib.lgnn_unit_slabs_supported(g, 256) == 1 for g in (2, 6, 10, 14))
    assert lib.lgnn_unit_slabs_supported(18, 256) == 0 and lib.lgnn_unit_slabs_supported(0, 256) == 0
    assert lib.lgnn_unit_pack_f32(None, 1536, None, 256, 10, 6, 256, None, None) == -1          # null pointers
    assert lib.lgnn_spmm_units_f32(5, 5, 10, None, None, None, None, 1536, None, 6, 256, None, 1536, 0, None) == -1
    assert lib.lgnn_unit_pack_f32(None, 3072, None, 256, 10, 11, 256, None, None) == -5          # LGNN_E_UNSUPPORTED
    assert lib.lgnn_unit_pack_f32(None, 3072, None, 256, 10, 12, 256, None, None) == -1          # null pointers
    assert lib.lgnn_unit_pack_f32(None, 3072, None, 256, 0, 12, 256, None, None) == 0            # nothing to do
    assert lib.lgnn_spmm_units_f32(5, 5, 10, None, None, None, None, 3072, None, 12, 256, None, 3072, 0, None) == -1
    assert lib.lgnn_spmm_units_f32(5, 5, 10, None, None, None, None, 3072, None, 20, 256, None, 3072, 0, None) == -5
    assert lib.lgnn_sddmm_f32(-1, None, None, None, 4, None, 4, 4, None, 0, None) == -1
    assert lib.lgnn_sddmm_f32(0, None, None, None, 4, None, 4, 4, None, 0, None) == 0
    assert lib.lgnn_sddmm_f32(3, None, None, None, 4, None, 4, 4, None, 0, None) == -1 and b"sddmm" in lib.lgnn_last_error()
    assert lib.lgnn_hess_rhs_pitched_f32(None, 4, 4, None, 1, 0, 4, 4, 8, 0, None, None) == -1
    # on-the-fly Hessian-sqrt SpMM: shape predicate and argument checks on the host
    assert lib.lgnn_spmm_hess_supported(47, 16) == 1 and lib.lgnn_spmm_hess_supported(64, 1) == 1
    assert lib.lgnn_spmm_hess_supported(65, 4) == 0 and lib.lgnn_spmm_hess_supported(47, 17) == 0
    assert lib.lgnn_hess_stats_f32(None, 48, 47, None, 10, 0, None, 240, None) == -1
    assert lib.lgnn_spmm_hess_f32(5, 10, None, None, None, None, 240, 47, 40, 8, 8, None, 384, 0, None) == -1   # c0 + ncols > C
    assert lib.lgnn_spmm_hess_f32(5, 10, None, None, None, None, 239, 47, 0, 8, 8, None, 384, 0, None) == -1    # stats pitch < 5 Cp
    assert lib.lgnn_spmm_hess_f32(5, 10, None, None, None, None, 240, 47, 0, 8, 20, None, 960, 0, None) == -5   # group too wide
    assert lib.lgnn_spmm_hess_f32(0, 0, None, None, None, None, 240, 47, 0, 8, 8, None, 384, 0x100, None) == 0


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    with pytest.raises(_lib.LgnnError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "nope.so"))


def test_cpu_tensors_are_refused():
    import pytest, torch
    with pytest.raises(_lib.LgnnError, match="CUDA tensors only"):
        _lib.ptr(torch.zeros(4))
