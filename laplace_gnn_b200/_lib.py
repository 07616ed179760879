"""ctypes binding of the C-ABI in ``include/lgnn.h`` (``liblgnn.so``, built in-tree by
``__graft_entry__.build()`` / ``laplace_gnn_b200/csrc/Makefile``).

There is NO CPU fallback: if the shared library is missing, or a tensor is not on a CUDA
device, the call raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LGNN_LIB_PATH: another build of the same ABI (kernel bisects on the GPU box); default: the in-tree library
LIB_PATH = os.environ.get("LGNN_LIB_PATH") or os.path.join(_HERE, "liblgnn.so")

OK = 0
HESS_REFERENCE, HESS_GGN = 0, 1
SPMM_NONE, SPMM_RELU, SPMM_FORCE_LDG, SPMM_FORCE_BULK, SPMM_NO_HUB_ROWS = 0, 1, 2, 4, 8
SYRK_AUTO, SYRK_SIMT, SYRK_TCGEN05 = 0, 1, 2

_i64, _i32, _f32, _vp, _sz = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); every symbol include/lgnn.h declares
PROTOTYPES = {
    "lgnn_abi_version": (C.c_int, []),
    "lgnn_last_error": (C.c_char_p, []),
    "lgnn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "lgnn_csr_build_workspace_bytes": (_sz, [_i64, _i64, C.c_int]),
    "lgnn_csr_build_count": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp, _sz, _vp, _vp]),
    "lgnn_csr_build_fill": (C.c_int, [_vp, _sz, _i64, _i64, C.c_int, _vp, _vp, _vp]),
    "lgnn_csr_transpose_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "lgnn_csr_transpose": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lgnn_degree_norm": (C.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "lgnn_edge_values": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "lgnn_row_partition": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "lgnn_halo_mark": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "lgnn_csr_slice_remap": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "lgnn_spmm_f32": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, C.c_int, _vp]),
    "lgnn_unit_slabs_supported": (C.c_int, [_i64, _i64]),
    "lgnn_unit_pack_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "lgnn_spmm_units_f32": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, C.c_int, _vp]),
    "lgnn_unit_pack_ragged_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "lgnn_spmm_units_ragged_f32": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, C.c_int, _vp]),
    "lgnn_sddmm_f32": (C.c_int, [_i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, C.c_int, _vp]),
    "lgnn_softmax_ce_sum": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp]),
    "lgnn_hess_rhs_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, C.c_int, _vp, _vp]),
    "lgnn_hess_rhs_pitched_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _i64, C.c_int, _vp, _vp]),
    "lgnn_spmm_hess_supported": (C.c_int, [_i64, _i64]),
    "lgnn_hess_stats_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, C.c_int, _vp, _i64, _vp]),
    "lgnn_spmm_hess_f32": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, C.c_int, _vp]),
    "lgnn_mask_edge_values": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "lgnn_relu_mask_mul_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i64, _vp]),
    "lgnn_gemm_mask_supported": (C.c_int, [_i64, _i64]),
    "lgnn_gemm_mask_kpad": (_i64, [_i64]),
    "lgnn_gemm_mask_prepare_f32": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "lgnn_gemm_mask_f32": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "lgnn_gemm_bias_f32": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "lgnn_syrk_workspace_bytes": (_sz, [_i64, _i64, C.c_int]),
    "lgnn_syrk_f32": (C.c_int, [_vp, _i64, _i64, _i64, _f32, _f32, _vp, _i64, _vp, _sz, C.c_int, _vp]),
}

_LIB = None


class LgnnError(RuntimeError):
    pass


def load(path: str | None = None):
    """Load liblgnn.so and attach prototypes.  Raises if the library is absent."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LgnnError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C laplace_gnn_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.lgnn_abi_version() != 1:
        raise LgnnError(f"ABI version mismatch: library {lib.lgnn_abi_version()}, binding 1")
    if path is None:
        _LIB = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().lgnn_last_error()
        raise LgnnError(f"{what or 'lgnn call'} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise LgnnError("lgnn kernels take CUDA tensors only (there is no CPU fallback); got a "
                        f"{t.device} tensor")
    if t.device.index != torch.cuda.current_device():
        # stream() is the current stream of the CURRENT device: launching there with another device's pointers
        # is an invalid launch at best.  The package's entry points (B200GGN.kron / diag, GCNConvFunction,
        # Graph.from_edge_index) switch to their tensors' device themselves (on_device_of); direct callers of
        # ops.* must do the same.
        raise LgnnError(f"tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                        "wrap the call in `with torch.cuda.device(tensor.device):`")
    return t.data_ptr()


def on_device_of(t: torch.Tensor):
    """Context manager making ``t``'s device the current CUDA device (no-op for CPU tensors: the CPU double of
    the tests drives the same host code)."""
    import contextlib
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return _LAUNCHES[0]


_LAUNCHES = [0]


def count_launches(n: int) -> None:
    """Bookkeeping for bench.py's gpu_launches claim (kernels of liblgnn launched so far)."""
    _LAUNCHES[0] += n
