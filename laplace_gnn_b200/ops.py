"""Tensor-level wrappers over the C-ABI (``include/lgnn.h``).  Every function enqueues
hand-written sm_100a kernels on torch's current CUDA stream and returns torch tensors that
only serve as device-memory handles.  No CPU path exists here.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import check, ptr, stream


@dataclass
class CSR:
    """Row-compressed sparse matrix on the device: rowptr int64 [n_rows+1], col int32 [nnz],
    val fp32 [nnz] (or None for a pattern)."""
    n_rows: int
    n_cols: int
    rowptr: torch.Tensor
    col: torch.Tensor
    val: torch.Tensor | None = None
    max_row_nnz: int | None = None     # longest row, if known (graph build): lets SpMM skip its hub passes
    masked: bool = False               # some edge values were zeroed on purpose (their gathers are skipped)
    _nnz_live: torch.Tensor | None = None   # 0-d device count of the live edges (profiling only)

    def nnz_gathered_dev(self) -> torch.Tensor | int:
        """Non-zeros whose source row is actually read — as an 8-byte DEVICE scalar when masked (enqueued, no
        host sync), so that a profile record never has to hold the matrix to count them later."""
        if not self.masked:
            return self.nnz
        if self._nnz_live is None:
            self._nnz_live = torch.count_nonzero(self.val)
        return self._nnz_live

    @property
    def nnz_gathered(self) -> int:
        """Host value of the above (one sync when masked; profiling / tests only)."""
        return int(self.nnz_gathered_dev())

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


# bench.py sets PROFILE = [] to collect one record per SpMM / SYRK launch: CUDA events on the
# launching stream plus the algorithmic bytes (SpMM) or flops (SYRK) of that launch.  A record holds NUMBERS
# only — python scalars, or 0-d device tensors where the count needs a reduction on the device (live edges of a
# masked matrix, live units gathered by the unit SpMM): a record that kept the slab, the activations or the masked
# edge values alive pinned ~5.7 GB of HBM per timed step (round-1 scaling runs at 2 and 4 GPUs ran out of memory).
PROFILE: list | None = None
_LIVE_GATHER: dict = {}     # (activation ptr, shape, col ptr) -> 0-d int64: sum over edges of the source's live units


def profile_begin() -> None:
    global PROFILE
    PROFILE = []


def profile_end() -> list:
    """Stop recording; returns the records with every count turned into a python number."""
    global PROFILE
    recs, PROFILE = PROFILE or [], None
    for r in recs:
        r["bytes"] = _number(r["bytes"])
    _LIVE_GATHER.clear()
    return recs


def _number(w):
    if callable(w):
        w = w()
    return float(w.item()) if isinstance(w, torch.Tensor) else float(w)


def live_gather_sum(act: torch.Tensor, h: int, col: torch.Tensor) -> torch.Tensor:
    """sum_e k[col[e]], k[j] = live units of node j — the unit SpMM's gathered slots per column, as a 0-d device
    tensor.  Profiling only; cached per (activation buffer, matrix): the column groups of a pass and — with the
    activations in the persistent workspace — the steps of a bench run share one evaluation."""
    key = (act.data_ptr(), tuple(act.shape), h, col.data_ptr(), col.numel())
    t = _LIVE_GATHER.get(key)
    if t is None:
        k = act if act.dim() == 1 else (act[:, :h] > 0).sum(1)      # 1-D: the live counts themselves
        t = _LIVE_GATHER[key] = k[col.to(torch.int64)].sum()
    return t


def timed(kind: str, d: int = 0, work: float = 0.0):
    """Context manager used by the backend around library GEMMs / small kernels when profiling."""
    return _Timed(kind, d, work)


# LGNN_NVTX=1: an NVTX range per launch group ("spmm_units d=4096", "syrk d=256", ...), so that ncu / nsys can
# filter by stage (ncu --nvtx --nvtx-include "spmm_units d=4096/").  Off by default: no call is made.
NVTX = os.environ.get("LGNN_NVTX") == "1"


class _Timed:
    def __init__(self, kind: str, d: int, work: float):
        self.rec = None
        self.nvtx = f"{kind} d={d}" if NVTX else None
        if PROFILE is not None:
            self.rec = {"kind": kind, "d": d, "bytes": work,
                        "start": torch.cuda.Event(enable_timing=True),
                        "end": torch.cuda.Event(enable_timing=True)}

    def __enter__(self):
        if self.nvtx is not None:
            torch.cuda.nvtx.range_push(self.nvtx)
        if self.rec is not None:
            self.rec["start"].record()
        return self

    def __exit__(self, *exc):
        if self.rec is not None:
            self.rec["end"].record()
            PROFILE.append(self.rec)
        if self.nvtx is not None:
            torch.cuda.nvtx.range_pop()
        return False


def spmm_algorithmic_bytes(n_rows: int, nnz: int, d: int) -> int:
    """col int32 + val fp32 per non-zero, int64 rowptr, one gathered source row of d floats per
    non-zero (no-reuse model), the output."""
    return nnz * 8 + (n_rows + 1) * 8 + nnz * d * 4 + n_rows * d * 4


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D row-major tensor (unit column stride)")
    return t


# ------------------------------------------------------------------------------ integer kernels
def csr_from_edge_index(edge_index: torch.Tensor, num_nodes: int, symmetric: bool = False) -> CSR:
    """Binary adjacency A with self loops as a CSR pattern (columns ascending, duplicates
    collapsed).  Mirrors edge_index_to_adj + clamp + fill_diagonal_(1) (+ optional A + A^T)."""
    lib = _lib.load()
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise TypeError("edge_index must be an int64 tensor of shape [2, E]")
    ei = edge_index.contiguous()
    n, e = int(num_nodes), int(ei.shape[1])
    dev = ei.device
    ws_bytes = lib.lgnn_csr_build_workspace_bytes(n, e, int(symmetric))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    src, dst = ei[0], ei[1]
    check(lib.lgnn_csr_build_count(ptr(src), ptr(dst), e, n, int(symmetric), ptr(ws), ws_bytes,
                                   ptr(rowptr), stream()), "lgnn_csr_build_count")
    _lib.count_launches(11)
    if int(ws[:4].view(torch.int32).item()) != 0:  # err flag (also syncs before reading nnz)
        raise ValueError("edge_index contains node ids outside [0, num_nodes)")
    nnz = int(rowptr[-1].item())
    col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    check(lib.lgnn_csr_build_fill(ptr(ws), ws_bytes, n, e, int(symmetric), ptr(rowptr), ptr(col),
                                  stream()), "lgnn_csr_build_fill")
    _lib.count_launches(1)
    return CSR(n, n, rowptr, col, None)


def csr_transpose(a: CSR) -> CSR:
    lib = _lib.load()
    dev = a.rowptr.device
    ws_bytes = lib.lgnn_csr_transpose_workspace_bytes(a.n_rows, a.n_cols, a.nnz)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    t_rowptr = torch.empty(a.n_cols + 1, dtype=torch.int64, device=dev)
    t_col = torch.empty(max(a.nnz, 1), dtype=torch.int32, device=dev)[: a.nnz]
    check(lib.lgnn_csr_transpose(a.n_rows, a.n_cols, ptr(a.rowptr), ptr(a.col), ptr(t_rowptr),
                                 ptr(t_col), ptr(ws), ws_bytes, stream()), "lgnn_csr_transpose")
    _lib.count_launches(7)
    return CSR(a.n_cols, a.n_rows, t_rowptr, t_col, None)


def degree_norm(a: CSR):
    """(deg int64 [n], dis fp32 [n]) from the pattern of A (row sums incl. self loop)."""
    lib = _lib.load()
    dev = a.rowptr.device
    deg = torch.empty(a.n_rows, dtype=torch.int64, device=dev)
    dis = torch.empty(a.n_rows, dtype=torch.float32, device=dev)
    check(lib.lgnn_degree_norm(a.n_rows, ptr(a.rowptr), ptr(deg), ptr(dis), stream()), "lgnn_degree_norm")
    _lib.count_launches(1)
    return deg, dis


def edge_values(a: CSR, dis: torch.Tensor, row_offset: int = 0) -> torch.Tensor:
    lib = _lib.load()
    val = torch.empty(max(a.nnz, 1), dtype=torch.float32, device=a.rowptr.device)[: a.nnz]
    check(lib.lgnn_edge_values(a.n_rows, row_offset, ptr(a.rowptr), ptr(a.col), ptr(dis), ptr(val),
                               stream()), "lgnn_edge_values")
    _lib.count_launches(1)
    return val


def row_partition(rowptr: torch.Tensor, nparts: int) -> torch.Tensor:
    lib = _lib.load()
    n = int(rowptr.numel()) - 1
    bounds = torch.empty(nparts + 1, dtype=torch.int64, device=rowptr.device)
    check(lib.lgnn_row_partition(ptr(rowptr), n, nparts, ptr(bounds), stream()), "lgnn_row_partition")
    _lib.count_launches(1)
    return bounds


def halo_columns(a: CSR, lo: int, hi: int) -> torch.Tensor:
    """Sorted unique column ids referenced by rows [lo, hi) outside [lo, hi)."""
    lib = _lib.load()
    flags = torch.zeros(a.n_cols, dtype=torch.uint8, device=a.rowptr.device)
    check(lib.lgnn_halo_mark(ptr(a.rowptr), ptr(a.col), lo, hi, ptr(flags), stream()), "lgnn_halo_mark")
    _lib.count_launches(1)
    return torch.nonzero(flags, as_tuple=False).flatten()


def csr_slice_remap(a: CSR, lo: int, hi: int, bounds: torch.Tensor, pad: int) -> CSR:
    """Rows [lo, hi) of ``a`` with columns remapped to the padded all-gather layout."""
    lib = _lib.load()
    dev = a.rowptr.device
    nparts = int(bounds.numel()) - 1
    b = int(a.rowptr[lo].item())
    e = int(a.rowptr[hi].item())
    nnz = e - b
    out_rowptr = torch.empty(hi - lo + 1, dtype=torch.int64, device=dev)
    out_col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    out_val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz] if a.val is not None else None
    check(lib.lgnn_csr_slice_remap(ptr(a.rowptr), ptr(a.col), ptr(a.val), lo, hi, ptr(bounds), nparts,
                                   pad, ptr(out_rowptr), ptr(out_col), ptr(out_val), stream()),
          "lgnn_csr_slice_remap")
    _lib.count_launches(1)
    return CSR(hi - lo, nparts * pad, out_rowptr, out_col, out_val, a.max_row_nnz)   # rows are kept whole


# ------------------------------------------------------------------------------ SpMM
_SPMM_IMPL = {"auto": 0, "ldg": _lib.SPMM_FORCE_LDG, "bulk": _lib.SPMM_FORCE_BULK}


def spmm(a: CSR, x: torch.Tensor, relu: bool = False, out: torch.Tensor | None = None,
         d: int | None = None, impl: str = "auto") -> torch.Tensor:
    """Y = A @ X[:, :d]  (optionally relu'd).  x: [>= n_cols, ld] row-major fp32."""
    lib = _lib.load()
    _f32c(x, "x")
    if x.shape[0] < a.n_cols:
        raise ValueError(f"spmm: x has {x.shape[0]} rows, matrix has {a.n_cols} columns")
    d = int(x.shape[1]) if d is None else int(d)
    if out is None:
        out = torch.empty(a.n_rows, d, dtype=torch.float32, device=x.device)
    _f32c(out, "out")
    if out.shape[0] < a.n_rows or out.shape[1] < d:
        raise ValueError("spmm: out too small")
    # masked edges still stream their (col, val) but pull no source row: nnz*8 + rowptr + output + live*d*4,
    # the live count a device scalar (no sync here, no reference to the matrix in the record)
    work = spmm_algorithmic_bytes(a.n_rows, a.nnz, d)
    if a.masked and PROFILE is not None:
        work = spmm_algorithmic_bytes(a.n_rows, 0, d) + a.nnz * 8 + a.nnz_gathered_dev() * (d * 4)
    with _Timed("spmm", d, work):
        check(lib.lgnn_spmm_f32(a.n_rows, a.nnz, ptr(a.rowptr), ptr(a.col), ptr(a.val), ptr(x), x.stride(0),
                                ptr(out), out.stride(0), d,
                                (_lib.SPMM_RELU if relu else _lib.SPMM_NONE) | _SPMM_IMPL[impl] |
                                (_lib.SPMM_NO_HUB_ROWS if a.max_row_nnz is not None and a.max_row_nnz <= 4096 else 0),
                                stream()), "lgnn_spmm_f32")
    _lib.count_launches(1)
    return out


# ------------------------------------------------------------------------------ unit-compacted slabs
@dataclass
class UnitSlab:
    """Slab [n_rows, g*h] whose rows hold only the live hidden units of their node, as [slot][g]
    (csrc/spmm_units.cu); ``hdr`` [n_rows, h/32, 2] int32 = (mask, first slot) per 32 units; ``act`` is
    the activation matrix that decided which units live (kept for the byte accounting only)."""
    n_rows: int
    g: int
    h: int
    slab: torch.Tensor
    hdr: torch.Tensor
    act: torch.Tensor | None = None
    _live: torch.Tensor | None = None
    ragged: bool = False        # rows back to back in a flat ``slab``, absolute slots in ``hdr`` (unit_pack_ragged)

    def live_units(self) -> torch.Tensor:
        """int64 [n_rows]: live units per node (profiling only; evaluated outside timed regions)."""
        if self._live is None:
            self._live = (self.act[:, : self.h] > 0).sum(1)
        return self._live


def unit_slabs_supported(g: int, h: int) -> bool:
    return bool(_lib.load().lgnn_unit_slabs_supported(int(g), int(h)))


def unit_pack(slab: torch.Tensor, act: torch.Tensor, g: int, hdr: torch.Tensor | None = None) -> UnitSlab:
    """Compact the dense rows [g][h] of ``slab`` ([n_rows, >= g*h], h = act.shape[1]) IN PLACE: units with
    act[n, u] <= 0 are dropped, whatever the slab holds there."""
    lib = _lib.load()
    _f32c(slab, "slab"); _f32c(act, "act")
    n, h = int(act.shape[0]), int(act.shape[1])
    if slab.shape[0] < n or slab.shape[1] < g * h:
        raise ValueError("unit_pack: slab smaller than [act rows, g*h]")
    if hdr is None:
        hdr = torch.empty(max(n, 1), h // 32, 2, dtype=torch.int32, device=slab.device)
    if hdr.dtype != torch.int32 or hdr.numel() < n * (h // 32) * 2 or not hdr.is_contiguous():
        raise ValueError("unit_pack: hdr must be a contiguous int32 buffer of n_rows * h/32 * 2 entries")
    with _Timed("unit_pack", g * h, float(n) * h * (g * 6 + 4)):     # dense read + ~half written back + act
        check(lib.lgnn_unit_pack_f32(ptr(slab), slab.stride(0), ptr(act), act.stride(0), n, int(g), h, ptr(hdr),
                                     stream()), "lgnn_unit_pack_f32")
    _lib.count_launches(1)
    return UnitSlab(n, int(g), h, slab, hdr, act)


def unit_row_slots(act: torch.Tensor, g: int) -> torch.Tensor:
    """int64 [n_rows]: slots the compacted row of each node takes (its live units; for g % 4 == 2 every 32-unit
    block rounded up to an even count, the layout of csrc/spmm_units_even.cu)."""
    n, h = int(act.shape[0]), int(act.shape[1])
    if g % 4 == 0:
        return (act > 0).sum(1)
    per_block = (act > 0).view(n, h // 32, 32).sum(2)
    return ((per_block + 1) // 2 * 2).sum(1)


def unit_pack_ragged(src: torch.Tensor | None, act: torch.Tensor, g: int, row_first: torch.Tensor,
                     dst: torch.Tensor | None, hdr: torch.Tensor | None) -> None:
    """Rows of ``src`` ([n_rows, >= g*h] dense [g][h]) packed back to back into the flat buffer ``dst``: node n's
    live slots start at the absolute slot ``row_first[n]`` (int64; see unit_row_slots).  ``hdr`` (int32
    [n_rows, h/32, 2], optional) receives (mask, absolute first slot) per 32 units; ``src=None`` writes the
    headers only.  What the ranks of the row-partitioned backward exchange (curvature.py)."""
    lib = _lib.load()
    _f32c(act, "act")
    n, h = int(act.shape[0]), int(act.shape[1])
    if row_first.dtype != torch.int64 or row_first.numel() < n or not row_first.is_contiguous():
        raise ValueError("unit_pack_ragged: row_first must be a contiguous int64 vector with one entry per row")
    if src is not None:
        _f32c(src, "src")
        if dst is None:
            raise ValueError("unit_pack_ragged: src without dst")
        if dst.dtype != torch.float32 or not dst.is_contiguous():
            raise ValueError("unit_pack_ragged: dst must be a contiguous float32 buffer")
        if src.shape[0] < n or src.shape[1] < g * h:
            raise ValueError("unit_pack_ragged: src smaller than [act rows, g*h]")
    if hdr is not None and (hdr.dtype != torch.int32 or hdr.numel() < n * (h // 32) * 2 or not hdr.is_contiguous()):
        raise ValueError("unit_pack_ragged: hdr must be a contiguous int32 buffer of n_rows * h/32 * 2 entries")
    with _Timed("unit_pack", g * h, float(n) * h * ((g * 6 + 4) if src is not None else 4)):
        check(lib.lgnn_unit_pack_ragged_f32(ptr(src) if src is not None else None, src.stride(0) if src is not None else 0,
                                            ptr(act), act.stride(0), n, int(g), h, ptr(row_first),
                                            ptr(dst) if src is not None else None,
                                            ptr(hdr) if hdr is not None else None, stream()),
              "lgnn_unit_pack_ragged_f32")
    _lib.count_launches(1)


def spmm_units(a: CSR, us: UnitSlab, out: torch.Tensor | None = None, variant: int = 0) -> torch.Tensor:
    """Y = A @ dense(us), Y: [n_rows, g*h] dense; bit-identical to ``spmm(a, dense slab)``."""
    lib = _lib.load()
    if us.n_rows < a.n_cols:
        raise ValueError("spmm_units: fewer slab rows than matrix columns")
    d = us.g * us.h
    if out is None:
        out = torch.empty(a.n_rows, d, dtype=torch.float32, device=us.slab.device)
    _f32c(out, "out")
    if out.shape[0] < a.n_rows or out.shape[1] < d:
        raise ValueError("spmm_units: out too small")
    # bytes actually asked of HBM: (col, val) + one header row per edge, rowptr, the live units of every
    # gathered row, the dense output
    nblk = us.h // 32
    work = a.nnz * (8 + 8 * nblk) + (a.n_rows + 1) * 8 + a.n_rows * d * 4
    if PROFILE is not None and (us.act is not None or us._live is not None):
        work = work + live_gather_sum(us.act if us.act is not None else us._live, us.h, a.col) * (us.g * 4)
    with _Timed("spmm_units", d, work) as rec:
        if us.ragged:
            check(lib.lgnn_spmm_units_ragged_f32(a.n_rows, us.n_rows, a.nnz, ptr(a.rowptr), ptr(a.col), ptr(a.val),
                                                 ptr(us.slab), us.slab.numel(), ptr(us.hdr), us.g, us.h, ptr(out),
                                                 out.stride(0), (int(variant) & 0xff) << 8, stream()),
                  "lgnn_spmm_units_ragged_f32")
        else:
            check(lib.lgnn_spmm_units_f32(a.n_rows, us.n_rows, a.nnz, ptr(a.rowptr), ptr(a.col), ptr(a.val),
                                          ptr(us.slab), us.slab.stride(0), ptr(us.hdr), us.g, us.h, ptr(out),
                                          out.stride(0), (int(variant) & 0xff) << 8, stream()), "lgnn_spmm_units_f32")
    if rec.rec is not None:
        rec.rec["dense_bytes"] = spmm_algorithmic_bytes(a.n_rows, a.nnz, d)
    _lib.count_launches(1)
    return out


# ------------------------------------------------------------------------------ SDDMM (adjacency gradient)
def sddmm(rows: torch.Tensor, cols: torch.Tensor, u: torch.Tensor, v: torch.Tensor, d: int | None = None,
          out: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
    """out[e] (+)= u[rows[e], :d] . v[cols[e], :d]  (rows / cols int32)."""
    lib = _lib.load()
    _f32c(u, "u"); _f32c(v, "v")
    if rows.dtype != torch.int32 or cols.dtype != torch.int32 or rows.numel() != cols.numel():
        raise TypeError("sddmm: rows and cols must be int32 tensors of the same length")
    rows, cols = rows.contiguous(), cols.contiguous()
    d = int(min(u.shape[1], v.shape[1])) if d is None else int(d)
    n_pairs = int(rows.numel())
    if out is None:
        if accumulate:
            raise ValueError("sddmm: accumulate needs out")
        out = torch.empty(max(n_pairs, 1), dtype=torch.float32, device=u.device)[:n_pairs]
    if out.dtype != torch.float32 or out.numel() < n_pairs or not out.is_contiguous():
        raise ValueError("sddmm: out must be a contiguous float32 tensor with one entry per pair")
    with _Timed("sddmm", d, float(n_pairs) * (8 + 4 * d) + 8.0 * n_pairs):
        check(lib.lgnn_sddmm_f32(n_pairs, ptr(rows), ptr(cols), ptr(u), u.stride(0), ptr(v), v.stride(0), d,
                                 ptr(out), int(bool(accumulate)), stream()), "lgnn_sddmm_f32")
    _lib.count_launches(1)
    return out


# ------------------------------------------------------------------------------ loss / Hessian sqrt
def softmax_ce_sum(logits: torch.Tensor, idx: torch.Tensor, y: torch.Tensor, C: int | None = None):
    """(sum CE as a 0-d float64 tensor, number of argmax hits as 0-d int64)."""
    lib = _lib.load()
    _f32c(logits, "logits")
    C = int(logits.shape[1]) if C is None else int(C)
    loss = torch.zeros((), dtype=torch.float64, device=logits.device)
    hits = torch.zeros((), dtype=torch.int64, device=logits.device)
    idx = idx.contiguous()
    y = y.contiguous()
    if idx.dtype != torch.int64 or y.dtype != torch.int64:
        raise TypeError("idx and y must be int64")
    check(lib.lgnn_softmax_ce_sum(ptr(logits), logits.stride(0), C, ptr(idx), ptr(y), idx.numel(),
                                  ptr(loss), ptr(hits), stream()), "lgnn_softmax_ce_sum")
    _lib.count_launches(1)
    return loss, hits


def hess_rhs(logits: torch.Tensor, idx: torch.Tensor, c0: int, ncols: int, delta: torch.Tensor,
             ldc: int, mode: str = "reference", C: int | None = None) -> torch.Tensor:
    """Scatter-add the Hessian-sqrt columns [c0, c0+ncols) into delta ([n_nodes, >= ncols*ldc], zeroed;
    a wider row leaves all-zero padding columns behind the group)."""
    lib = _lib.load()
    _f32c(logits, "logits")
    C = int(logits.shape[1]) if C is None else int(C)
    m = {"reference": _lib.HESS_REFERENCE, "ggn": _lib.HESS_GGN}.get(mode)
    if m is None:
        raise ValueError(f"hess_sqrt must be 'reference' or 'ggn', got {mode!r}")
    idx = idx.contiguous()
    _f32c(delta, "delta")
    check(lib.lgnn_hess_rhs_pitched_f32(ptr(logits), logits.stride(0), C, ptr(idx), idx.numel(), c0, ncols, ldc,
                                        delta.stride(0), m, ptr(delta), stream()), "lgnn_hess_rhs_pitched_f32")
    _lib.count_launches(1)
    return delta


def spmm_hess_supported(C: int, width: int) -> bool:
    return bool(_lib.load().lgnn_spmm_hess_supported(int(C), int(width)))


def hess_stats(logits: torch.Tensor, idx: torch.Tensor, mode: str = "reference", C: int | None = None,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """[n_nodes, 5*Cp] softmax statistics of the batch's train nodes (zero rows elsewhere) from which
    ``spmm_hess`` rebuilds the Hessian-sqrt right-hand sides on the fly (csrc/spmm_hess.cu)."""
    lib = _lib.load()
    _f32c(logits, "logits")
    C = int(logits.shape[1]) if C is None else int(C)
    m = {"reference": _lib.HESS_REFERENCE, "ggn": _lib.HESS_GGN}.get(mode)
    if m is None:
        raise ValueError(f"hess_sqrt must be 'reference' or 'ggn', got {mode!r}")
    cp = (C + 3) // 4 * 4
    n = int(logits.shape[0])
    if out is None:
        out = torch.empty(n, 5 * cp, dtype=torch.float32, device=logits.device)
    _f32c(out, "stats")
    if out.shape[0] < n or out.shape[1] < 5 * cp:
        raise ValueError("hess_stats: out too small")
    idx = idx.contiguous()
    with _Timed("hess_rhs", C, 0.0):
        out.zero_()
        check(lib.lgnn_hess_stats_f32(ptr(logits), logits.stride(0), C, ptr(idx), idx.numel(), m, ptr(out),
                                      out.stride(0), stream()), "lgnn_hess_stats_f32")
    _lib.count_launches(1)
    return out


SPMM_HESS_STAGED = True      # lab switch: the cp.async-ring kernel of csrc/spmm_hess.cu instead of the register one


def spmm_hess(a: CSR, stats: torch.Tensor, C: int, c0: int, ncols: int, width: int,
              out: torch.Tensor | None = None, staged: bool | None = None) -> torch.Tensor:
    """Y[i, c*Cp + k] = sum_j A[i, j] v_{j, c0+c}[k] for c < ncols (zero columns up to ``width``): the output-layer
    SpMM of a column group with its Hessian-sqrt right-hand sides rebuilt from ``stats`` per edge."""
    lib = _lib.load()
    _f32c(stats, "stats")
    cp = (int(C) + 3) // 4 * 4
    if stats.shape[0] < a.n_cols:
        raise ValueError("spmm_hess: fewer stats rows than matrix columns")
    d = int(width) * cp
    if out is None:
        out = torch.empty(a.n_rows, d, dtype=torch.float32, device=stats.device)
    _f32c(out, "out")
    if out.shape[0] < a.n_rows or out.shape[1] < d:
        raise ValueError("spmm_hess: out too small")
    # bytes asked of HBM: (col, val), rowptr, per gathered edge P and Q (2 Cp floats) and the group's A, S, V,
    # the output; the same launch priced as the materialised slab goes into dense_bytes
    per_edge = (2 * cp + 3 * int(ncols)) * 4
    work = a.nnz * 8 + (a.n_rows + 1) * 8 + a.n_rows * d * 4
    if PROFILE is not None:
        work = work + a.nnz_gathered_dev() * per_edge
    with _Timed("spmm_hess", d, work):
        check(lib.lgnn_spmm_hess_f32(a.n_rows, a.nnz, ptr(a.rowptr), ptr(a.col), ptr(a.val), ptr(stats),
                                     stats.stride(0), int(C), int(c0), int(ncols), int(width), ptr(out),
                                     out.stride(0), 0x100 if (SPMM_HESS_STAGED if staged is None else staged) else 0,
                                     stream()), "lgnn_spmm_hess_f32")
    _lib.count_launches(1)
    return out


def csr_with_masked_sources(a: CSR, keep: torch.Tensor) -> CSR:
    """Same pattern, edge values zeroed where the source row (column index) is not flagged in
    ``keep`` (uint8 [n_cols]): SpMM then skips the gathers of rows known to be all zero."""
    lib = _lib.load()
    if keep.dtype != torch.uint8 or keep.numel() < a.n_cols:
        raise TypeError("keep must be a uint8 tensor with one flag per column")
    val = torch.empty_like(a.val)
    check(lib.lgnn_mask_edge_values(a.nnz, ptr(a.col), ptr(a.val), ptr(keep), ptr(val), stream()),
          "lgnn_mask_edge_values")
    _lib.count_launches(1)
    return CSR(a.n_rows, a.n_cols, a.rowptr, a.col, val, a.max_row_nnz, masked=True)


def relu_mask_mul(inp: torch.Tensor, act: torch.Tensor, group: int, out: torch.Tensor | None = None,
                  d: int | None = None) -> torch.Tensor:
    """out[r, :] = inp[r, :] * (act[r // group, :] > 0)."""
    lib = _lib.load()
    _f32c(inp, "inp"); _f32c(act, "act")
    d = int(inp.shape[1]) if d is None else int(d)
    out = inp if out is None else out
    n_rows = int(act.shape[0])
    if inp.shape[0] < n_rows * group:
        raise ValueError("relu_mask_mul: inp has too few rows")
    check(lib.lgnn_relu_mask_mul_f32(ptr(inp), inp.stride(0), ptr(act), act.stride(0), ptr(out),
                                     out.stride(0), n_rows, group, d, stream()), "lgnn_relu_mask_mul_f32")
    _lib.count_launches(1)
    return out


# ------------------------------------------------------------------------------ fused (gZ W) ⊙ relu'
@dataclass
class PreparedWeight:
    """W [k, n] as the two K-major tensor-core operands of the fused GEMM: wt_hi = W^T (the tensor
    core truncates it to tf32), wt_lo = W^T - trunc_tf32(W^T); both [n, kpad], zero padded."""
    k: int
    n: int
    wt_hi: torch.Tensor
    wt_lo: torch.Tensor


def gemm_mask_supported(k: int, n: int) -> bool:
    return bool(_lib.load().lgnn_gemm_mask_supported(int(k), int(n)))


def gemm_mask_prepare(w: torch.Tensor) -> PreparedWeight:
    lib = _lib.load()
    _f32c(w, "w")
    k, n = int(w.shape[0]), int(w.shape[1])
    kpad = int(lib.lgnn_gemm_mask_kpad(k))
    hi = torch.empty(n, kpad, dtype=torch.float32, device=w.device)
    lo = torch.empty(n, kpad, dtype=torch.float32, device=w.device)
    check(lib.lgnn_gemm_mask_prepare_f32(ptr(w), w.stride(0), k, n, ptr(hi), ptr(lo), stream()),
          "lgnn_gemm_mask_prepare_f32")
    _lib.count_launches(1)
    return PreparedWeight(k, n, hi, lo)


def gemm_mask(a: torch.Tensor, w: PreparedWeight, act: torch.Tensor | None, group: int,
              out: torch.Tensor | None = None, m_rows: int | None = None) -> torch.Tensor:
    """out[r, :] = (a[r, :k] @ W) * (act[r // group, :] > 0)   (act=None: plain product)."""
    lib = _lib.load()
    _f32c(a, "a")
    m_rows = int(a.shape[0]) if m_rows is None else int(m_rows)
    if a.shape[1] < w.k:
        raise ValueError("gemm_mask: a has fewer columns than W has rows")
    if out is None:
        out = torch.empty(m_rows, w.n, dtype=torch.float32, device=a.device)
    _f32c(out, "out")
    if act is not None:
        _f32c(act, "act")
        if act.shape[0] * group < m_rows or act.shape[1] < w.n:
            raise ValueError("gemm_mask: act too small")
    with _Timed("gemm_mask", w.n, 2.0 * m_rows * w.k * w.n) as rec:
        check(lib.lgnn_gemm_mask_f32(ptr(a), a.stride(0), m_rows, w.k, ptr(w.wt_hi), ptr(w.wt_lo), w.n,
                                     ptr(act), 0 if act is None else act.stride(0), int(group),
                                     ptr(out), out.stride(0), stream()), "lgnn_gemm_mask_f32")
    if rec.rec is not None:       # 3xTF32 issues hi.hi + hi.lo + lo.hi over the K padded to the 8-wide MMA steps
        rec.rec["issued"] = 3.0 * 2.0 * m_rows * ((w.k + 7) // 8 * 8) * w.n
    _lib.count_launches(1)
    return out


GEMM_WIDTHS = (64, 128, 256)


def linear_prepare(weight: torch.Tensor) -> PreparedWeight | None:
    """``nn.Linear`` weight [d_out, d_in] as the resident operands of ``gemm_bias`` (W^T, output width zero-padded
    to 64 / 128 / 256), or None when the fused kernel does not take the shape (d_in > 256 or d_out > 256)."""
    d_out, d_in = int(weight.shape[0]), int(weight.shape[1])
    n_pad = next((w for w in GEMM_WIDTHS if w >= d_out), None)
    if n_pad is None or not gemm_mask_supported(d_in, n_pad):
        return None
    wt = torch.zeros(d_in, n_pad, dtype=torch.float32, device=weight.device)
    wt[:, :d_out] = weight.detach().t()
    return gemm_mask_prepare(wt)


def gemm_bias(a: torch.Tensor, w: PreparedWeight, bias: torch.Tensor | None, out: torch.Tensor,
              m_rows: int | None = None) -> torch.Tensor:
    """out[r, :w.n] = a[r, :w.k] @ W + bias  (the GCNConv linear layer on the 3xTF32 tcgen05 kernel; ``bias`` [w.n]
    or None; ``out`` needs a pitch >= w.n)."""
    lib = _lib.load()
    _f32c(a, "a"); _f32c(out, "out")
    m_rows = int(a.shape[0]) if m_rows is None else int(m_rows)
    if a.shape[1] < w.k or out.shape[0] < m_rows or out.stride(0) < w.n:
        raise ValueError("gemm_bias: a has fewer columns than W has rows, or out is too small")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() < w.n or not bias.is_contiguous()):
        raise ValueError("gemm_bias: bias must be a contiguous float32 vector with one entry per (padded) output column")
    with _Timed("gemm_fwd", w.n, 2.0 * m_rows * w.k * w.n) as rec:
        check(lib.lgnn_gemm_bias_f32(ptr(a), a.stride(0), m_rows, w.k, ptr(w.wt_hi), ptr(w.wt_lo), w.n, ptr(bias),
                                     ptr(out), out.stride(0), stream()), "lgnn_gemm_bias_f32")
    if rec.rec is not None:
        rec.rec["issued"] = 3.0 * 2.0 * m_rows * ((w.k + 7) // 8 * 8) * w.n
    _lib.count_launches(1)
    return out


# ------------------------------------------------------------------------------ SYRK
_SYRK_IMPL = {"auto": _lib.SYRK_AUTO, "simt": _lib.SYRK_SIMT, "tcgen05": _lib.SYRK_TCGEN05}
_ws_cache: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Scratch for split-K partials, one buffer per (device, stream): two streams may run SYRKs
    concurrently (overlapped column groups of the multi-GPU backward)."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def syrk(x: torch.Tensor, n: int | None = None, alpha: float = 1.0, beta: float = 0.0,
         out: torch.Tensor | None = None, impl: str = "auto", k_rows: int | None = None) -> torch.Tensor:
    """out = beta*out + alpha * X[:k_rows, :n]^T X[:k_rows, :n]  (both triangles)."""
    lib = _lib.load()
    _f32c(x, "x")
    n = int(x.shape[1]) if n is None else int(n)
    k_rows = int(x.shape[0]) if k_rows is None else int(k_rows)
    if out is None:
        if beta != 0.0:
            raise ValueError("syrk: beta != 0 needs out")
        out = torch.empty(n, n, dtype=torch.float32, device=x.device)
    _f32c(out, "out")
    code = _SYRK_IMPL[impl]
    nbytes = lib.lgnn_syrk_workspace_bytes(k_rows, n, code)
    ws = _workspace(nbytes, x.device)
    with _Timed("syrk", n, float(k_rows) * n * (n + 1)) as rec:
        check(lib.lgnn_syrk_f32(ptr(x), x.stride(0), k_rows, n, float(alpha), float(beta), ptr(out),
                                out.stride(0), ptr(ws), ws.numel(), code, stream()), "lgnn_syrk_f32")
    if rec.rec is not None and n <= 256 and impl != "simt" and x.stride(0) % 4 == 0:
        # tensor-core path: three products on M = 128-row blocks of the upper block-triangle (rows 0..127 against all
        # np columns, rows 128..255 against columns 128..np-1), K rounded up to the 16-row steps
        np_ = (n + 15) // 16 * 16
        rec.rec["issued"] = 3.0 * 2.0 * ((k_rows + 15) // 16 * 16) * 128.0 * (np_ + (np_ - 128 if np_ > 128 else 0))
    _lib.count_launches(2)
    return out
