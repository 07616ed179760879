/*
 * lgnn.h — C ABI of the B200-native GCN + KFAC-GGN Laplace hot path.
 *
 * Plain C types only: device pointers, sizes, a CUDA stream handle.  No torch types,
 * no exceptions, no hidden allocation (the caller owns every buffer; kernels that need
 * scratch take a caller-provided workspace whose size comes from a *_workspace_bytes()
 * query).  Every entry point returns 0 on success or a negative LGNN_E_* code and
 * records a message retrievable with lgnn_last_error() (thread-local).
 * All pointers are DEVICE pointers unless a parameter is documented as "host".
 * Kernels are enqueued on `stream` and never synchronise the device.
 *
 * The reference (anitasyang/Laplace-GNN, pure Python) has no FFI; each entry point
 * cites the reference lines whose arithmetic it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 */
#ifndef LGNN_H_
#define LGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGNN_ABI_VERSION 1

enum {
  LGNN_OK = 0,
  LGNN_E_BADARG = -1,      /* null pointer, negative size, unknown mode */
  LGNN_E_ALIGN = -2,       /* pointer / leading dimension not aligned as documented */
  LGNN_E_NOMEM = -3,       /* workspace too small */
  LGNN_E_CUDA = -4,        /* a CUDA runtime call or launch failed */
  LGNN_E_UNSUPPORTED = -5  /* shape outside what the kernel supports */
};

/* Hessian-square-root mode of lgnn_hess_rhs_f32 */
enum {
  LGNN_HESS_REFERENCE = 0, /* curvlinops/kfac.py:631-661 as vendored (sqrt not detached) */
  LGNN_HESS_GGN = 1        /* textbook sqrt(p_c)(e_c - p): upstream curvlinops / asdl / backpack */
};

/* SpMM epilogue flags */
enum {
  LGNN_SPMM_NONE = 0,
  LGNN_SPMM_RELU = 1,       /* y = max(y, 0) fused into the store */
  LGNN_SPMM_FORCE_LDG = 2,  /* kernel selection override (tests / profiling): warp-per-row LDG.128 kernel */
  LGNN_SPMM_FORCE_BULK = 4, /* kernel selection override: bulk-async (TMA 1-D copy) shared-memory ring kernel */
  LGNN_SPMM_NO_HUB_ROWS = 8 /* caller guarantees that no row has more than 4096 non-zeros: the passes that
                               split hub rows of power-law graphs across warps / CTAs are skipped */
};

/* SYRK implementation selector */
enum {
  LGNN_SYRK_AUTO = 0,      /* tcgen05 3xTF32 when the shape allows, else SIMT fp32 */
  LGNN_SYRK_SIMT = 1,      /* CUDA-core fp32 FMA */
  LGNN_SYRK_TCGEN05 = 2    /* tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators */
};

typedef void* lgnn_stream_t; /* cudaStream_t */

int lgnn_abi_version(void);
const char* lgnn_last_error(void);
/* sm count and compute capability of the current device (host out-params). */
int lgnn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * (c) integer kernels — bit-exact contract
 * ---------------------------------------------------------------------------------------- */

/* Edge list -> CSR pattern of the binary adjacency A with self loops.
 * Replaces gnn/utils.py:325-330 (edge_index_to_adj: scipy coo -> dense, duplicates summed),
 * gnn/marglik_training.py:403-405 (clamp to 1), gnn/models/base_gnn.py:68-73 (optional
 * A + A^T clamp) and gnn/models/models.py:23 (fill_diagonal_(1)).
 * A[src[e], dst[e]] = 1; duplicates collapse; the diagonal is set; columns ascend within a row.
 *
 * Two calls because nnz is only known after de-duplication:
 *   lgnn_csr_build_count  fills rowptr[0..n] (rowptr[n] = nnz) and keeps the sorted rows in ws;
 *   the caller reads rowptr[n], allocates col[nnz], then
 *   lgnn_csr_build_fill   compacts the unique columns into col.
 * src/dst: int64 [n_edges] (the two rows of the reference's edge_index). */
size_t lgnn_csr_build_workspace_bytes(int64_t n, int64_t n_edges, int symmetric);
int lgnn_csr_build_count(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n,
                         int symmetric, void* ws, size_t ws_bytes, int64_t* rowptr,
                         lgnn_stream_t stream);
int lgnn_csr_build_fill(const void* ws, size_t ws_bytes, int64_t n, int64_t n_edges, int symmetric,
                        const int64_t* rowptr, int32_t* col, lgnn_stream_t stream);

/* Pattern transpose of an [n_rows x n_cols] CSR (columns ascend within each output row).
 * Â's pattern is A^T (gnn/models/utils.py:112: (adj @ D).T @ D). */
size_t lgnn_csr_transpose_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t nnz);
int lgnn_csr_transpose(int64_t n_rows, int64_t n_cols, const int64_t* rowptr, const int32_t* col,
                       int64_t* t_rowptr, int32_t* t_col, void* ws, size_t ws_bytes,
                       lgnn_stream_t stream);

/* Degree normalisation, gnn/models/utils.py:106-109: deg = row sums of A (int64, exact),
 * dis = deg^-1/2 in fp32 computed as IEEE 1/sqrt (inf -> 0). */
int lgnn_degree_norm(int64_t n, const int64_t* a_rowptr, int64_t* deg, float* dis,
                     lgnn_stream_t stream);

/* Edge values of Â (or Â^T): val[k] = dis[row_offset + i] * dis[col[k]] for k in row i.
 * gnn/models/utils.py:112. */
int lgnn_edge_values(int64_t n_rows, int64_t row_offset, const int64_t* rowptr, const int32_t* col,
                     const float* dis, float* val, lgnn_stream_t stream);

/* nnz-balanced contiguous row blocks (new: the reference is single-device).
 * bounds[r] = first row i with rowptr[i] >= floor(r*nnz/nparts); bounds[0]=0, bounds[nparts]=n. */
int lgnn_row_partition(const int64_t* rowptr, int64_t n, int32_t nparts, int64_t* bounds,
                       lgnn_stream_t stream);

/* Halo of a row block: flags[c] = 1 for every column c referenced by rows [lo, hi) that lies
 * outside [lo, hi); flags (uint8 [n_cols]) must be zeroed by the caller. */
int lgnn_halo_mark(const int64_t* rowptr, const int32_t* col, int64_t lo, int64_t hi,
                   uint8_t* flags, lgnn_stream_t stream);

/* Slice rows [lo, hi) of a CSR into a local CSR whose columns are remapped for a padded
 * all-gather layout: column c owned by rank q (bounds[q] <= c < bounds[q+1]) becomes
 * q*pad + (c - bounds[q]).  out_rowptr has hi-lo+1 entries starting at 0. */
int lgnn_csr_slice_remap(const int64_t* rowptr, const int32_t* col, const float* val, int64_t lo,
                         int64_t hi, const int64_t* bounds, int32_t nparts, int64_t pad,
                         int64_t* out_rowptr, int32_t* out_col, float* out_val,
                         lgnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (a) CSR SpMM — HBM-bound
 * ---------------------------------------------------------------------------------------- */

/* Y[i, 0:d] = epilogue( sum_k val[k] * X[col[k], 0:d] ), k in [rowptr[i], rowptr[i+1]).
 * Replaces `adj @ self.lin(x)` (gnn/models/layers.py:45-46) with a sparse Â, its autograd
 * backward Â^T @ grad, and — with d = (number of Hessian-sqrt columns) x (layer width) — the C
 * backward passes of curvlinops/kfac.py:653-661 in ONE multi-RHS pass.
 * X: [*, ldx] row-major fp32; Y: [n_rows, ldy]; nnz = rowptr[n_rows] (host copy, sizes the grid).
 * The 128-bit paths need X, Y 16-byte aligned and d, ldx, ldy multiples of 4; anything else takes
 * the scalar path.  Wide slabs whose column ranges (<= 4096 floats) give >= 14 KB ring stages
 * (d >= 3584) use the bulk-async shared-memory ring kernel, everything else the warp-per-row kernel. */
int lgnn_spmm_f32(int64_t n_rows, int64_t nnz, const int64_t* rowptr, const int32_t* col,
                  const float* val, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t d,
                  int flags, lgnn_stream_t stream);

/* Unit-compacted slabs.  delta[n, c, u] = (gZ W)[n, c, u] * 1[H[n, u] > 0] (curvlinops/kfac.py:653-661
 * through the relu of gnn/models/base_gnn.py:150): the g Hessian-sqrt columns of a node share ONE zero
 * pattern, that of the node's hidden units.  The slab row [g][h] of node n is rewritten in place as
 * [active unit slot][g] (the g values of a live unit contiguous) with a header word per 32 units,
 * hdr[n][w] = {uint32 mask, uint32 slot of the block's first live unit}; the SpMM then moves only the
 * live units' bytes and adds them in the neighbour order of lgnn_spmm_f32 (bit-identical result).
 * g even, 2 .. 16, h a multiple of 32 up to 1024 (lgnn_unit_slabs_supported).  For g % 4 == 2 a slot is
 * 8*odd bytes, so every block's run starts on an EVEN slot (hdr "first" = the previous blocks' live counts,
 * each rounded up to even): the runs stay 16-byte aligned for the SpMM's 16-byte copies, and the compact row
 * still fits the dense pitch.
 *   lgnn_unit_pack_f32    slab [n_rows, lds] dense [g][h] rows -> compact, in place; act = H [n_rows, lda]
 *                         decides which units live (values of dead units are dropped, whatever they hold);
 *                         hdr: uint2 [n_rows, h/32]
 *   lgnn_spmm_units_f32   Y[i, c*h + u] = sum_k val[k] * delta[col[k], c, u];  Y: [n_rows, ldy >= g*h] dense;
 *                         n_cols = rows of the slab (columns of the matrix).
 *                         Rows are walked whole by one warp per 32 units: meant for graphs without hub rows. */
int lgnn_unit_slabs_supported(int64_t g, int64_t h);
int lgnn_unit_pack_f32(float* slab, int64_t lds, const float* act, int64_t lda, int64_t n_rows, int64_t g,
                       int64_t h, void* hdr, lgnn_stream_t stream);
int lgnn_spmm_units_f32(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* rowptr, const int32_t* col,
                        const float* val, const float* slab, int64_t lds, const void* hdr, int64_t g,
                        int64_t h, float* y, int64_t ldy, int flags, lgnn_stream_t stream);

/* Ragged unit-compacted rows, for the row-partitioned backward (a rank packs its row block, the packed rows are
 * what the ranks exchange before the SpMM — the reference's per-layer Â^T product of curvlinops/kfac.py:653-661
 * on a graph split by rows).  Node n's live slots start at the ABSOLUTE slot row_first[n] of dst (rows back to
 * back: row_first = exclusive scan of the rows' slot counts, + the rank's base in the gathered buffer; for
 * g % 4 == 2 a 32-unit block takes its live count rounded up to even, as in lgnn_unit_pack_f32), and the header
 * words carry absolute slots, so the SpMM needs no pitch.
 *   lgnn_unit_pack_ragged_f32   src [n_rows, lds] dense [g][h] rows -> dst (must not overlap src);
 *                               src == NULL: headers only;  hdr == NULL: values only
 *   lgnn_spmm_units_ragged_f32  as lgnn_spmm_units_f32 on a slab of slab_floats floats packed that way
 *                               (at most 2^32 slots) */
int lgnn_unit_pack_ragged_f32(const float* src, int64_t lds, const float* act, int64_t lda, int64_t n_rows,
                              int64_t g, int64_t h, const int64_t* row_first, float* dst, void* hdr,
                              lgnn_stream_t stream);
int lgnn_spmm_units_ragged_f32(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* rowptr,
                               const int32_t* col, const float* val, const float* slab, int64_t slab_floats,
                               const void* hdr, int64_t g, int64_t h, float* y, int64_t ldy, int flags,
                               lgnn_stream_t stream);

/* Sampled dense-dense products: out[e] (+)= U[rows[e], 0:d] . V[cols[e], 0:d] for n_pairs (row, column)
 * pairs (int32).  The adjoint of Â[i, j] wherever the path computes Y = Â X is Ybar[i, :] . X[j, :]; summed
 * over the forward and the KFAC backward this is d marglik / dÂ on the requested entries — what the
 * reference materialises as the dense model.adj.grad (gnn/marglik_training.py:197-215,
 * gnn/models/models.py:100-115).  accumulate != 0 adds onto out. */
int lgnn_sddmm_f32(int64_t n_pairs, const int32_t* rows, const int32_t* cols, const float* u, int64_t ldu,
                   const float* v, int64_t ldv, int64_t d, float* out, int accumulate,
                   lgnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * loss + Hessian-sqrt right-hand sides
 * ---------------------------------------------------------------------------------------- */

/* loss[0] += sum_m CE(logits[idx[m], :], y[m])  (laplace/curvature/curvlinops.py:106,
 * CrossEntropyLoss(reduction="sum")).  loss is a device double the caller zeroes.
 * If n_correct != NULL it also counts argmax hits (int64 device counter). */
int lgnn_softmax_ce_sum(const float* logits, int64_t ld, int32_t C, const int64_t* idx,
                        const int64_t* y, int64_t m, double* loss, int64_t* n_correct,
                        lgnn_stream_t stream);

/* delta[idx[m], g, 0:C] += v_{m, c0+g}  for g in [0, ncols): the vector the reference injects at
 * the logits of train node idx[m] for Hessian-sqrt column c0+g
 * (curvlinops/kfac_utils.py:122-126 + curvlinops/kfac.py:631-661).  delta: [n_nodes, ncols, ldc]
 * row-major, zeroed by the caller (duplicate idx accumulate, like autograd's index backward).
 *   GGN:        v = sqrt(p_c) (e_c - p)
 *   REFERENCE:  v = sqrt(p_c) [ (e_c - p)(1 + (f_c - fbar)/2) - p*(f - fbar) ],  fbar = p.f */
int lgnn_hess_rhs_f32(const float* logits, int64_t ld, int32_t C, const int64_t* idx, int64_t m,
                      int32_t c0, int32_t ncols, int32_t ldc, int mode, float* delta,
                      lgnn_stream_t stream);
/* Same with an explicit row pitch ld_delta >= ncols*ldc (floats) of delta: column groups padded with
 * all-zero right-hand sides (the unit-compacted slabs want a multiple of 4 columns per group). */
int lgnn_hess_rhs_pitched_f32(const float* logits, int64_t ld, int32_t C, const int64_t* idx, int64_t m,
                              int32_t c0, int32_t ncols, int32_t ldc, int64_t ld_delta, int mode,
                              float* delta, lgnn_stream_t stream);

/* The same right-hand sides generated inside the output-layer SpMM instead of being materialised
 * (csrc/spmm_hess.cu).  v_{m,c} is a closed form of the node's softmax:
 *   v_c[k] = -(A_c P_k + S_c Q_k) (k != c),  V_c (k == c);   P = p, Q = p*(f - fbar), S_c = sqrt(p_c),
 *   A_c = sqrt(p_c)(1 + (f_c - fbar)/2)          (GGN mode: A = S, Q = 0)
 *   lgnn_hess_stats_f32   stats[idx[m]] = [P | Q | A | S | V], five vectors of Cp = roundup4(C) floats per node
 *                         (ld_stats >= 5*Cp, zeroed by the caller; A, S, V accumulate for duplicate idx);
 *   lgnn_spmm_hess_f32    Y[i, c*Cp + k] = sum_j val[ij] * v_{col[ij], c0+c}[k]  for c < ncols, zero for
 *                         ncols <= c < width; gathers 2*Cp + 3*ncols floats per edge instead of width*Cp.
 *                         C <= 64, width <= 16 (lgnn_spmm_hess_supported); rows are walked whole by one warp.
 *                         flags 0x100: the staged kernel (records copied through a cp.async ring; c0 % 4 == 0). */
int lgnn_spmm_hess_supported(int64_t C, int64_t width);
int lgnn_hess_stats_f32(const float* logits, int64_t ld, int32_t C, const int64_t* idx, int64_t m, int mode,
                        float* stats, int64_t ld_stats, lgnn_stream_t stream);
int lgnn_spmm_hess_f32(int64_t n_rows, int64_t nnz, const int64_t* rowptr, const int32_t* col, const float* val,
                       const float* stats, int64_t ld_stats, int32_t C, int32_t c0, int32_t ncols, int32_t width,
                       float* y, int64_t ldy, int flags, lgnn_stream_t stream);

/* out[k] = keep[col[k]] ? val[k] : 0 for the nnz entries of a CSR.  The right-hand sides injected at
 * the logits are zero outside the batch's train nodes (curvlinops/kfac.py:653-661 back-propagates
 * through model(X)[idx]); with the edges into those rows zeroed, the output-layer SpMM of the KFAC
 * backward skips their gathers (a zero edge value costs no load in lgnn_spmm_f32's row kernel). */
int lgnn_mask_edge_values(int64_t nnz, const int32_t* col, const float* val, const uint8_t* keep,
                          float* out, lgnn_stream_t stream);

/* out[r, j] = in[r, j] * (act[(r / group), j] > 0)   (relu' mask of the layer below;
 * autograd of base_gnn.py:150 inside kfac.py:653-661).  in/out: [n_rows*group, d], act: [n_rows, d]. */
int lgnn_relu_mask_mul_f32(const float* in, int64_t ldi, const float* act, int64_t lda, float* out,
                           int64_t ldo, int64_t n_rows, int32_t group, int64_t d,
                           lgnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (b) Kronecker-factor accumulation — dense SYRK, tensor-core bound
 * ---------------------------------------------------------------------------------------- */

/* C[n x n] = beta * C + alpha * X^T X,  X: [k_rows, ldx] row-major fp32 (first n columns used),
 * C: [n, ldc] row-major, both triangles written.
 * A_l = H^T H / N (curvlinops/kfac.py:819-875), G_l = sum_c gZ^T gZ (kfac.py:777-817).
 * Deterministic split-K: partial tiles go to the workspace and are reduced in fixed order. */
size_t lgnn_syrk_workspace_bytes(int64_t k_rows, int64_t n, int impl);
int lgnn_syrk_f32(const float* x, int64_t ldx, int64_t k_rows, int64_t n, float alpha, float beta,
                  float* c, int64_t ldc, void* ws, size_t ws_bytes, int impl,
                  lgnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * fused back-propagation step between two SpMMs — tensor-core / HBM balanced
 * ---------------------------------------------------------------------------------------- */

/* out[r, 0:n] = (A[r, 0:k] . W[0:k, 0:n]) * (act[r / group, 0:n] > 0),  r in [0, m_rows).
 * The relu-masked product delta_{l-1} = (gZ_l W_l) ⊙ 1[H_{l-1} > 0] that autograd computes between
 * the aggregation backward passes of curvlinops/kfac.py:653-661 (gnn/models/base_gnn.py:150,
 * gnn/models/layers.py:45).  A = gZ viewed as [nodes*group, k] (pitch lda), W = conv.lin.weight
 * [k = d_l, n = d_{l-1}] row-major (pitch ldw); act = H_{l-1} [nodes, n]; act == NULL skips the mask.
 * 3xTF32 on tcgen05 tensor cores; supported for k <= 256 and n in {64, 128, 256}
 * (lgnn_gemm_mask_supported); all pointers 16-byte aligned, pitches multiples of 4 floats.
 * The weights are prepared once per W into two [n, kpad] arrays (kpad = lgnn_gemm_mask_kpad(k)). */
int lgnn_gemm_mask_supported(int64_t k, int64_t n);
int64_t lgnn_gemm_mask_kpad(int64_t k);
int lgnn_gemm_mask_prepare_f32(const float* w, int64_t ldw, int64_t k, int64_t n, float* wt_hi,
                               float* wt_lo, lgnn_stream_t stream);
int lgnn_gemm_mask_f32(const float* a, int64_t lda, int64_t m_rows, int64_t k, const float* wt_hi,
                       const float* wt_lo, int64_t n, const float* act, int64_t ld_act, int32_t group,
                       float* out, int64_t ldo, lgnn_stream_t stream);

/* out[r, 0:n] = A[r, 0:k] . W[0:k, 0:n] + bias[0:n],  r in [0, m_rows): the linear layer of GCNConv,
 * Z_l = H_{l-1} W_l^T + 1 b_l^T (gnn/models/layers.py:45, torch.nn.Linear), on the same 3xTF32 tcgen05 kernel with
 * the bias added in its epilogue.  W here is conv.lin.weight TRANSPOSED ([k = d_{l-1}, n = d_l], prepared with
 * lgnn_gemm_mask_prepare_f32; a d_l outside {64, 128, 256} is zero-padded to the next of them by the caller, ldo
 * then >= the padded n); bias [n] 16-byte aligned, or NULL.  Same shape and alignment rules as lgnn_gemm_mask_f32. */
int lgnn_gemm_bias_f32(const float* a, int64_t lda, int64_t m_rows, int64_t k, const float* wt_hi,
                       const float* wt_lo, int64_t n, const float* bias, float* out, int64_t ldo,
                       lgnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LGNN_H_ */
