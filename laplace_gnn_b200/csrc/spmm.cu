// CSR SpMM  Y = Â X  (fp32), the HBM-bound core of the GCN forward, its backward and the
// multi-RHS KFAC backward (all C Hessian-sqrt columns of one layer as extra feature columns).
//
// Mapping: one warp per (row, 512-float column chunk).  The 32 lanes first fetch 32 (col, val)
// pairs with one coalesced load each, then walk them with warp shuffles; for every neighbour the
// warp issues VEC 128-bit loads per lane (a contiguous VEC*512-byte piece of the source row), and
// UNROLL neighbours are in flight before the first FMA, i.e. UNROLL*VEC independent 16-byte
// requests per lane — enough outstanding bytes per SM to cover HBM latency at <= 50 % occupancy.
// Warps of one block work on the same rows' adjacent chunks, so (col, val) hit L1 and the pieces
// of a gathered row that different warps touch are adjacent in DRAM.
//
// Algorithmic bytes per launch:  nnz*(4+4) + (n_rows+1)*8 + nnz*d*4 + n_rows*d*4.
#include <limits.h>

#include "common.cuh"
#include "spmm_internal.cuh"

namespace lgnn {

constexpr int SPMM_THREADS = 256;

__global__ void spmm_zero_hub_rows_kernel(int64_t n_rows, const int64_t* __restrict__ rowptr, int64_t hub_len,
                                          float* __restrict__ y, int64_t ldy, int d4);

// acc[t] += sum_{k in [beg, end)} val[k] * X[col[k], chunk columns]  — the warp-wide gather loop shared by
// the row kernel and the hub-segment kernel.
template <int VEC, int UNROLL>
__device__ __forceinline__ void gather_range(int64_t beg, int64_t end, const int32_t* __restrict__ col,
                                             const float* __restrict__ val, const float* __restrict__ xb,
                                             int64_t ldx, const bool (&act)[VEC], float4 (&acc)[VEC], int lane) {
  for (int64_t k0 = beg; k0 < end; k0 += 32) {
    int cnt = (int)((end - k0) < 32 ? (end - k0) : 32);
    int32_t my_c = 0;
    float my_v = 0.f;
    if (lane < cnt) {
      my_c = __ldg(col + k0 + lane);
      my_v = __ldg(val + k0 + lane);
    }
    int j = 0;
    for (; j + UNROLL <= cnt; j += UNROLL) {
      float4 buf[UNROLL][VEC];
      float vv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        int32_t c = __shfl_sync(0xffffffffu, my_c, j + u);
        vv[u] = __shfl_sync(0xffffffffu, my_v, j + u);
        const float* src = xb + (int64_t)c * ldx;
        // an explicit zero (lgnn_mask_edge_values: source row known to be all zero) costs no gather;
        // (c, v) are warp-uniform, so is the branch
#pragma unroll
        for (int t = 0; t < VEC; ++t)
          if (act[t]) buf[u][t] = vv[u] != 0.f ? ldg_f4(src + 128 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int t = 0; t < VEC; ++t)
          if (act[t]) fma4(acc[t], vv[u], buf[u][t]);
    }
    for (; j < cnt; ++j) {
      int32_t c = __shfl_sync(0xffffffffu, my_c, j);
      float v = __shfl_sync(0xffffffffu, my_v, j);
      const float* src = xb + (int64_t)c * ldx;
      if (v != 0.f) {
#pragma unroll
        for (int t = 0; t < VEC; ++t)
          if (act[t]) fma4(acc[t], v, ldg_f4(src + 128 * t));
      }
    }
  }
}

// Rows longer than LDG_HUB_LEN non-zeros are left to spmm_hub_kernel: one warp walking a 100 k-entry
// hub row of a power-law graph is a 10 ms tail (measured 29-54 % of the HBM roofline on R-MAT).
constexpr int64_t LDG_HUB_LEN = 4096;
constexpr int LDG_HUB_SEG = 2048;   // non-zeros per block of the hub kernel (8 warps x 256)

// VEC float4 per lane per chunk (chunk = VEC*128 floats); d4 = d/4.
template <int VEC, int UNROLL>
__global__ void __launch_bounds__(SPMM_THREADS, 2) spmm_vec_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
    int64_t ldy, int d4, int n_chunks, int64_t hub_len, int flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
  const int64_t row = warp / n_chunks;
  if (row >= n_rows) return;
  const int chunk = (int)(warp - row * n_chunks);
  const int c4 = chunk * (VEC * 32) + lane;  // first float4 column of this lane

  bool act[VEC];
#pragma unroll
  for (int t = 0; t < VEC; ++t) act[t] = (c4 + 32 * t) < d4;

  float4 acc[VEC];
#pragma unroll
  for (int t = 0; t < VEC; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  const float* xb = x + (int64_t)c4 * 4;
  const int64_t beg = rowptr[row], end = rowptr[row + 1];
  if (end - beg > hub_len) return;   // hub row: spmm_hub_kernel adds it up in segments
  gather_range<VEC, UNROLL>(beg, end, col, val, xb, ldx, act, acc, lane);
  float* yb = y + row * ldy + (int64_t)c4 * 4;
#pragma unroll
  for (int t = 0; t < VEC; ++t) {
    if (!act[t]) continue;
    float4 a = acc[t];
    if (flags & LGNN_SPMM_RELU) {
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    }
    *reinterpret_cast<float4*>(yb + 128 * t) = a;
  }
}

// Narrow rows (d <= 64 floats, e.g. the 48-float logits of the products shape): with one float4 per lane only d/4 of
// the 32 lanes would work (12 at d = 48: the forward logits SpMM ran at 0.25 of the copy rate).  The warp is cut into
// SPLIT groups of W = 32 / SPLIT lanes; group s takes the neighbours j = s (mod SPLIT) of every batch of 32, so
// SPLIT source rows are in flight per step and every lane that has a column works; the groups' partial sums are
// added with xor-shuffles at the end (the summation order therefore differs from the other kernels' — rounding
// level, far inside the 1e-5 SpMM tolerance).  Hub rows are left to spmm_hub_kernel like in spmm_vec_kernel.
template <int SPLIT, int UNROLL>
__global__ void __launch_bounds__(SPMM_THREADS, 3) spmm_narrow_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
    int64_t ldy, int d4, int64_t hub_len, int flags) {
  constexpr int W = 32 / SPLIT;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int sub = lane / W, l = lane - sub * W;
  const bool act = l < d4;
  const float* xb = x + 4 * l;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t beg = rowptr[row], end = rowptr[row + 1];
  if (end - beg > hub_len) return;   // hub row: spmm_hub_kernel adds it up in segments
  for (int64_t k0 = beg; k0 < end; k0 += 32) {
    const int cnt = (int)((end - k0) < 32 ? (end - k0) : 32);
    int32_t my_c = 0;
    float my_v = 0.f;                 // lanes beyond the batch carry a zero weight: their group step loads nothing
    if (lane < cnt) {
      my_c = __ldg(col + k0 + lane);
      my_v = __ldg(val + k0 + lane);
    }
    for (int j = 0; j < cnt; j += SPLIT * UNROLL) {
      float4 buf[UNROLL];
      float vv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int src_lane = (j + u * SPLIT + sub) & 31;
        const int32_t c = __shfl_sync(0xffffffffu, my_c, src_lane);
        vv[u] = __shfl_sync(0xffffffffu, my_v, src_lane);
        if (j + u * SPLIT + sub >= 32) vv[u] = 0.f;
        buf[u] = (act && vv[u] != 0.f) ? ldg_f4(xb + (int64_t)c * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) fma4(acc, vv[u], buf[u]);
    }
  }
#pragma unroll
  for (int off = W; off < 32; off <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
  }
  if (sub == 0 && act) {
    if (flags & LGNN_SPMM_RELU) {
      acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    }
    *reinterpret_cast<float4*>(y + row * ldy + 4 * l) = acc;
  }
}

// Hub rows of the warp-per-row path.  Block b owns the non-zeros [b*SEG, (b+1)*SEG); for every hub row
// intersecting that range its 8 warps split the intersection evenly and add their partial sums onto
// the row (cleared beforehand by spmm_zero_hub_rows_kernel) with red.global.add.v4.f32.  Blocks
// whose range touches no hub row — all of them on a uniform graph — leave after one ballot.
template <int VEC, int UNROLL>
__global__ void __launch_bounds__(SPMM_THREADS, 2) spmm_hub_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
    int64_t ldy, int d4) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t nnz = __ldg(rowptr + n_rows);
  const int64_t t_beg = (int64_t)blockIdx.x * LDG_HUB_SEG;
  if (t_beg >= nnz) return;
  const int64_t t_end = (t_beg + LDG_HUB_SEG < nnz) ? t_beg + LDG_HUB_SEG : nnz;
  // last row whose first non-zero is at or before t_beg
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid) <= t_beg) lo = mid + 1; else hi = mid;
  }
  const int c4 = blockIdx.y * (VEC * 32) + lane;
  bool act[VEC];
#pragma unroll
  for (int t = 0; t < VEC; ++t) act[t] = (c4 + 32 * t) < d4;
  const float* xb = x + (int64_t)c4 * 4;
  for (int64_t row = lo - 1; row < n_rows; ++row) {
    const int64_t r_beg = __ldg(rowptr + row);
    if (r_beg >= t_end) break;
    const int64_t r_end = __ldg(rowptr + row + 1);
    if (r_end - r_beg <= LDG_HUB_LEN) continue;
    const int64_t s_beg = r_beg > t_beg ? r_beg : t_beg;
    const int64_t s_end = r_end < t_end ? r_end : t_end;
    const int64_t per_warp = (s_end - s_beg + SPMM_THREADS / 32 - 1) / (SPMM_THREADS / 32);
    const int64_t my_beg = s_beg + w * per_warp;
    int64_t my_end = my_beg + per_warp;
    if (my_end > s_end) my_end = s_end;
    if (my_beg >= my_end) continue;
    float4 acc[VEC];
#pragma unroll
    for (int t = 0; t < VEC; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_range<VEC, UNROLL>(my_beg, my_end, col, val, xb, ldx, act, acc, lane);
    float* yb = y + row * ldy + (int64_t)c4 * 4;
#pragma unroll
    for (int t = 0; t < VEC; ++t)
      if (act[t])
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yb + 128 * t), "f"(acc[t].x),
                     "f"(acc[t].y), "f"(acc[t].z), "f"(acc[t].w)
                     : "memory");
  }
}

// y[row, :] = max(y[row, :], 0) for hub rows (their relu cannot be fused: the sum arrives in pieces)
__global__ void spmm_relu_hub_rows_kernel(int64_t n_rows, const int64_t* __restrict__ rowptr, int64_t hub_len,
                                          float* __restrict__ y, int64_t ldy, int d4) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  if (__ldg(rowptr + row + 1) - __ldg(rowptr + row) <= hub_len) return;
  float4* dst = reinterpret_cast<float4*>(y + row * ldy);
  for (int c = lane; c < d4; c += 32) {
    float4 a = dst[c];
    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    dst[c] = a;
  }
}

// scalar path: any d / ld / alignment.  One warp per (row, 128-column chunk).
__global__ void __launch_bounds__(SPMM_THREADS) spmm_scalar_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
    int64_t ldy, int64_t d, int n_chunks, int flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
  const int64_t row = warp / n_chunks;
  if (row >= n_rows) return;
  const int chunk = (int)(warp - row * n_chunks);
  const int64_t c0 = (int64_t)chunk * 128 + lane;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t beg = rowptr[row], end = rowptr[row + 1];
  for (int64_t k0 = beg; k0 < end; k0 += 32) {
    int cnt = (int)((end - k0) < 32 ? (end - k0) : 32);
    int32_t my_c = 0;
    float my_v = 0.f;
    if (lane < cnt) {
      my_c = __ldg(col + k0 + lane);
      my_v = __ldg(val + k0 + lane);
    }
    for (int j = 0; j < cnt; ++j) {
      int32_t c = __shfl_sync(0xffffffffu, my_c, j);
      float v = __shfl_sync(0xffffffffu, my_v, j);
      const float* src = x + (int64_t)c * ldx + c0;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (c0 + 32 * t < d) acc[t] = fmaf(v, __ldg(src + 32 * t), acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (c0 + 32 * t < d) {
      float a = acc[t];
      if (flags & LGNN_SPMM_RELU) a = fmaxf(a, 0.f);
      y[row * ldy + c0 + 32 * t] = a;
    }
}


// ------------------------------------------------------------------------------------------
// Wide SpMM (d >= 512): bulk-async gather through a shared-memory ring.
//
// One CTA owns a block of consecutive rows holding ~BULK_NNZ_PER_CTA non-zeros (found by a binary
// search in rowptr, so CTAs are nnz-balanced) and one column range of at most 4096 floats (16 KB stages).  A
// producer warp walks the (col, val) stream of those rows; every lane issues, for its neighbour, ONE
// cp.async.bulk (the TMA engine's 1-D copy) of the whole <= 16 KB piece of the source row into a
// ring stage, completion signalled on the stage's mbarrier.  Up to ~190 KB of gathered rows are in
// flight per SM without holding a single register, which is what it takes to keep HBM busy with
// dependent random accesses.  Eight consumer warps wait on the stage, read it back conflict-free
// (consecutive float4 per lane), FMA into register accumulators, and release the stage; at the end
// of a row they store the row of Y (relu fused).
// Rows longer than BULK_HUB_FACTOR CTA budgets ("hubs" of power-law graphs) are split at the CTA
// budget boundaries: every CTA adds its piece of the row with red.global.add.v4.f32 onto a row that
// spmm_zero_hub_rows_kernel cleared beforehand.  All other rows belong entirely to the CTA in whose
// budget they start and are written with plain stores in a fixed summation order (bit-reproducible,
// bit-equal to the warp-per-row kernel).  With a fused relu the split is off (relu needs the full sum).
constexpr int BULK_HUB_FACTOR = 4;

__global__ void spmm_zero_hub_rows_kernel(int64_t n_rows, const int64_t* __restrict__ rowptr, int64_t hub_len,
                                          float* __restrict__ y, int64_t ldy, int d4) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  if (__ldg(rowptr + row + 1) - __ldg(rowptr + row) <= hub_len) return;
  float4* dst = reinterpret_cast<float4*>(y + row * ldy);
  for (int c = lane; c < d4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int VPT>  // float4 accumulators per consumer thread; column range <= VPT*256 float4
__global__ void __launch_bounds__(BULK_THREADS, 1) spmm_bulk_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
    int64_t ldy, int d4, int range4, int stages, int stage_bytes, int64_t nnz_per_cta, int64_t hub_len,
    int flags) {
  extern __shared__ uint8_t bulk_smem_[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bulk_smem_) + 127) & ~(uintptr_t)127);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)stages * stage_bytes);
  uint64_t* empty = full + BULK_MAX_STAGES;
  float* meta = reinterpret_cast<float*>(empty + BULK_MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c4_0 = blockIdx.x * range4;                       // first float4 column of this range
  const int w4 = min(range4, d4 - c4_0);                      // float4 columns in this range
  const uint32_t bytes = (uint32_t)w4 * 16u;

  const int64_t nnz = __ldg(rowptr + n_rows);
  const bool last_cta = blockIdx.y == gridDim.y - 1;
  const int64_t t_beg = (int64_t)blockIdx.y * nnz_per_cta;
  if (blockIdx.y > 0 && t_beg >= nnz) return;
  const int64_t t_end = (last_cta || t_beg + nnz_per_cta >= nnz) ? nnz : t_beg + nnz_per_cta;
  const int64_t row_beg = blockIdx.y == 0 ? 0 : row_lower_bound(rowptr, n_rows, t_beg);
  const int64_t row_end = (last_cta || t_end >= nnz) ? n_rows : row_lower_bound(rowptr, n_rows, t_end);
  // piece of a hub row that started in an earlier CTA's budget and reaches into this one
  const int64_t main_beg = __ldg(rowptr + row_beg);
  const bool has_prefix = blockIdx.y > 0 && row_beg > 0 && main_beg > t_beg &&
                          (main_beg - __ldg(rowptr + row_beg - 1)) > hub_len;
  const int64_t pre_end = main_beg < t_end ? main_beg : t_end;
  // a hub row that starts in this budget (necessarily its last row) stops at the budget boundary
  int64_t s_end = has_prefix ? pre_end : main_beg;
  if (row_end > row_beg) {
    s_end = __ldg(rowptr + row_end);
    const int64_t last_beg = __ldg(rowptr + row_end - 1);
    if (s_end - last_beg > hub_len && s_end > t_end) s_end = t_end;
  }
  const int64_t s_beg = has_prefix ? t_beg : main_beg;
  if (row_beg >= row_end && !has_prefix) return;

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      bulk_mbar_init(smem_addr(&full[i]), 1);
      bulk_mbar_init(smem_addr(&empty[i]), BULK_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == BULK_CONSUMERS / 32) {
    // ===================================================================== producer warp
    const float* xb = x + (int64_t)c4_0 * 4;
    for (int64_t k0 = s_beg; k0 < s_end; k0 += 32) {
      const int64_t k = k0 + lane;
      const bool valid = k < s_end;
      int32_t c = 0;
      float v = 0.f;
      if (valid) {
        c = __ldg(col + k);
        v = __ldg(val + k);
      }
      const int64_t it = k - s_beg;
      const int st = (int)(it % stages);
      const uint32_t ph = (uint32_t)((it / stages) & 1);
      // lanes whose stages coincide (stages < 32) go in successive sub-batches
      for (int sub = 0; sub * stages < 32; ++sub) {
        if (valid && lane / stages == sub) {
          bulk_mbar_wait(smem_addr(&empty[st]), ph ^ 1u);
          meta[st] = v;
          const uint32_t bar = smem_addr(&full[st]);
          bulk_mbar_expect_tx(bar, bytes);
          bulk_g2s(smem_addr(ring + (size_t)st * stage_bytes), xb + (int64_t)c * ldx, bytes, bar);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== consumer warps
    const int t = threadIdx.x;
    bool act[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) act[j] = (t + j * BULK_CONSUMERS) < w4;
    int64_t it = 0;
    // one segment = `cnt` consecutive ring stages accumulated into row `row`
    auto segment = [&](int64_t row, int64_t cnt, bool atomic) {
      float4 acc[VPT];
#pragma unroll
      for (int j = 0; j < VPT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t e = 0; e < cnt; ++e, ++it) {
        const int st = (int)(it % stages);
        const uint32_t ph = (uint32_t)((it / stages) & 1);
        bulk_mbar_wait(smem_addr(&full[st]), ph);
        const float v = meta[st];
        const float4* src = reinterpret_cast<const float4*>(ring + (size_t)st * stage_bytes) + t;
        float4 b[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j)
          if (act[j]) b[j] = src[j * BULK_CONSUMERS];
#pragma unroll
        for (int j = 0; j < VPT; ++j)
          if (act[j]) fma4(acc[j], v, b[j]);
        __syncwarp();
        if (lane == 0) bulk_mbar_arrive(smem_addr(&empty[st]));
      }
      float4* dst = reinterpret_cast<float4*>(y + row * ldy) + c4_0 + t;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        if (!act[j]) continue;
        float4 a = acc[j];
        if (atomic) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j * BULK_CONSUMERS), "f"(a.x),
                       "f"(a.y), "f"(a.z), "f"(a.w)
                       : "memory");
          continue;
        }
        if (flags & LGNN_SPMM_RELU) {
          a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
        }
        dst[j * BULK_CONSUMERS] = a;
      }
    };
    if (has_prefix) segment(row_beg - 1, pre_end - t_beg, true);
    int64_t k_row = main_beg;
    for (int64_t row = row_beg; row < row_end; ++row) {
      const int64_t k_next = __ldg(rowptr + row + 1);
      const bool hub = (k_next - k_row) > hub_len;
      const int64_t k_stop = (hub && k_next > s_end) ? s_end : k_next;
      segment(row, k_stop - k_row, hub);
      k_row = k_next;
    }
  }
}

template <int VPT>
static int launch_bulk(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                       const float* x, int64_t ldx, float* y, int64_t ldy, int d4, int n_ranges,
                       int range4, int64_t nnz_hint, int flags, cudaStream_t st) {
  const int stage_bytes = (range4 * 16 + 127) / 128 * 128;
  int stages = BULK_SMEM_RING / stage_bytes;
  if (stages > BULK_MAX_STAGES) stages = BULK_MAX_STAGES;
  const size_t smem = 128 + (size_t)stages * stage_bytes + 2 * BULK_MAX_STAGES * sizeof(uint64_t) +
                      BULK_MAX_STAGES * sizeof(float);
  LGNN_CUDA_TRY(cudaFuncSetAttribute(spmm_bulk_kernel<VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks_y = (nnz_hint + BULK_NNZ_PER_CTA - 1) / BULK_NNZ_PER_CTA;
  if (blocks_y < 1) blocks_y = 1;
  int64_t per_cta = BULK_NNZ_PER_CTA;
  if (blocks_y > 65535) {  // grid.y limit: fatter CTAs
    per_cta = (nnz_hint + 65534) / 65535;
    blocks_y = (nnz_hint + per_cta - 1) / per_cta;
  }
  dim3 grid((unsigned)n_ranges, (unsigned)blocks_y);
  int64_t hub_len = per_cta * BULK_HUB_FACTOR;
  if (flags & (LGNN_SPMM_RELU | LGNN_SPMM_NO_HUB_ROWS)) {
    hub_len = INT64_MAX;   // relu needs the complete row sum / the caller vouches for short rows: no split
  } else {
    const int64_t zb = (n_rows * 32 + 255) / 256;
    if (zb > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm: grid too large");
    spmm_zero_hub_rows_kernel<<<(unsigned)zb, 256, 0, st>>>(n_rows, rowptr, hub_len, y, ldy, d4);
    LGNN_LAUNCH_CHECK("spmm_zero_hub_rows_kernel");
  }
  spmm_bulk_kernel<VPT><<<grid, BULK_THREADS, smem, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4,
                                                         range4, stages, stage_bytes, per_cta, hub_len, flags);
  LGNN_LAUNCH_CHECK("spmm_bulk_kernel");
  return LGNN_OK;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" int lgnn_spmm_f32(int64_t n_rows, int64_t nnz, const int64_t* rowptr, const int32_t* col,
                             const float* val, const float* x, int64_t ldx, float* y, int64_t ldy,
                             int64_t d, int flags, lgnn_stream_t stream) {
  if (n_rows < 0 || nnz < 0 || d < 0 || ldx < d || ldy < d) return fail(LGNN_E_BADARG, "spmm: bad shape (n_rows=%lld d=%lld ldx=%lld ldy=%lld)", (long long)n_rows, (long long)d, (long long)ldx, (long long)ldy);
  if (n_rows == 0 || d == 0) return LGNN_OK;
  if (!rowptr || !x || !y) return fail(LGNN_E_BADARG, "spmm: null pointer");
  if (nnz > 0 && (!col || !val)) return fail(LGNN_E_BADARG, "spmm: null col / val");
  if (d > (int64_t)1 << 24) return fail(LGNN_E_UNSUPPORTED, "spmm: d too large");
  cudaStream_t st = as_stream(stream);
  const bool vec_ok = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  const int warps_per_block = SPMM_THREADS / 32;
  const int epi = flags & LGNN_SPMM_RELU;
  // measured on B200 (profiles/): both kernels sit at the HBM roofline for wide slabs; with small ring
  // stages the bulk kernel's per-stage hand-off dominates and the warp-per-row kernel is faster
  const int d4_all = (int)(d / 4);
  const int bulk_ranges = (d4_all + 1023) / 1024;                       // column ranges of <= 4096 floats
  const int bulk_range4 = bulk_ranges > 0 ? (d4_all + bulk_ranges - 1) / bulk_ranges : 0;
  bool bulk = vec_ok && bulk_range4 >= 896;   // paired lab runs (profiles/r1f_spmm_lab.txt): the ring wins from ~14 KB stages up
  if (flags & LGNN_SPMM_FORCE_LDG) bulk = false;
  if (flags & LGNN_SPMM_FORCE_BULK) {
    if (!vec_ok) return fail(LGNN_E_ALIGN, "spmm: the bulk path needs 16-byte aligned x / y and d, ldx, ldy %% 4 == 0");
    bulk = true;
  }
  if (bulk) {
    const int d4 = d4_all;
    const int n_ranges = bulk_ranges;
    const int range4 = (d4 + n_ranges - 1) / n_ranges;
    const int vpt = (range4 + BULK_CONSUMERS - 1) / BULK_CONSUMERS;
    switch (vpt) {
      case 1: return launch_bulk<1>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, n_ranges, range4, nnz, epi | (flags & LGNN_SPMM_NO_HUB_ROWS), st);
      case 2: return launch_bulk<2>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, n_ranges, range4, nnz, epi | (flags & LGNN_SPMM_NO_HUB_ROWS), st);
      case 3: return launch_bulk<3>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, n_ranges, range4, nnz, epi | (flags & LGNN_SPMM_NO_HUB_ROWS), st);
      default: return launch_bulk<4>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, n_ranges, range4, nnz, epi | (flags & LGNN_SPMM_NO_HUB_ROWS), st);
    }
  }
  if (vec_ok) {
    int d4 = (int)(d / 4);
    const int64_t zb = (n_rows * 32 + 255) / 256;                      // one warp per row
    // hub kernel: one block per 2048 non-zeros; skipped when the caller vouches for short rows
    const int64_t hub_blocks = (flags & LGNN_SPMM_NO_HUB_ROWS) ? 0 : (nnz + LDG_HUB_SEG - 1) / LDG_HUB_SEG;
    const int64_t hub_len = hub_blocks > 0 ? LDG_HUB_LEN : INT64_MAX;
    if (zb > 0x7fffffffLL || hub_blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm: grid too large");
    if (hub_blocks > 0) {
      spmm_zero_hub_rows_kernel<<<(unsigned)zb, 256, 0, st>>>(n_rows, rowptr, LDG_HUB_LEN, y, ldy, d4);
      LGNN_LAUNCH_CHECK("spmm_zero_hub_rows_kernel");
    }
    if (d4 <= 16 && !(flags & LGNN_SPMM_FORCE_LDG)) {
      const unsigned nb = (unsigned)((n_rows + warps_per_block - 1) / warps_per_block);
      if (d4 <= 4) spmm_narrow_kernel<8, 2><<<nb, SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, hub_len, epi);
      else if (d4 <= 8) spmm_narrow_kernel<4, 4><<<nb, SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, hub_len, epi);
      else spmm_narrow_kernel<2, 4><<<nb, SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4, hub_len, epi);
      if (hub_blocks > 0)
        spmm_hub_kernel<1, 8><<<dim3((unsigned)hub_blocks, 1), SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4);
    } else if (d4 <= 32) {
      int64_t warps = n_rows;
      spmm_vec_kernel<1, 8><<<(unsigned)((warps + warps_per_block - 1) / warps_per_block), SPMM_THREADS, 0, st>>>(
          n_rows, rowptr, col, val, x, ldx, y, ldy, d4, 1, hub_len, epi);
      if (hub_blocks > 0)
        spmm_hub_kernel<1, 8><<<dim3((unsigned)hub_blocks, 1), SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4);
    } else if (d4 <= 64) {
      int64_t warps = n_rows;
      spmm_vec_kernel<2, 4><<<(unsigned)((warps + warps_per_block - 1) / warps_per_block), SPMM_THREADS, 0, st>>>(
          n_rows, rowptr, col, val, x, ldx, y, ldy, d4, 1, hub_len, epi);
      if (hub_blocks > 0)
        spmm_hub_kernel<2, 4><<<dim3((unsigned)hub_blocks, 1), SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy, d4);
    } else {
      // one warp per (row, chunk of VEC*32 float4): pick the chunk width that leaves the fewest idle lane
      // slots in the last chunk (d = 576: one chunk of 160 float4 instead of 128 + a warp that walks the
      // whole row for 16); ties go to VEC = 4, the width the wide slabs were tuned on
      static const int cand[4] = {4, 5, 6, 3};
      int vec = 4, best = INT32_MAX;
      for (int i = 0; i < 4; ++i) {
        const int slots = (d4 + cand[i] * 32 - 1) / (cand[i] * 32) * cand[i] * 32;
        if (slots < best) { best = slots; vec = cand[i]; }
      }
      const int n_chunks = (d4 + vec * 32 - 1) / (vec * 32);
      int64_t warps = n_rows * n_chunks;
      int64_t blocks = (warps + warps_per_block - 1) / warps_per_block;
      if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm: grid too large");
      if (n_chunks > 65535) return fail(LGNN_E_UNSUPPORTED, "spmm: d too large");
#define LGNN_VEC_LAUNCH(VEC_)                                                                                    \
  do {                                                                                                           \
    spmm_vec_kernel<VEC_, 2><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y, ldy,  \
                                                                        d4, n_chunks, hub_len, epi);             \
    if (hub_blocks > 0)                                                                                          \
      spmm_hub_kernel<VEC_, 2><<<dim3((unsigned)hub_blocks, (unsigned)n_chunks), SPMM_THREADS, 0, st>>>(           \
          n_rows, rowptr, col, val, x, ldx, y, ldy, d4);                                                         \
  } while (0)
      switch (vec) {
        case 3: LGNN_VEC_LAUNCH(3); break;
        case 5: LGNN_VEC_LAUNCH(5); break;
        case 6: LGNN_VEC_LAUNCH(6); break;
        default: LGNN_VEC_LAUNCH(4); break;
      }
#undef LGNN_VEC_LAUNCH
    }
    if (hub_blocks > 0 && epi) {
      spmm_relu_hub_rows_kernel<<<(unsigned)zb, 256, 0, st>>>(n_rows, rowptr, LDG_HUB_LEN, y, ldy, d4);
      LGNN_LAUNCH_CHECK("spmm_relu_hub_rows_kernel");
    }
    LGNN_LAUNCH_CHECK("spmm_vec_kernel");
  } else {
    int n_chunks = (int)((d + 127) / 128);
    int64_t warps = n_rows * n_chunks;
    int64_t blocks = (warps + warps_per_block - 1) / warps_per_block;
    if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm: grid too large");
    spmm_scalar_kernel<<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y,
                                                                 ldy, d, n_chunks, epi);
    LGNN_LAUNCH_CHECK("spmm_scalar_kernel");
  }
  return LGNN_OK;
}
