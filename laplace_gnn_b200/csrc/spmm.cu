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
#include "common.cuh"

namespace lgnn {

constexpr int SPMM_THREADS = 256;

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
  a.x = fmaf(v, x.x, a.x);
  a.y = fmaf(v, x.y, a.y);
  a.z = fmaf(v, x.z, a.z);
  a.w = fmaf(v, x.w, a.w);
}

// VEC float4 per lane per chunk (chunk = VEC*128 floats); d4 = d/4.
template <int VEC, int UNROLL>
__global__ void __launch_bounds__(SPMM_THREADS) spmm_vec_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
    int64_t ldy, int d4, int n_chunks, int flags) {
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
#pragma unroll
        for (int t = 0; t < VEC; ++t)
          if (act[t]) buf[u][t] = ldg_f4(src + 128 * t);
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
#pragma unroll
      for (int t = 0; t < VEC; ++t)
        if (act[t]) fma4(acc[t], v, ldg_f4(src + 128 * t));
    }
  }
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

}  // namespace lgnn

using namespace lgnn;

extern "C" int lgnn_spmm_f32(int64_t n_rows, const int64_t* rowptr, const int32_t* col,
                             const float* val, const float* x, int64_t ldx, float* y, int64_t ldy,
                             int64_t d, int flags, lgnn_stream_t stream) {
  if (n_rows < 0 || d < 0 || ldx < d || ldy < d) return fail(LGNN_E_BADARG, "spmm: bad shape (n_rows=%lld d=%lld ldx=%lld ldy=%lld)", (long long)n_rows, (long long)d, (long long)ldx, (long long)ldy);
  if (n_rows == 0 || d == 0) return LGNN_OK;
  if (!rowptr || !x || !y) return fail(LGNN_E_BADARG, "spmm: null pointer");
  if (d > (int64_t)1 << 24) return fail(LGNN_E_UNSUPPORTED, "spmm: d too large");
  cudaStream_t st = as_stream(stream);
  const bool vec_ok = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  const int warps_per_block = SPMM_THREADS / 32;
  if (vec_ok) {
    int d4 = (int)(d / 4);
    if (d4 <= 32) {
      int64_t warps = n_rows;
      spmm_vec_kernel<1, 8><<<(unsigned)((warps + warps_per_block - 1) / warps_per_block), SPMM_THREADS, 0, st>>>(
          n_rows, rowptr, col, val, x, ldx, y, ldy, d4, 1, flags);
    } else if (d4 <= 64) {
      int64_t warps = n_rows;
      spmm_vec_kernel<2, 4><<<(unsigned)((warps + warps_per_block - 1) / warps_per_block), SPMM_THREADS, 0, st>>>(
          n_rows, rowptr, col, val, x, ldx, y, ldy, d4, 1, flags);
    } else {
      int n_chunks = (d4 + 127) / 128;
      int64_t warps = n_rows * n_chunks;
      int64_t blocks = (warps + warps_per_block - 1) / warps_per_block;
      if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm: grid too large");
      spmm_vec_kernel<4, 2><<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(
          n_rows, rowptr, col, val, x, ldx, y, ldy, d4, n_chunks, flags);
    }
    LGNN_LAUNCH_CHECK("spmm_vec_kernel");
  } else {
    int n_chunks = (int)((d + 127) / 128);
    int64_t warps = n_rows * n_chunks;
    int64_t blocks = (warps + warps_per_block - 1) / warps_per_block;
    if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm: grid too large");
    spmm_scalar_kernel<<<(unsigned)blocks, SPMM_THREADS, 0, st>>>(n_rows, rowptr, col, val, x, ldx, y,
                                                                 ldy, d, n_chunks, flags);
    LGNN_LAUNCH_CHECK("spmm_scalar_kernel");
  }
  return LGNN_OK;
}
