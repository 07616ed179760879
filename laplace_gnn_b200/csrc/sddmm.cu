// Sampled dense-dense products on a list of (row, column) pairs:
//     out[e] (+)= U[rows[e], 0:d] . V[cols[e], 0:d]
// — the edge-wise dot products of the adjacency gradient (laplace_gnn_b200/structure.py): wherever the
// path computes Y = Â X, the adjoint of Â[i, j] is Ybar[i, :] . X[j, :].  The reference gets the same
// numbers as a DENSE N x N outer product from autograd (gnn/marglik_training.py:213,
// `neg_marglik.backward()` into `model.adj.grad`); here only the requested entries are evaluated.
//
// One warp per pair: the lanes stride over the row with 128-bit loads (4 independent loads of U and of V
// in flight per lane), FMA into 4 partial sums, one shuffle reduction per pair.  Consecutive pairs of
// a CSR pattern share their U row (L1 / L2 hits); the V rows are the random gathers — HBM-bound with
// the SpMM's byte count: nnz * d * 4 gathered + the (row, column) stream.
#include "common.cuh"
#include "spmm_internal.cuh"

namespace lgnn {

constexpr int SDDMM_THREADS = 256;

__global__ void __launch_bounds__(SDDMM_THREADS) sddmm_vec_kernel(
    int64_t n_pairs, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
    const float* __restrict__ u, int64_t ldu, const float* __restrict__ v, int64_t ldv, int d4,
    float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * SDDMM_THREADS + threadIdx.x) >> 5;
  if (e >= n_pairs) return;
  const float* pu = u + (int64_t)__ldg(rows + e) * ldu;
  const float* pv = v + (int64_t)__ldg(cols + e) * ldv;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = lane;
  for (; c + 96 < d4; c += 128) {
    const float4 a0 = ldg_f4(pu + 4 * c), a1 = ldg_f4(pu + 4 * (c + 32));
    const float4 a2 = ldg_f4(pu + 4 * (c + 64)), a3 = ldg_f4(pu + 4 * (c + 96));
    const float4 b0 = ldg_f4(pv + 4 * c), b1 = ldg_f4(pv + 4 * (c + 32));
    const float4 b2 = ldg_f4(pv + 4 * (c + 64)), b3 = ldg_f4(pv + 4 * (c + 96));
    s0 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s0))));
    s1 = fmaf(a1.x, b1.x, fmaf(a1.y, b1.y, fmaf(a1.z, b1.z, fmaf(a1.w, b1.w, s1))));
    s2 = fmaf(a2.x, b2.x, fmaf(a2.y, b2.y, fmaf(a2.z, b2.z, fmaf(a2.w, b2.w, s2))));
    s3 = fmaf(a3.x, b3.x, fmaf(a3.y, b3.y, fmaf(a3.z, b3.z, fmaf(a3.w, b3.w, s3))));
  }
  for (; c < d4; c += 32) {
    const float4 a0 = ldg_f4(pu + 4 * c);
    const float4 b0 = ldg_f4(pv + 4 * c);
    s0 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s0))));
  }
  float s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[e] = accumulate ? out[e] + s : s;
}

__global__ void __launch_bounds__(SDDMM_THREADS) sddmm_scalar_kernel(
    int64_t n_pairs, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
    const float* __restrict__ u, int64_t ldu, const float* __restrict__ v, int64_t ldv, int64_t d,
    float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * SDDMM_THREADS + threadIdx.x) >> 5;
  if (e >= n_pairs) return;
  const float* pu = u + (int64_t)__ldg(rows + e) * ldu;
  const float* pv = v + (int64_t)__ldg(cols + e) * ldv;
  float s = 0.f;
  for (int64_t c = lane; c < d; c += 32) s = fmaf(__ldg(pu + c), __ldg(pv + c), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[e] = accumulate ? out[e] + s : s;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" int lgnn_sddmm_f32(int64_t n_pairs, const int32_t* rows, const int32_t* cols, const float* u,
                              int64_t ldu, const float* v, int64_t ldv, int64_t d, float* out,
                              int accumulate, lgnn_stream_t stream) {
  if (n_pairs < 0 || d < 0 || ldu < d || ldv < d) return fail(LGNN_E_BADARG, "sddmm: bad shape");
  if (n_pairs == 0) return LGNN_OK;
  if (!rows || !cols || !out || (d > 0 && (!u || !v))) return fail(LGNN_E_BADARG, "sddmm: null pointer");
  const int64_t blocks = (n_pairs * 32 + SDDMM_THREADS - 1) / SDDMM_THREADS;
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "sddmm: too many pairs");
  cudaStream_t st = as_stream(stream);
  const bool vec = (d % 4 == 0) && (ldu % 4 == 0) && (ldv % 4 == 0) && (d / 4 <= 0x7fffffffLL) &&
                   ((reinterpret_cast<uintptr_t>(u) & 15) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  if (vec)
    sddmm_vec_kernel<<<(unsigned)blocks, SDDMM_THREADS, 0, st>>>(n_pairs, rows, cols, u, ldu, v, ldv, (int)(d / 4),
                                                                  out, accumulate);
  else
    sddmm_scalar_kernel<<<(unsigned)blocks, SDDMM_THREADS, 0, st>>>(n_pairs, rows, cols, u, ldu, v, ldv, d, out,
                                                                     accumulate);
  LGNN_LAUNCH_CHECK("sddmm_kernel");
  return LGNN_OK;
}
