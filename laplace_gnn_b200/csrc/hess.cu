// Loss, softmax Hessian-sqrt right-hand sides and the relu' mask — small fused elementwise kernels.
#include "common.cuh"

namespace lgnn {

constexpr int HESS_THREADS = 256;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// warp per train sample: CE = logsumexp(f) - f_y, summed in double.
__global__ void __launch_bounds__(HESS_THREADS) softmax_ce_sum_kernel(
    const float* __restrict__ logits, int64_t ld, int C, const int64_t* __restrict__ idx,
    const int64_t* __restrict__ y, int64_t m, double* __restrict__ loss,
    unsigned long long* __restrict__ n_correct) {
  __shared__ double sh_loss[HESS_THREADS / 32];
  __shared__ int sh_hit[HESS_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t s = ((int64_t)blockIdx.x * HESS_THREADS + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * HESS_THREADS) >> 5;
  double acc = 0.0;
  int hits = 0;
  for (; s < m; s += stride) {
    const float* f = logits + idx[s] * ld;
    float mx = -INFINITY;
    int arg = 0;
    for (int k = lane; k < C; k += 32) {
      float v = f[k];
      if (v > mx) { mx = v; arg = k; }
    }
    // warp arg-max (lowest index wins ties, like torch.argmax on CPU)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, mx, o);
      int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float se = 0.f;
    for (int k = lane; k < C; k += 32) se += expf(f[k] - mx);
    se = warp_sum(se);
    int64_t label = y[s];
    if (lane == 0) {
      acc += (double)(logf(se) + mx - f[label]);
      hits += (arg == (int)label);
    }
  }
  if (lane == 0) { sh_loss[w] = acc; sh_hit[w] = hits; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int h = 0;
    for (int i = 0; i < HESS_THREADS / 32; ++i) { t += sh_loss[i]; h += sh_hit[i]; }
    atomicAdd(loss, t);
    if (n_correct != nullptr && h) atomicAdd(n_correct, (unsigned long long)h);
  }
}

// warp per train sample.  Shared memory per warp: p[C], fc[C] (= f - fbar).
__global__ void __launch_bounds__(HESS_THREADS) hess_rhs_kernel(
    const float* __restrict__ logits, int64_t ld, int C, const int64_t* __restrict__ idx, int64_t m,
    int c0, int ncols, int ldc, int64_t ld_delta, int mode, float* __restrict__ delta) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* p = sh + (size_t)w * 2 * C;
  float* fc = p + C;
  int64_t s = ((int64_t)blockIdx.x * HESS_THREADS + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * HESS_THREADS) >> 5;
  for (; s < m; s += stride) {
    const int64_t node = idx[s];
    const float* f = logits + node * ld;
    float mx = -INFINITY;
    for (int k = lane; k < C; k += 32) mx = fmaxf(mx, f[k]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < C; k += 32) {
      float e = expf(f[k] - mx);
      p[k] = e;
      se += e;
    }
    se = warp_sum(se);
    float fbar = 0.f;
    for (int k = lane; k < C; k += 32) {
      float pk = p[k] / se;
      p[k] = pk;
      fbar = fmaf(pk, f[k], fbar);
    }
    fbar = warp_sum(fbar);
    for (int k = lane; k < C; k += 32) fc[k] = f[k] - fbar;
    __syncwarp();
    float* out = delta + node * ld_delta;
    for (int g = 0; g < ncols; ++g) {
      const int c = c0 + g;
      const float pc = p[c];
      const float spc = sqrtf(pc);
      const float a = (mode == LGNN_HESS_REFERENCE) ? fmaf(0.5f, fc[c], 1.0f) : 1.0f;
      for (int k = lane; k < C; k += 32) {
        float e = (k == c ? 1.0f : 0.0f) - p[k];
        float v = e * a;
        if (mode == LGNN_HESS_REFERENCE) v -= p[k] * fc[k];
        atomicAdd(out + (int64_t)g * ldc + k, spc * v);
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(HESS_THREADS) relu_mask_mul_kernel(
    const float* __restrict__ in, int64_t ldi, const float* __restrict__ act, int64_t lda,
    float* __restrict__ out, int64_t ldo, int64_t n_rows, int group, int64_t d, int vec) {
  // one thread per (row of in, float4 / float column); rows of in = n_rows * group
  const int64_t dv = vec ? d / 4 : d;
  const int64_t total = n_rows * group * dv;
  int64_t i = (int64_t)blockIdx.x * HESS_THREADS + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * HESS_THREADS;
  for (; i < total; i += stride) {
    int64_t r = i / dv;
    int64_t j = i - r * dv;
    int64_t node = r / group;
    if (vec) {
      float4 a = __ldg(reinterpret_cast<const float4*>(act + node * lda) + j);
      float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ldi) + j);
      v.x = a.x > 0.f ? v.x : 0.f;
      v.y = a.y > 0.f ? v.y : 0.f;
      v.z = a.z > 0.f ? v.z : 0.f;
      v.w = a.w > 0.f ? v.w : 0.f;
      reinterpret_cast<float4*>(out + r * ldo)[j] = v;
    } else {
      float a = act[node * lda + j];
      float v = in[r * ldi + j];
      out[r * ldo + j] = a > 0.f ? v : 0.f;
    }
  }
}

// out[k] = keep[col[k]] ? val[k] : 0
__global__ void __launch_bounds__(HESS_THREADS) mask_edge_values_kernel(
    int64_t nnz, const int32_t* __restrict__ col, const float* __restrict__ val,
    const uint8_t* __restrict__ keep, float* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * HESS_THREADS + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * HESS_THREADS;
  for (; k < nnz; k += stride) out[k] = keep[__ldg(col + k)] ? __ldg(val + k) : 0.f;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" {

int lgnn_softmax_ce_sum(const float* logits, int64_t ld, int32_t C, const int64_t* idx,
                        const int64_t* y, int64_t m, double* loss, int64_t* n_correct,
                        lgnn_stream_t stream) {
  if (!logits || !loss || C < 1 || ld < C || m < 0) return fail(LGNN_E_BADARG, "softmax_ce_sum: bad argument");
  if (m == 0) return LGNN_OK;
  if (!idx || !y) return fail(LGNN_E_BADARG, "softmax_ce_sum: null idx / y");
  int64_t blocks = (m + HESS_THREADS / 32 - 1) / (HESS_THREADS / 32);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  softmax_ce_sum_kernel<<<(unsigned)blocks, HESS_THREADS, 0, as_stream(stream)>>>(
      logits, ld, C, idx, y, m, loss, reinterpret_cast<unsigned long long*>(n_correct));
  LGNN_LAUNCH_CHECK("softmax_ce_sum_kernel");
  return LGNN_OK;
}

int lgnn_hess_rhs_f32(const float* logits, int64_t ld, int32_t C, const int64_t* idx, int64_t m,
                      int32_t c0, int32_t ncols, int32_t ldc, int mode, float* delta,
                      lgnn_stream_t stream) {
  return lgnn_hess_rhs_pitched_f32(logits, ld, C, idx, m, c0, ncols, ldc, (int64_t)ncols * ldc, mode, delta, stream);
}

int lgnn_hess_rhs_pitched_f32(const float* logits, int64_t ld, int32_t C, const int64_t* idx, int64_t m,
                              int32_t c0, int32_t ncols, int32_t ldc, int64_t ld_delta, int mode,
                              float* delta, lgnn_stream_t stream) {
  if (!logits || !delta || C < 1 || ld < C || m < 0 || c0 < 0 || ncols < 0 || c0 + ncols > C || ldc < C ||
      ld_delta < (int64_t)ncols * ldc)
    return fail(LGNN_E_BADARG, "hess_rhs: bad argument");
  if (mode != LGNN_HESS_REFERENCE && mode != LGNN_HESS_GGN) return fail(LGNN_E_BADARG, "hess_rhs: unknown mode %d", mode);
  if (m == 0 || ncols == 0) return LGNN_OK;
  if (!idx) return fail(LGNN_E_BADARG, "hess_rhs: null idx");
  size_t smem = (size_t)(HESS_THREADS / 32) * 2 * C * sizeof(float);
  if (smem > 200 * 1024) return fail(LGNN_E_UNSUPPORTED, "hess_rhs: C=%d too large", C);
  if (smem > 48 * 1024)
    LGNN_CUDA_TRY(cudaFuncSetAttribute(hess_rhs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (m + HESS_THREADS / 32 - 1) / (HESS_THREADS / 32);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  hess_rhs_kernel<<<(unsigned)blocks, HESS_THREADS, smem, as_stream(stream)>>>(
      logits, ld, C, idx, m, c0, ncols, ldc, ld_delta, mode, delta);
  LGNN_LAUNCH_CHECK("hess_rhs_kernel");
  return LGNN_OK;
}

int lgnn_mask_edge_values(int64_t nnz, const int32_t* col, const float* val, const uint8_t* keep,
                          float* out, lgnn_stream_t stream) {
  if (nnz < 0 || (nnz > 0 && (!col || !val || !keep || !out))) return fail(LGNN_E_BADARG, "mask_edge_values: bad argument");
  if (nnz == 0) return LGNN_OK;
  int64_t blocks = (nnz + HESS_THREADS - 1) / HESS_THREADS;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  mask_edge_values_kernel<<<(unsigned)blocks, HESS_THREADS, 0, as_stream(stream)>>>(nnz, col, val, keep, out);
  LGNN_LAUNCH_CHECK("mask_edge_values_kernel");
  return LGNN_OK;
}

int lgnn_relu_mask_mul_f32(const float* in, int64_t ldi, const float* act, int64_t lda, float* out,
                           int64_t ldo, int64_t n_rows, int32_t group, int64_t d,
                           lgnn_stream_t stream) {
  if (!in || !act || !out || n_rows < 0 || group < 1 || d < 0 || ldi < d || lda < d || ldo < d)
    return fail(LGNN_E_BADARG, "relu_mask_mul: bad argument");
  if (n_rows == 0 || d == 0) return LGNN_OK;
  int vec = (d % 4 == 0) && (ldi % 4 == 0) && (lda % 4 == 0) && (ldo % 4 == 0) &&
            ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(act) & 15) == 0) &&
            ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  int64_t total = n_rows * group * (vec ? d / 4 : d);
  int64_t blocks = (total + HESS_THREADS - 1) / HESS_THREADS;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  relu_mask_mul_kernel<<<(unsigned)blocks, HESS_THREADS, 0, as_stream(stream)>>>(
      in, ldi, act, lda, out, ldo, n_rows, group, d, vec);
  LGNN_LAUNCH_CHECK("relu_mask_mul_kernel");
  return LGNN_OK;
}

}  // extern "C"
