// Device helpers shared by the SpMM kernels (spmm.cu, spmm_packed.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lgnn {

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
  a.x = fmaf(v, x.x, a.x);
  a.y = fmaf(v, x.y, a.y);
  a.z = fmaf(v, x.z, a.z);
  a.w = fmaf(v, x.w, a.w);
}

constexpr int BULK_CONSUMERS = 256;
constexpr int BULK_THREADS = BULK_CONSUMERS + 32;
constexpr int BULK_MAX_STAGES = 32;
constexpr int BULK_NNZ_PER_CTA = 4096;
constexpr int BULK_SMEM_RING = 192 * 1024;

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bulk_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bulk_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 27)) __trap();  // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// first row r in [0, n_rows] with rowptr[r] >= target
__device__ __forceinline__ int64_t row_lower_bound(const int64_t* __restrict__ rowptr, int64_t n_rows,
                                                   int64_t target) {
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// unit-compacted slabs for column groups with g % 4 == 2 (spmm_units_even.cu); the C ABI entry points of
// spmm_units.cu validate the arguments and dispatch here
// src [n_rows, lds] dense [g][h] rows (nullptr: headers only) -> dst: row n at dst + n*ldd (in place when
// dst == src, ldd == lds), or, with row_first, at the absolute slot row_first[n] (ragged rows, back to back)
struct UnitPackArgs {
  const float* src;
  int64_t lds;
  float* dst;
  int64_t ldd;
  const int64_t* row_first;
  const float* act;
  int64_t lda;
  int h;
  uint2* hdr;     // nullptr: values only
};
int unit_pack_even(const UnitPackArgs& A, int64_t n_rows, int g, cudaStream_t st);
int spmm_units_even(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val, const float* slab,
                    int64_t lds, const uint2* hdr, int g, int nblk, float* y, int64_t ldy, int variant,
                    cudaStream_t st);

}  // namespace lgnn
