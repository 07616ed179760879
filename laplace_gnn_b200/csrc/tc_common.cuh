// PTX wrappers shared by the tcgen05 kernels (SYRK, fused GEMM + relu' mask): mbarriers, TMA,
// tcgen05 fences / commit / MMA, shared-memory matrix descriptors, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lgnn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
#ifdef LGNN_MBAR_DEBUG
// Debug builds (-DLGNN_MBAR_DEBUG, tools/round2_*): a wait that times out records {source line, block, thread, barrier
// offset} in a device word the host can read afterwards (lgnn_debug_mbar_timeout) and gives up instead of trapping —
// a trap poisons the context and says nothing about WHICH barrier starved.
static __device__ unsigned int g_mbar_timeout[8];
__device__ __forceinline__ void mbar_wait_dbg(uint32_t bar, uint32_t parity, int line) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 21)) {
      if (atomicCAS(&g_mbar_timeout[0], 0u, (unsigned)line) == 0u) {
        g_mbar_timeout[1] = blockIdx.x;
        g_mbar_timeout[2] = threadIdx.x;
        g_mbar_timeout[3] = bar;
        g_mbar_timeout[4] = parity;
      }
      atomicAdd(&g_mbar_timeout[5], 1u);
      return;
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
#endif
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// same box delivered to the same shared-memory offset (and signalled on the same barrier offset)
// of every CTA of the cluster whose bit is set in cta_mask
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                                      uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(cta_mask)
      : "memory");
}
// bring the box into L2 only (no shared memory, no barrier): latency hiding beyond the ring depth
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// arrive on the barrier at the same offset in every CTA of cta_mask once all prior MMAs completed
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- warp-collective issue (the whole MMA warp runs the loop, ONE elected lane issues) ----------------------
// Issuing from `if (lane == 0) { ... }` makes ptxas treat every tcgen05 operand as possibly divergent: each MMA
// becomes a waterfall loop (ELECT, R2UR.BROADCAST per operand, UTCHMMA, BRA.U.ANY) of ~60-100 issue cycles —
// more than the 32-128 cycles the tensor pipe needs for it, so the pipe idled 43-50 % of the time (ncu,
// profiles/r2b_*).  With elect.sync in the same asm block and warp-uniform operands the operands live in uniform
// registers and an MMA is one predicated UTCHMMA.
__device__ __forceinline__ void tc_mma_tf32_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void tc_mma_tf32_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
// one lane of a converged warp (elect.sync): ptxas knows the guarded region runs single-lane, so TMA / tcgen05
// operands are moved to uniform registers with plain R2UR instead of waterfall loops
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
  return pred != 0;
}
// a value every lane of the warp holds, told to the compiler (lane 0's copy broadcast): keeps address arithmetic
// derived from it in uniform registers
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor, version 1).
//   K-major, no swizzle (layout 0): 8-row x 16-byte core matrices; LBO = stride between the 16-byte
//     k-chunks, SBO = stride between 8-row groups.
//   K-major, 128-byte swizzle (layout 2): rows of 128 bytes, 8-row atoms of 1024 bytes (what TMA
//     writes with CU_TENSOR_MAP_SWIZZLE_128B); SBO = 1024, LBO unused; the k-step inside the atom is
//     taken by adding its byte offset to the start address.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo,
                                                   uint32_t layout_type = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, dense
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t m, uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  if (q != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

}  // namespace lgnn
