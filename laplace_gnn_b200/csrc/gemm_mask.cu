// delta_{l-1} = (gZ_l W_l) ⊙ 1[H_{l-1} > 0] — the step between two multi-RHS SpMMs of the KFAC
// backward (autograd of gnn/models/base_gnn.py:150 + layers.py:45 inside curvlinops/kfac.py:653-661),
// as ONE kernel on the 5th-generation tensor cores: 3xTF32 (hi·hi + hi·lo + lo·hi) for fp32-faithful
// products, the relu' mask fused into the epilogue.
//
//   out[r, 0:N] = (A[r, 0:K] · W[0:K, 0:N]) ⊙ (act[r / group, 0:N] > 0)
//   A = gZ viewed as [n_nodes * group, K] row-major (K-major operand as it lies in HBM).
//
// The weights do not fit one SM as resident hi/lo operands (N x K x 8 B = 512 KB at 256 x 256), and
// streaming them from L2 per tile would cost ~12 TB/s of L2 traffic.  So a CLUSTER of N/64 CTAs
// splits the N columns: every CTA keeps its 64-column slice of W^T (hi and lo, 128 KB) resident in
// shared memory for its whole life, and the cluster shares ONE read of each A tile: every CTA
// TMA-loads 128/cluster rows of the tile and multicasts them into all CTAs' shared memory.
//
// Shared memory operands are K-major with the 128-byte swizzle exactly as TMA writes them, so the
// raw fp32 tile IS the "hi" operand (kind::tf32 ignores the low 13 mantissa bits); only "lo" =
// x - trunc_tf32(x) is materialised by the transform warps, chunk for chunk at the same swizzled
// address — no transposition, no bank conflicts.
//
// Per CTA (10 warps): warp 0 TMA producer; warp 1 MMA issuer (one thread; 12 MMAs M=128 N=64 K=8 per
// 32-wide k-block; accumulator double-buffered in TMEM); warps 2-5 lo transform; warps 6-9 epilogue
// (tcgen05.ld -> mask -> 128-bit stores).  A ring stage is recycled when the MMAs of ALL CTAs of the
// cluster have consumed it (tcgen05.commit multicast onto every CTA's "empty" barrier), because the
// next multicast overwrites it everywhere.
//
// Roofline: useful flops 2·M·K·N; HBM bytes M·(K + N)·4 + mask; at K = N = 256 both
// bounds are ~40 ms per 115 M rows — the kernel is balanced between the tensor pipe and HBM.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace lgnn {

constexpr int GM_BM = 128;            // rows per tile
constexpr int GM_BN = 64;             // output columns per CTA
constexpr int GM_BK = 32;             // floats per k-block = 128 bytes = one swizzle row
constexpr int GM_STAGES = 3;
constexpr int GM_THREADS = 320;
constexpr int GM_A_TILE = GM_BM * GM_BK * 4;   // 16 KB
constexpr int GM_B_TILE = GM_BN * GM_BK * 4;   // 8 KB
constexpr int GM_TILES_PER_CLUSTER = 32;
constexpr int GM_MAX_KB = 8;                   // K <= 256

struct GmParams {
  int64_t m_rows;
  int64_t tiles_total;
  int n_kb;          // k-blocks (K_pad / 32)
  int last_ksteps;   // 8-wide k-steps in the last k-block (1..4)
  int cl;            // cluster size = N / 64
  int group;
  const float* act;  // may be null (no mask)
  int64_t ld_act;
  float* out;
  int64_t ldo;
};

__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_mask_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bhi,
                 const __grid_constant__ CUtensorMap tm_blo, const GmParams P) {
  extern __shared__ uint8_t gm_smem_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gm_smem_) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = b_hi + (size_t)P.n_kb * GM_B_TILE;
  uint8_t* a_base = b_lo + (size_t)P.n_kb * GM_B_TILE;   // stage s: raw at +s*32 KB, lo 16 KB behind it
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + (size_t)GM_STAGES * 2 * GM_A_TILE);
  uint64_t* full_raw = bars;                    // [STAGES] TMA bytes of all cluster slices landed
  uint64_t* lo_ready = full_raw + GM_STAGES;    // [STAGES] transform warps wrote lo
  uint64_t* empty = lo_ready + GM_STAGES;       // [STAGES] MMAs of every CTA of the cluster consumed it
  uint64_t* b_full = empty + GM_STAGES;         // [1]
  uint64_t* acc_full = b_full + 1;              // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = P.cl;
  const uint32_t rank = cl > 1 ? cluster_ctarank() : 0u;
  const int64_t cluster_id = blockIdx.x / cl;
  const int64_t tile_beg = cluster_id * GM_TILES_PER_CLUSTER;
  int64_t tile_end = tile_beg + GM_TILES_PER_CLUSTER;
  if (tile_end > P.tiles_total) tile_end = P.tiles_total;
  const int64_t n_tiles = tile_end - tile_beg;              // >= 1 by construction of the grid
  const int64_t total_it = n_tiles * P.n_kb;
  const uint16_t cta_mask = (uint16_t)((1u << cl) - 1u);
  const int n0 = (int)rank * GM_BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GM_STAGES; ++i) {
      mbar_init(smem_u32(&full_raw[i]), 1);
      mbar_init(smem_u32(&lo_ready[i]), 4);
      mbar_init(smem_u32(&empty[i]), (uint32_t)cl);
    }
    mbar_init(smem_u32(b_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // every CTA's barriers exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
      // resident weights: this CTA's 64 rows of W^T, hi and lo, all k-blocks
      mbar_arrive_expect_tx(smem_u32(b_full), (uint32_t)(2 * P.n_kb * GM_B_TILE));
      for (int kb = 0; kb < P.n_kb; ++kb) {
        tma_load_2d(smem_u32(b_hi + (size_t)kb * GM_B_TILE), &tm_bhi, kb * GM_BK, n0, smem_u32(b_full));
        tma_load_2d(smem_u32(b_lo + (size_t)kb * GM_B_TILE), &tm_blo, kb * GM_BK, n0, smem_u32(b_full));
      }
      const int rows_per_cta = GM_BM / cl;
      for (int64_t it = 0; it < total_it; ++it) {
        const int s = (int)(it % GM_STAGES);
        const uint32_t ph = (uint32_t)((it / GM_STAGES) & 1);
        const int64_t tile = tile_beg + it / P.n_kb;
        const int kb = (int)(it % P.n_kb);
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        const uint32_t bar = smem_u32(&full_raw[s]);
        mbar_arrive_expect_tx(bar, (uint32_t)GM_A_TILE);
        const uint32_t dst = smem_u32(a_base + (size_t)s * 2 * GM_A_TILE + (size_t)rank * rows_per_cta * GM_BK * 4);
        const int row0 = (int)(tile * GM_BM + (int64_t)rank * rows_per_cta);
        if (cl > 1) tma_load_2d_multicast(dst, &tm_a, kb * GM_BK, row0, bar, cta_mask);
        else tma_load_2d(dst, &tm_a, kb * GM_BK, row0, bar);
      }
      // tail: the last uses of every stage have been released by ALL CTAs, i.e. no remote arrive is
      // still heading for this CTA's barriers when it leaves
      for (int64_t it = total_it; it < total_it + GM_STAGES; ++it) {
        if (it < GM_STAGES) continue;   // stage never used
        mbar_wait(smem_u32(&empty[it % GM_STAGES]), (uint32_t)((it / GM_STAGES) & 1) ^ 1u);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(GM_BM, GM_BN);
      mbar_wait(smem_u32(b_full), 0);
      for (int64_t t = 0; t < n_tiles; ++t) {
        const int as = (int)(t & 1);
        const uint32_t aph = (uint32_t)((t >> 1) & 1);
        mbar_wait(smem_u32(&acc_empty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(as * GM_BN);
        for (int kb = 0; kb < P.n_kb; ++kb) {
          const int64_t it = t * P.n_kb + kb;
          const int s = (int)(it % GM_STAGES);
          const uint32_t ph = (uint32_t)((it / GM_STAGES) & 1);
          mbar_wait(smem_u32(&full_raw[s]), ph);
          mbar_wait(smem_u32(&lo_ready[s]), ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(a_base + (size_t)s * 2 * GM_A_TILE);
          const uint32_t a_lo = a_hi + GM_A_TILE;
          const uint32_t bh = smem_u32(b_hi + (size_t)kb * GM_B_TILE);
          const uint32_t bl = smem_u32(b_lo + (size_t)kb * GM_B_TILE);
          const int ksteps = (kb == P.n_kb - 1) ? P.last_ksteps : GM_BK / 8;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t off = (uint32_t)ks * 32u;   // 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint64_t da_hi = make_smem_desc(a_hi + off, 0, 1024, 2);
            const uint64_t da_lo = make_smem_desc(a_lo + off, 0, 1024, 2);
            const uint64_t db_hi = make_smem_desc(bh + off, 0, 1024, 2);
            const uint64_t db_lo = make_smem_desc(bl + off, 0, 1024, 2);
            tc_mma_tf32(d, da_hi, db_hi, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
            tc_mma_tf32(d, da_hi, db_lo, idesc, 1u);
            tc_mma_tf32(d, da_lo, db_hi, idesc, 1u);
          }
          if (cl > 1) tc_commit_multicast(smem_u32(&empty[s]), cta_mask);
          else tc_commit(smem_u32(&empty[s]));
        }
        tc_commit(smem_u32(&acc_full[as]));
      }
    }
  } else if (warp < 6) {
    // ===================================================================== lo transform warps
    const int t = threadIdx.x - 64;  // 0..127
    for (int64_t it = 0; it < total_it; ++it) {
      const int s = (int)(it % GM_STAGES);
      const uint32_t ph = (uint32_t)((it / GM_STAGES) & 1);
      mbar_wait(smem_u32(&full_raw[s]), ph);
      const float4* raw = reinterpret_cast<const float4*>(a_base + (size_t)s * 2 * GM_A_TILE);
      float4* lo = reinterpret_cast<float4*>(a_base + (size_t)s * 2 * GM_A_TILE + GM_A_TILE);
#pragma unroll
      for (int c = 0; c < GM_A_TILE / 16 / 128; ++c) {
        const float4 x = raw[t + 128 * c];
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        lo[t + 128 * c] = l;
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&lo_ready[s]));
    }
  } else {
    // ===================================================================== epilogue warps
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int as = (int)(t & 1);
      const uint32_t aph = (uint32_t)((t >> 1) & 1);
      mbar_wait(smem_u32(&acc_full[as]), aph);
      tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * GM_BN);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[32 * h + 0]), "=r"(v[32 * h + 1]), "=r"(v[32 * h + 2]), "=r"(v[32 * h + 3]),
              "=r"(v[32 * h + 4]), "=r"(v[32 * h + 5]), "=r"(v[32 * h + 6]), "=r"(v[32 * h + 7]),
              "=r"(v[32 * h + 8]), "=r"(v[32 * h + 9]), "=r"(v[32 * h + 10]), "=r"(v[32 * h + 11]),
              "=r"(v[32 * h + 12]), "=r"(v[32 * h + 13]), "=r"(v[32 * h + 14]), "=r"(v[32 * h + 15]),
              "=r"(v[32 * h + 16]), "=r"(v[32 * h + 17]), "=r"(v[32 * h + 18]), "=r"(v[32 * h + 19]),
              "=r"(v[32 * h + 20]), "=r"(v[32 * h + 21]), "=r"(v[32 * h + 22]), "=r"(v[32 * h + 23]),
              "=r"(v[32 * h + 24]), "=r"(v[32 * h + 25]), "=r"(v[32 * h + 26]), "=r"(v[32 * h + 27]),
              "=r"(v[32 * h + 28]), "=r"(v[32 * h + 29]), "=r"(v[32 * h + 30]), "=r"(v[32 * h + 31])
            : "r"(taddr + (uint32_t)(32 * h)));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q)   // pin the uses of v[] behind the wait (volatile asms keep their order)
        asm volatile("" : "+r"(v[16 * q + 0]), "+r"(v[16 * q + 1]), "+r"(v[16 * q + 2]), "+r"(v[16 * q + 3]),
                          "+r"(v[16 * q + 4]), "+r"(v[16 * q + 5]), "+r"(v[16 * q + 6]), "+r"(v[16 * q + 7]),
                          "+r"(v[16 * q + 8]), "+r"(v[16 * q + 9]), "+r"(v[16 * q + 10]), "+r"(v[16 * q + 11]),
                          "+r"(v[16 * q + 12]), "+r"(v[16 * q + 13]), "+r"(v[16 * q + 14]), "+r"(v[16 * q + 15]));
      // the accumulator is in registers: hand the TMEM buffer back before touching global memory
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
      const int64_t row = (tile_beg + t) * GM_BM + quarter * 32 + lane;
      if (row < P.m_rows) {
        float4* dst = reinterpret_cast<float4*>(P.out + row * P.ldo + n0);
        if (P.act != nullptr) {
          const float4* m = reinterpret_cast<const float4*>(P.act + (row / P.group) * P.ld_act + n0);
#pragma unroll
          for (int j = 0; j < GM_BN / 4; ++j) {
            const float4 a = __ldg(m + j);
            float4 o;
            o.x = a.x > 0.f ? __uint_as_float(v[4 * j + 0]) : 0.f;
            o.y = a.y > 0.f ? __uint_as_float(v[4 * j + 1]) : 0.f;
            o.z = a.z > 0.f ? __uint_as_float(v[4 * j + 2]) : 0.f;
            o.w = a.w > 0.f ? __uint_as_float(v[4 * j + 3]) : 0.f;
            dst[j] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < GM_BN / 4; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// wt_hi[n, k] = W[k, n] (the tensor core truncates it to tf32 itself), wt_lo[n, k] = W[k, n] - trunc_tf32(W[k, n]);
// columns k in [K, k_pad) are zero.
__global__ void gemm_mask_prepare_kernel(const float* __restrict__ w, int64_t ldw, int k, int n, int k_pad,
                                         float* __restrict__ wt_hi, float* __restrict__ wt_lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * k_pad) return;
  const int nn = i / k_pad, kk = i - nn * k_pad;
  float x = 0.f;
  if (kk < k) x = w[(int64_t)kk * ldw + nn];
  wt_hi[i] = x;
  wt_lo[i] = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

static int encode_2d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(LGNN_E_CUDA, "gemm_mask: cuTensorMapEncodeTiled not available");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LGNN_E_CUDA, "gemm_mask: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return LGNN_OK;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" {

int lgnn_gemm_mask_supported(int64_t k, int64_t n) {
  return (k >= 1 && k <= GM_MAX_KB * GM_BK && (n == 64 || n == 128 || n == 256)) ? 1 : 0;
}

int64_t lgnn_gemm_mask_kpad(int64_t k) { return (k + GM_BK - 1) / GM_BK * GM_BK; }

int lgnn_gemm_mask_prepare_f32(const float* w, int64_t ldw, int64_t k, int64_t n, float* wt_hi, float* wt_lo,
                               lgnn_stream_t stream) {
  if (!w || !wt_hi || !wt_lo || ldw < n) return fail(LGNN_E_BADARG, "gemm_mask_prepare: bad argument");
  if (!lgnn_gemm_mask_supported(k, n)) return fail(LGNN_E_UNSUPPORTED, "gemm_mask: needs K <= 256 and N in {64, 128, 256}");
  const int k_pad = (int)lgnn_gemm_mask_kpad(k);
  const int total = (int)n * k_pad;
  gemm_mask_prepare_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(w, ldw, (int)k, (int)n, k_pad, wt_hi, wt_lo);
  LGNN_LAUNCH_CHECK("gemm_mask_prepare_kernel");
  return LGNN_OK;
}

int lgnn_gemm_mask_f32(const float* a, int64_t lda, int64_t m_rows, int64_t k, const float* wt_hi,
                       const float* wt_lo, int64_t n, const float* act, int64_t ld_act, int32_t group,
                       float* out, int64_t ldo, lgnn_stream_t stream) {
  if (m_rows < 0 || !wt_hi || !wt_lo || !out || group < 1 || lda < k || ldo < n || (act && ld_act < n))
    return fail(LGNN_E_BADARG, "gemm_mask: bad argument");
  if (!lgnn_gemm_mask_supported(k, n)) return fail(LGNN_E_UNSUPPORTED, "gemm_mask: needs K <= 256 and N in {64, 128, 256}");
  if (m_rows == 0) return LGNN_OK;
  if (!a) return fail(LGNN_E_BADARG, "gemm_mask: null a");
  if (m_rows >= ((int64_t)1 << 31) - GM_BM) return fail(LGNN_E_UNSUPPORTED, "gemm_mask: more than 2^31 rows");
  if ((lda % 4) || (ldo % 4) || (act && (ld_act % 4)) || (reinterpret_cast<uintptr_t>(a) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (act && (reinterpret_cast<uintptr_t>(act) & 15)) ||
      (reinterpret_cast<uintptr_t>(wt_hi) & 15) || (reinterpret_cast<uintptr_t>(wt_lo) & 15))
    return fail(LGNN_E_ALIGN, "gemm_mask: operands must be 16-byte aligned with pitches that are multiples of 4 floats");
  cudaStream_t st = as_stream(stream);
  const int k_pad = (int)lgnn_gemm_mask_kpad(k);
  GmParams P;
  P.m_rows = m_rows;
  P.tiles_total = (m_rows + GM_BM - 1) / GM_BM;
  P.n_kb = k_pad / GM_BK;
  P.last_ksteps = (int)((k - (int64_t)(P.n_kb - 1) * GM_BK + 7) / 8);
  P.cl = (int)(n / GM_BN);
  P.group = group;
  P.act = act;
  P.ld_act = ld_act;
  P.out = out;
  P.ldo = ldo;
  CUtensorMap tm_a, tm_bhi, tm_blo;
  int rc;
  if ((rc = encode_2d(&tm_a, a, (uint64_t)k, (uint64_t)m_rows, (uint64_t)lda * 4, GM_BK, GM_BM / P.cl))) return rc;
  if ((rc = encode_2d(&tm_bhi, wt_hi, (uint64_t)k_pad, (uint64_t)n, (uint64_t)k_pad * 4, GM_BK, GM_BN))) return rc;
  if ((rc = encode_2d(&tm_blo, wt_lo, (uint64_t)k_pad, (uint64_t)n, (uint64_t)k_pad * 4, GM_BK, GM_BN))) return rc;
  const size_t smem = 1024 + (size_t)2 * P.n_kb * GM_B_TILE + (size_t)GM_STAGES * 2 * GM_A_TILE + 256;
  LGNN_CUDA_TRY(cudaFuncSetAttribute(gemm_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_clusters = (P.tiles_total + GM_TILES_PER_CLUSTER - 1) / GM_TILES_PER_CLUSTER;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * P.cl), 1, 1);
  cfg.blockDim = dim3(GM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LGNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_mask_kernel, tm_a, tm_bhi, tm_blo, P));
  LGNN_LAUNCH_CHECK("gemm_mask_kernel");
  return LGNN_OK;
}

}  // extern "C"
