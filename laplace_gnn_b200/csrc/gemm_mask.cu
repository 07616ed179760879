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
// The A operand lives in TENSOR MEMORY: TMA writes the raw fp32 tile into a 5-deep shared-memory
// ring (128-byte swizzle), four transform warps — one thread per tile row, which is also one TMEM
// lane — read their row, split it into hi = x (kind::tf32 ignores the low 13 mantissa bits) and
// lo = x - trunc_tf32(x), and tcgen05.st both into a 4-deep ring of TMEM operand stages.  The MMAs
// then take A from TMEM and only the resident weight slice from shared memory (K-major, 128-byte
// swizzle, as TMA wrote it).  With M=128 N=64 an all-shared-memory operand path needs 192 B/clk
// of shared-memory bandwidth (more than the 128 B/clk an SM has) — measured 9.5 us per tile in v1;
// with A in TMEM the tensor path reads 64 B/clk.
//
// Per CTA (14 warps): warp 0 TMA producer; warp 1 MMA issuer (one thread; 12 MMAs M=128 N=64 K=8 per
// 32-wide k-block; accumulator double-buffered in TMEM); warps 2-9 transform (every k-block by all eight, in
// order: warps 2-5 its columns 0..15, warps 6-9 its columns 16..31); warps 10-13 epilogue (tcgen05.ld ->
// shared-memory staging -> 128-bit stores of 64 contiguous bytes per row; the relu' mask words of the
// NEXT tile are prefetched while the current one is stored).  A raw ring
// stage is recycled when the transform warps of ALL CTAs of the cluster have read it (remote
// mbarrier arrives), because the next multicast overwrites it everywhere.
//
// Roofline: useful flops 2·M·K·N; HBM bytes M·(K + N)·4 + mask; at K = N = 256 both
// bounds are ~40 ms per 115 M rows.  Measured (profiles/r1c_gemm_lab.txt): 97 useful TFLOP/s at
// K = N = 256 (291 issued, 25 % of the TF32 peak) — with N = 64 per CTA every MMA still reads a
// full 4 KB A operand, so the operand path, not the MMA floor, paces the k-loop (~90 cycles per
// MMA instead of 32).  An L2 prefetch ahead of the ring and doubling the transform warps changed
// nothing, which rules out HBM latency and the transform chain.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace lgnn {

constexpr int GM_BM = 128;            // rows per tile
constexpr int GM_BN = 64;             // output columns per CTA
constexpr int GM_BK = 32;             // floats per k-block = 128 bytes = one swizzle row
constexpr int GM_RAW_STAGES = 5;      // shared-memory ring of raw A k-blocks
constexpr int GM_TM_STAGES = 4;       // TMEM ring of (hi, lo) A operand k-blocks, 64 columns each
constexpr int GM_THREADS = 448;            // 1 TMA + 1 MMA + 8 transform + 4 epilogue warps
constexpr int GM_A_TILE = GM_BM * GM_BK * 4;   // 16 KB
constexpr int GM_B_TILE = GM_BN * GM_BK * 4;   // 8 KB
constexpr int GM_MAX_KB = 8;                   // K <= 256
constexpr int GM_STG_PITCH = 80;               // epilogue staging: 16 floats per row + 16 bytes pad
constexpr int GM_STG_WARP = 32 * GM_STG_PITCH; // per epilogue warp
constexpr uint32_t GM_TMEM_COLS = 512;         // D: 2 x 64, A: 4 x 64
constexpr uint32_t GM_TMEM_A0 = 128;

// Ablation switches for the lab copy of this kernel (gemm_mask_lab.cu compiles this file a second time with
// LGNN_GM_ABLATE defined and every global symbol renamed): which stage paces a tile?  In the product build GM_ABL
// is the constant false and every guarded statement is what it was.
//   1: the MMA thread issues no MMA (commits only)      2: only hi.hi of the three products is issued
//   4: transform warps skip the tcgen05.st of hi / lo    8: epilogue skips the global stores
#ifdef LGNN_GM_ABLATE
#define GM_ABL(bit) ((P.ablate & (bit)) != 0)
#else
#define GM_ABL(bit) false
#endif

struct GmParams {
#ifdef LGNN_GM_ABLATE
  int ablate;
#endif
  int64_t m_rows;
  int64_t tiles_total;
  int n_kb;          // k-blocks (K_pad / 32)
  int last_ksteps;   // 8-wide k-steps in the last k-block (1..4)
  int cl;            // cluster size = N / 64
  int tiles_per_cluster;
  int group;
  const float* act;  // may be null (no mask)
  int64_t ld_act;
  const float* bias; // may be null; added after the mask (the forward linear layer passes a bias and no mask)
  float* out;
  int64_t ldo;
};

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
  // default semantics (release at CTA scope), like cutlass::arch::ClusterBarrier::arrive(cta_id): a
  // cluster-scope release costs a MEMBAR.ALL.GPU + ERRBAR per call (26 % of all stall samples in the
  // first ncu capture); the data this arrive orders (shared-memory reads) is already in registers.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// MASKED: the relu' mask is applied in the epilogue (dense slabs); false: plain product (+ bias) — the unit-compacted
// hot path and the forward linear layer, whose instantiation carries neither the 64 mask-prefetch registers nor the
// spills they caused
template <bool MASKED>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_mask_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bhi,
                 const __grid_constant__ CUtensorMap tm_blo, const GmParams P) {
  extern __shared__ uint8_t gm_smem_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gm_smem_) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = b_hi + (size_t)P.n_kb * GM_B_TILE;
  uint8_t* a_raw = b_lo + (size_t)P.n_kb * GM_B_TILE;                    // [RAW_STAGES][16 KB]
  uint8_t* stg = a_raw + (size_t)GM_RAW_STAGES * GM_A_TILE;               // [4 warps][32 rows][80 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 4 * GM_STG_WARP);
  uint64_t* full_raw = bars;                       // [RAW] TMA bytes of all cluster slices landed
  uint64_t* empty_raw = full_raw + GM_RAW_STAGES;  // [RAW] transform warps of every CTA have read it
  uint64_t* ta_full = empty_raw + GM_RAW_STAGES;   // [TM]  hi / lo operand stage written to TMEM
  uint64_t* ta_empty = ta_full + GM_TM_STAGES;     // [TM]  MMAs that read it completed
  uint64_t* b_full = ta_empty + GM_TM_STAGES;      // [1]
  uint64_t* acc_full = b_full + 1;                 // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = (int)warp_uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int cl = P.cl;
  const uint32_t rank = cl > 1 ? cluster_ctarank() : 0u;
  const int64_t cluster_id = blockIdx.x / cl;
  const int64_t tile_beg = cluster_id * P.tiles_per_cluster;
  int64_t tile_end = tile_beg + P.tiles_per_cluster;
  if (tile_end > P.tiles_total) tile_end = P.tiles_total;
  // 32-bit loop state everywhere below, ring stages and phases stepped incrementally: a 64-bit division per
  // k-block (it / n_kb, it % 5) was ~500 issue cycles of the single producer thread — the whole 625-cycle k-block
  // budget of the round-1 kernel (profiles/r2c_*)
  const int n_tiles = (int)(tile_end - tile_beg);           // >= 1 by construction of the grid
  const int total_it = n_tiles * P.n_kb;
  const uint16_t cta_mask = (uint16_t)((1u << cl) - 1u);
  const int n0 = (int)rank * GM_BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GM_RAW_STAGES; ++i) {
      mbar_init(smem_u32(&full_raw[i]), 1);
      mbar_init(smem_u32(&empty_raw[i]), (uint32_t)(8 * cl));
    }
    for (int i = 0; i < GM_TM_STAGES; ++i) {
      mbar_init(smem_u32(&ta_full[i]), 8);
      mbar_init(smem_u32(&ta_empty[i]), 1);
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
                 "r"(GM_TMEM_COLS)
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
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
      // resident weights: this CTA's 64 rows of W^T, hi and lo, all k-blocks
      mbar_arrive_expect_tx(smem_u32(b_full), (uint32_t)(2 * P.n_kb * GM_B_TILE));
      for (int kb = 0; kb < P.n_kb; ++kb) {
        tma_load_2d(smem_u32(b_hi + (size_t)kb * GM_B_TILE), &tm_bhi, kb * GM_BK, n0, smem_u32(b_full));
        tma_load_2d(smem_u32(b_lo + (size_t)kb * GM_B_TILE), &tm_blo, kb * GM_BK, n0, smem_u32(b_full));
      }
      const int rows_per_cta = GM_BM / cl;
      const uint32_t raw0 = smem_u32(a_raw) + (uint32_t)rank * (uint32_t)(rows_per_cta * GM_BK * 4);
      int row0 = (int)(tile_beg * GM_BM) + (int)rank * rows_per_cta;
      int s = 0, kb = 0;
      uint32_t ph = 0;
      for (int it = 0; it < total_it; ++it) {
        mbar_wait(smem_u32(&empty_raw[s]), ph ^ 1u);
        const uint32_t bar = smem_u32(&full_raw[s]);
        mbar_arrive_expect_tx(bar, (uint32_t)GM_A_TILE);
        const uint32_t dst = raw0 + (uint32_t)s * GM_A_TILE;
        if (cl > 1) tma_load_2d_multicast(dst, &tm_a, kb * GM_BK, row0, bar, cta_mask);
        else tma_load_2d(dst, &tm_a, kb * GM_BK, row0, bar);
        if (++kb == P.n_kb) { kb = 0; row0 += GM_BM; }
        if (++s == GM_RAW_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // the WHOLE warp walks the loop (uniform control flow, uniform operands), one elected lane issues each MMA /
    // commit (tc_common.cuh: tc_mma_tf32_ts_e)
    {
      const uint32_t tb = warp_uniform(tmem_base);
      const uint32_t idesc = make_idesc_tf32(GM_BM, GM_BN);
      const uint32_t bh0 = smem_u32(b_hi), bl0 = smem_u32(b_lo);
      mbar_wait(smem_u32(b_full), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int as = t & 1;
        const uint32_t aph = (uint32_t)((t >> 1) & 1);
        mbar_wait(smem_u32(&acc_empty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tb + (uint32_t)(as * GM_BN);
        for (int kb = 0; kb < P.n_kb; ++kb) {
          mbar_wait(smem_u32(&ta_full[s]), ph);
          tc_fence_after();
          const uint32_t a_hi = tb + GM_TMEM_A0 + (uint32_t)(s * 64);   // 32 columns hi, 32 columns lo
          const uint32_t a_lo = a_hi + 32;
          const uint32_t bh = bh0 + (uint32_t)kb * GM_B_TILE;
          const uint32_t bl = bl0 + (uint32_t)kb * GM_B_TILE;
          const int ksteps = (kb == P.n_kb - 1) ? P.last_ksteps : GM_BK / 8;
#pragma unroll 4
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t off = (uint32_t)ks * 32u;   // 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint64_t db_hi = make_smem_desc(bh + off, 0, 1024, 2);
            const uint64_t db_lo = make_smem_desc(bl + off, 0, 1024, 2);
            const uint32_t ca = (uint32_t)(ks * 8);    // 8 tf32 = 8 TMEM columns
            if (!GM_ABL(1)) tc_mma_tf32_ts_e(d, a_hi + ca, db_hi, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
            if (!GM_ABL(1) && !GM_ABL(2)) {
              tc_mma_tf32_ts_e(d, a_hi + ca, db_lo, idesc, 1u);
              tc_mma_tf32_ts_e(d, a_lo + ca, db_hi, idesc, 1u);
            }
          }
          tc_commit_e(smem_u32(&ta_empty[s]));
          if (++s == GM_TM_STAGES) { s = 0; ph ^= 1u; }
        }
        tc_commit_e(smem_u32(&acc_full[as]));
      }
    }
  } else if (warp < 10) {
    // ===================================================================== transform warps
    // thread = tile row = TMEM lane (a warp may only touch its own quarter of the lanes).  EVERY k-block is
    // handled by all eight warps, in order: warps 2-5 take its columns 0..15, warps 6-9 its columns 16..31
    // (4 of the row's 8 swizzled 16-byte chunks each).  (Round 1 let two sets of four warps take alternate
    // k-blocks.  With a ring of 5 stages a set then revisits a stage's barrier every 10 k-blocks — two phases
    // apart, which a parity wait cannot tell from "already there": once k-blocks landed out of order the set read
    // a stage early, arrived on its `empty` barrier a second time, the producer re-armed a `full` barrier inside
    // an open phase and the kernel died with an illegal-instruction exception.  Seen at N = 64 after the MMA
    // issue got fast, cuda-gdb, profiles/r2f_*.)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;             // 0: columns 0..15 of the k-block, 1: columns 16..31
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    int rs = 0, ts = 0;
    uint32_t rph = 0, tph = 0;
    for (int it = 0; it < total_it; ++it) {
      mbar_wait(smem_u32(&full_raw[rs]), rph);
      const uint8_t* row = a_raw + (size_t)rs * GM_A_TILE + (size_t)r * 128;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {   // 16-byte chunk c of row r sits at chunk position c ^ (r & 7)
        const float4 x = *reinterpret_cast<const float4*>(row + (((4 * half + c) ^ (r & 7)) << 4));
        hi[4 * c + 0] = __float_as_uint(x.x);
        hi[4 * c + 1] = __float_as_uint(x.y);
        hi[4 * c + 2] = __float_as_uint(x.z);
        hi[4 * c + 3] = __float_as_uint(x.w);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        lo[j] = __float_as_uint(__uint_as_float(hi[j]) - __uint_as_float(hi[j] & 0xffffe000u));
      // the raw stage is in registers: hand it back to every producer of the cluster
      __syncwarp();
      if (lane < cl) {
        if (cl > 1) mbar_arrive_remote(smem_u32(&empty_raw[rs]), (uint32_t)lane);
        else mbar_arrive(smem_u32(&empty_raw[rs]));
      }
      mbar_wait(smem_u32(&ta_empty[ts]), tph ^ 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_addr + GM_TMEM_A0 + (uint32_t)(ts * 64 + 16 * half);
      if (!GM_ABL(4)) {
        tmem_st_x16(taddr, hi);
        tmem_st_x16(taddr + 32, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ta_full[ts]));
      if (++rs == GM_RAW_STAGES) { rs = 0; rph ^= 1u; }
      if (++ts == GM_TM_STAGES) { ts = 0; tph ^= 1u; }
    }
  } else {
    // ===================================================================== epilogue warps
    // (A variant in which every thread stores its own row straight from the tcgen05.ld registers — no staging —
    // was 7 % slower at K = 256 and 15 % faster at K = 47, profiles/r2h_gemm_epilogue_ab.txt; not kept.)
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    uint8_t* my_stg = stg + (size_t)(warp - 10) * GM_STG_WARP;
    const int sub_r = lane >> 2, sub_c = lane & 3;
    constexpr bool masked = MASKED;
    // relu' mask words of one tile for this thread: mk[4*p + i] covers pass p (columns 16p + 4*sub_c ..),
    // row i*8 + sub_r.  They do not depend on the MMAs, so tile t+1's are fetched while tile t is stored.
    float4 mk[16];
    auto fetch_mask = [&](int t) {
      const int64_t row_base = (tile_beg + t) * GM_BM + quarter * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = row_base + i * 8 + sub_r;
        const bool ok = t < n_tiles && row < P.m_rows;
        const float* src = P.act + (int64_t)((uint32_t)row / (uint32_t)P.group) * P.ld_act + n0 + 4 * sub_c;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          mk[4 * p + i] = ok ? __ldg(reinterpret_cast<const float4*>(src + 16 * p)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (masked) fetch_mask(0);
    // bias of this thread's output columns (pass p: columns 16p + 4 sub_c ..): re-read per pass from L1 rather than
    // held in 16 registers — the epilogue is at the register limit and spills are paid once per tile
    const float* bias_src = P.bias ? P.bias + n0 + 4 * sub_c : nullptr;
    for (int t = 0; t < n_tiles; ++t) {
      const int as = t & 1;
      const uint32_t aph = (uint32_t)((t >> 1) & 1);
      const int64_t row_base = (tile_beg + t) * GM_BM + quarter * 32;
      unsigned long long keep = ~0ull;   // bit 4*(4*p + i) + e
      if (masked) {
        keep = 0ull;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const unsigned long long b = (mk[q].x > 0.f ? 1ull : 0ull) | (mk[q].y > 0.f ? 2ull : 0ull) |
                                       (mk[q].z > 0.f ? 4ull : 0ull) | (mk[q].w > 0.f ? 8ull : 0ull);
          keep |= b << (4 * q);
        }
        fetch_mask(t + 1);               // in flight during this tile's TMEM drain and stores
      }
      mbar_wait(smem_u32(&acc_full[as]), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * GM_BN);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + (uint32_t)(32 * h)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 2; ++q)   // pin the uses of v[] behind the wait (volatile asms keep their order)
          asm volatile("" : "+r"(v[16 * q + 0]), "+r"(v[16 * q + 1]), "+r"(v[16 * q + 2]), "+r"(v[16 * q + 3]),
                            "+r"(v[16 * q + 4]), "+r"(v[16 * q + 5]), "+r"(v[16 * q + 6]), "+r"(v[16 * q + 7]),
                            "+r"(v[16 * q + 8]), "+r"(v[16 * q + 9]), "+r"(v[16 * q + 10]), "+r"(v[16 * q + 11]),
                            "+r"(v[16 * q + 12]), "+r"(v[16 * q + 13]), "+r"(v[16 * q + 14]), "+r"(v[16 * q + 15]));
        if (h == 1) {   // the whole accumulator has left TMEM: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
        }
        // thread = row in TMEM; global memory wants lanes along a row.  Passes of 16 columns through a
        // padded staging tile: 128-bit stores by row (conflict-free at an 80-byte pitch), then every
        // instruction writes 8 rows x 64 contiguous bytes.
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          const int p = 2 * h + pp;
          const float4 bs = bias_src ? __ldg(reinterpret_cast<const float4*>(bias_src + 16 * p)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4* mine = reinterpret_cast<float4*>(my_stg + (size_t)lane * GM_STG_PITCH);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mine[j] = make_float4(__uint_as_float(v[16 * pp + 4 * j + 0]), __uint_as_float(v[16 * pp + 4 * j + 1]),
                                  __uint_as_float(v[16 * pp + 4 * j + 2]), __uint_as_float(v[16 * pp + 4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + sub_r;
            const int64_t row = row_base + rl;
            float4 o = *reinterpret_cast<const float4*>(my_stg + (size_t)rl * GM_STG_PITCH + sub_c * 16);
            const unsigned kb4 = (unsigned)(keep >> (4 * (4 * p + i))) & 15u;
            o.x = ((kb4 & 1u) ? o.x : 0.f) + bs.x;
            o.y = ((kb4 & 2u) ? o.y : 0.f) + bs.y;
            o.z = ((kb4 & 4u) ? o.z : 0.f) + bs.z;
            o.w = ((kb4 & 8u) ? o.w : 0.f) + bs.w;
            if (row < P.m_rows && !GM_ABL(8)) *reinterpret_cast<float4*>(P.out + row * P.ldo + n0 + 16 * p + 4 * sub_c) = o;
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into it / arrive on it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GM_TMEM_COLS)
                 : "memory");
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

static int gemm_launch(const float* a, int64_t lda, int64_t m_rows, int64_t k, const float* wt_hi,
                       const float* wt_lo, int64_t n, const float* act, int64_t ld_act, int32_t group,
                       const float* bias, float* out, int64_t ldo, lgnn_stream_t stream) {
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
#ifdef LGNN_GM_ABLATE
  P.ablate = LGNN_GM_ABLATE_VALUE;
#endif
  P.m_rows = m_rows;
  P.tiles_total = (m_rows + GM_BM - 1) / GM_BM;
  P.n_kb = k_pad / GM_BK;
  P.last_ksteps = (int)((k - (int64_t)(P.n_kb - 1) * GM_BK + 7) / 8);
  P.cl = (int)(n / GM_BN);
  // tiles per cluster: long enough to amortise a CTA's prologue (barriers, TMEM allocation, the 128 KB weight
  // slice, two cluster syncs), short enough that many more clusters than fit the device at once exist — not every
  // SM can join a 4-CTA cluster, and the tail of a few long work items costs more than the prologues of many short
  // ones.  Measured at K = N = 256, 20 M rows: 32 or 256 tiles 15.0 ms, 1,056 tiles 16.2 ms, 4,096 tiles (one
  // persistent wave) 23.5 ms (profiles/r2e_gemm_lab.txt)
  {
    const int64_t resident = sm_count() / P.cl > 0 ? sm_count() / P.cl : 1;
    int64_t tpc = (P.tiles_total + resident * 16 - 1) / (resident * 16);
    if (tpc < 32) tpc = 32;
    if (tpc > 256) tpc = 256;
    if (const char* t = getenv("LGNN_GEMM_TILES_PER_CLUSTER")) { int v = atoi(t); if (v >= 1) tpc = v; }   // lab switch
    P.tiles_per_cluster = (int)tpc;
  }
  P.group = group;
  P.act = act;
  P.ld_act = ld_act;
  P.bias = bias;
  P.out = out;
  P.ldo = ldo;
  CUtensorMap tm_a, tm_bhi, tm_blo;
  int rc;
  if ((rc = encode_2d(&tm_a, a, (uint64_t)k, (uint64_t)m_rows, (uint64_t)lda * 4, GM_BK, GM_BM / P.cl))) return rc;
  if ((rc = encode_2d(&tm_bhi, wt_hi, (uint64_t)k_pad, (uint64_t)n, (uint64_t)k_pad * 4, GM_BK, GM_BN))) return rc;
  if ((rc = encode_2d(&tm_blo, wt_lo, (uint64_t)k_pad, (uint64_t)n, (uint64_t)k_pad * 4, GM_BK, GM_BN))) return rc;
  const size_t smem = 1024 + (size_t)2 * P.n_kb * GM_B_TILE + (size_t)GM_RAW_STAGES * GM_A_TILE + 4 * GM_STG_WARP + 256;
  LGNN_CUDA_TRY(cudaFuncSetAttribute(gemm_mask_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LGNN_CUDA_TRY(cudaFuncSetAttribute(gemm_mask_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_clusters = (P.tiles_total + P.tiles_per_cluster - 1) / P.tiles_per_cluster;
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
  if (act) LGNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_mask_kernel<true>, tm_a, tm_bhi, tm_blo, P));
  else LGNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_mask_kernel<false>, tm_a, tm_bhi, tm_blo, P));
  LGNN_LAUNCH_CHECK("gemm_mask_kernel");
  return LGNN_OK;
}

int lgnn_gemm_mask_f32(const float* a, int64_t lda, int64_t m_rows, int64_t k, const float* wt_hi,
                       const float* wt_lo, int64_t n, const float* act, int64_t ld_act, int32_t group,
                       float* out, int64_t ldo, lgnn_stream_t stream) {
  return gemm_launch(a, lda, m_rows, k, wt_hi, wt_lo, n, act, ld_act, group, nullptr, out, ldo, stream);
}

int lgnn_gemm_bias_f32(const float* a, int64_t lda, int64_t m_rows, int64_t k, const float* wt_hi,
                       const float* wt_lo, int64_t n, const float* bias, float* out, int64_t ldo,
                       lgnn_stream_t stream) {
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return fail(LGNN_E_ALIGN, "gemm_bias: bias must be 16-byte aligned");
  return gemm_launch(a, lda, m_rows, k, wt_hi, wt_lo, n, nullptr, 0, 1, bias, out, ldo, stream);
}

}  // extern "C"

#if defined(LGNN_MBAR_DEBUG) && !defined(LGNN_GM_ABLATE)
// debug builds only: {line, block, thread, barrier, parity, count} of the first timed-out mbarrier wait since the last call
extern "C" int lgnn_debug_mbar_timeout(unsigned int* out8) {
  unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out8, lgnn::g_mbar_timeout, sizeof(z)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(lgnn::g_mbar_timeout, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif
