// C += X^T X on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM),
// with the 3xTF32 split for fp32-faithful results:  x = hi + lo (hi = x with the 13 low mantissa
// bits cleared — exactly the value the tensor core sees — lo = x - hi, exact in fp32), and
//     X^T X  ~=  hi^T hi + hi^T lo + lo^T hi        (the lo^T lo term is below fp32 rounding).
//
// Shape: extreme-K, tiny-N (K = nodes x Hessian columns up to 1e8 rows, n <= 256).  One persistent
// CTA per SM owns a contiguous slice of rows and the FULL n x n accumulator in TMEM (two M=128
// blocks x up to 256 fp32 columns = all 512 TMEM columns), so X is read from HBM exactly once.
//
// Tensor-core accumulation rounds toward zero, so a long all-positive chain (H^T H of relu outputs,
// the diagonal of any Gram matrix) drifts low by ~4e-8 per accumulate.  The chain is therefore cut
// into SEGMENTS of TC_SEG_STEPS steps (192 accumulates): after each segment four dedicated epilogue
// warps drain TMEM and add it to the CTA's fp32 partial in global memory (L2-resident) with
// round-to-nearest adds while the TMA / transform warps keep the operand ring full; the next
// segment restarts the TMEM accumulator from zero.
//
// Only the upper block-triangle is computed: M block 0 (rows 0..127) against all np columns, M block
// 1 (rows 128..255) against columns 128..np-1 only — 25 % fewer MMAs and TMEM columns at n = 256.
//
// Pipeline per CTA (18 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes [BK rows x NP cols] of raw fp32 into a
//               6-deep shared-memory ring (out-of-bounds rows / columns arrive as zeros, which makes
//               every edge case — row tail, n not a multiple of 16 — free);
//   warps 2..9  transform (2..5: features 0..127, 6..9: features 128..255): thread = feature = TMEM lane.  It reads its COLUMN of the raw box (16 values,
//               consecutive lanes -> consecutive words), splits hi / lo and writes the operand twice:
//               (A) into TENSOR MEMORY (tcgen05.st: lane = feature, column = row of the box) for the M side,
//               (B) as K-major no-swizzle UMMA operand rows in shared memory (8-row x 16-byte core matrices,
//               padded 144-byte group stride) for the N side; a 2-deep operand ring;
//   warp 1      the whole warp walks the loop, one elected lane issues, per 8-wide k-step and per M block,
//               three tcgen05.mma with A FROM TMEM and B from shared memory (hi.hi, hi.lo, lo.hi), then
//               tcgen05.commit frees the operand stage.  (Round 1 took both operands from shared memory:
//               120 KB of operand reads per 16-row step on top of the transform's 64 KB is more than the
//               128 B/clk an SM's shared memory delivers in the 1,152 cycles the step's MMAs take — ncu showed
//               8.5-way "bank conflicts" on conflict-free loads and the tensor pipe 57 % busy.  With A in TMEM
//               the MMAs read 72 KB per step.)
//   warps 10..17 epilogue, once per segment: tcgen05.ld the accumulator (two warps per 32-lane TMEM
//               quarter, alternating 16-column chunks) and add it into this CTA's zero-initialised
//               partial n x NP block with fire-and-forget red.global.add.v4.f32 (RN adds in L2; every
//               address is only ever touched by one thread, so the order of adds — and the result —
//               is the same on every run); a second kernel reduces the <=148 partials in fixed
//               order and writes both triangles of C.
//
// Tensor roofline accounting: useful flops = k_rows * n * (n+1); the 3x issue factor
// of the split is not counted as useful work.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "syrk_internal.cuh"
#include "tc_common.cuh"

namespace lgnn {

constexpr int TC_BK = 16;          // rows of X per pipeline stage (two k=8 UMMA steps)
constexpr int TC_RAW_STAGES = 6;
constexpr int TC_OP_STAGES = 2;    // operand stages: B rows in shared memory + A columns in TMEM (32 per M block)
constexpr int TC_SBO = 144;        // byte stride between 8-row core-matrix groups (128 + 16 pad)
constexpr int TC_THREADS = 576;            // 1 TMA + 1 MMA + 8 transform + 8 epilogue warps
constexpr int TC_EPILOGUE_WARPS = 8;
constexpr int TC_SEG_STEPS_DEFAULT = 32;  // 32 steps x 2 k-steps x 3 products = 192 accumulates per chain

struct TcGeom {
  int n, np, mb, npa;   // n, n padded to 16, #128-row M blocks, A-operand rows (mb*128)
  int lbo;              // byte stride between the 16-byte k-chunks of one row group set
  int op_bytes;         // bytes of one (hi or lo) operand stage
  int raw_bytes;        // bytes of one raw stage
  int tmem_cols;        // power of two >= accumulator columns + A operand stages, >= 32
  int a_col0;           // first TMEM column of the A operand stages (behind the accumulators)
  size_t smem_bytes;
};

static TcGeom tc_geom(int n) {
  TcGeom g;
  g.n = n;
  g.np = (n + 15) / 16 * 16;
  g.mb = (g.np + 127) / 128;
  g.npa = g.mb * 128;
  g.lbo = (g.npa / 8) * TC_SBO;
  g.op_bytes = (TC_BK / 4) * g.lbo;
  g.raw_bytes = TC_BK * g.np * 4;
  int cols = g.mb == 1 ? g.np : g.np + (g.np - 128);   // block 1 only holds columns 128..np-1
  cols = (cols + 31) / 32 * 32;                        // A operand stages start on a 32-column boundary
  g.a_col0 = cols;
  cols += TC_OP_STAGES * g.mb * 32;                    // per stage and M block: 16 columns hi + 16 columns lo
  g.tmem_cols = 32;
  while (g.tmem_cols < cols) g.tmem_cols <<= 1;
  g.smem_bytes = 1024 /*align slack*/ + (size_t)TC_RAW_STAGES * g.raw_bytes +
                 (size_t)TC_OP_STAGES * 2 * g.op_bytes + 256 /*barriers*/;
  return g;
}

// Ablation switches for the lab copy of this kernel (syrk_tcgen05_lab.cu compiles this file a second time with
// LGNN_TC_ABLATE defined and the global symbols renamed).  In the product build TC_ABL is the constant false.
//   1: no MMA issued (commits only)   2: only hi.hi of the three products   4: transform warps skip their
//   shared-memory stores of hi / lo   8: epilogue skips the red.global adds
#ifdef LGNN_TC_ABLATE
#define TC_ABL(bit) ((P.ablate & (bit)) != 0)
#else
#define TC_ABL(bit) false
#endif

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

struct TcParams {
#ifdef LGNN_TC_ABLATE
  int ablate;
#endif
  int n, np, mb, lbo, op_bytes, raw_bytes, tmem_cols, a_col0;
  int seg_steps;     // steps per TMEM accumulation segment
  int64_t steps_total;
  int64_t steps_per_cta;
  float* part;  // [gridDim.x][n][np]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
syrk_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams P) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  uint8_t* raw_base = smem;
  uint8_t* op_base = raw_base + (size_t)TC_RAW_STAGES * P.raw_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(op_base + (size_t)TC_OP_STAGES * 2 * P.op_bytes);
  // barrier slots
  uint64_t* full_raw = bars;                       // [RAW]
  uint64_t* empty_raw = bars + TC_RAW_STAGES;      // [RAW]
  uint64_t* full_op = empty_raw + TC_RAW_STAGES;   // [OP]
  uint64_t* empty_op = full_op + TC_OP_STAGES;     // [OP]
  uint64_t* acc_full = empty_op + TC_OP_STAGES;    // [1]
  uint64_t* acc_empty = acc_full + 1;              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = (int)warp_uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int64_t step_beg = (int64_t)blockIdx.x * P.steps_per_cta;
  int64_t step_end = step_beg + P.steps_per_cta;
  if (step_end > P.steps_total) step_end = P.steps_total;
  const int64_t my_steps = step_end > step_beg ? step_end - step_beg : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_RAW_STAGES; ++i) {
      mbar_init(smem_u32(&full_raw[i]), 1);
      mbar_init(smem_u32(&empty_raw[i]), (uint32_t)(4 * P.mb));
    }
    for (int i = 0; i < TC_OP_STAGES; ++i) {
      mbar_init(smem_u32(&full_op[i]), (uint32_t)(4 * P.mb));
      mbar_init(smem_u32(&empty_op[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(acc_empty), TC_EPILOGUE_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      // 32-bit loop state, stages and phases stepped incrementally (no 64-bit division per step)
      const uint32_t raw0 = smem_u32(raw_base);
      int row = (int)(step_beg * TC_BK), st = 0;
      uint32_t ph = 0;
      for (int s = 0; s < (int)my_steps; ++s) {
        mbar_wait(smem_u32(&empty_raw[st]), ph ^ 1);
        uint32_t bar = smem_u32(&full_raw[st]);
        mbar_arrive_expect_tx(bar, (uint32_t)P.raw_bytes);
        tma_load_2d(raw0 + (uint32_t)st * (uint32_t)P.raw_bytes, &tmap, 0, row, bar);
        row += TC_BK;
        if (++st == TC_RAW_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // the WHOLE warp walks the loop (uniform control flow, uniform operands), one elected lane issues each MMA /
    // commit (tc_common.cuh: tc_mma_tf32_e)
    if (my_steps > 0) {
      const uint32_t tb = warp_uniform(tmem_base);
      // instruction descriptor: D fp32, A/B tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc0 = idesc_base | ((uint32_t)(P.np >> 3) << 17);
      const uint32_t idesc1 = idesc_base | ((uint32_t)((P.np - 128) >> 3) << 17);
      const uint32_t lbo = P.lbo, sbo = TC_SBO;
      const uint32_t op0 = smem_u32(op_base);
      int st = 0, in_seg = 0;             // operand stage, step inside the current accumulation segment
      uint32_t ph = 0, seg_ph = 0;        // operand-ring phase, parity of the segments drained so far
      const int n_steps = (int)my_steps;
      for (int s = 0; s < n_steps; ++s) {
        const bool seg_first = in_seg == 0;
        if (seg_first && s > 0) {
          // the epilogue warps have drained the previous segment out of TMEM
          mbar_wait(smem_u32(acc_empty), seg_ph);
          seg_ph ^= 1u;
          tc_fence_after();
        }
        mbar_wait(smem_u32(&full_op[st]), ph);
        tc_fence_after();
        const uint32_t hi = op0 + (uint32_t)st * 2u * (uint32_t)P.op_bytes;
        const uint32_t lo = hi + P.op_bytes;
        const uint32_t a_st = tb + (uint32_t)P.a_col0 + (uint32_t)(st * P.mb * 32);   // A stage: [block][hi 16 | lo 16]
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          const uint32_t koff = (uint32_t)ks * 2u * (uint32_t)P.lbo;  // two 16-byte k-chunks per k=8 step
          const uint32_t acc = (seg_first && ks == 0) ? 0u : 1u;
          {  // M block 0: features 0..127 (TMEM) x features 0..np-1 (shared memory)
            const uint64_t b_hi = make_smem_desc(hi + koff, lbo, sbo);
            const uint64_t b_lo = make_smem_desc(lo + koff, lbo, sbo);
            const uint32_t a_hi = a_st + (uint32_t)(ks * 8), a_lo = a_hi + 16u;
            if (!TC_ABL(1)) tc_mma_tf32_ts_e(tb, a_hi, b_hi, idesc0, acc);
            if (!TC_ABL(1) && !TC_ABL(2)) {
              tc_mma_tf32_ts_e(tb, a_hi, b_lo, idesc0, 1u);
              tc_mma_tf32_ts_e(tb, a_lo, b_hi, idesc0, 1u);
            }
          }
          if (P.mb == 2) {  // M block 1: features 128..255 x features 128..np-1 (B rows 128.. = 16 groups on)
            const uint32_t off = koff + 16u * TC_SBO;
            const uint64_t b_hi = make_smem_desc(hi + off, lbo, sbo);
            const uint64_t b_lo = make_smem_desc(lo + off, lbo, sbo);
            const uint32_t a_hi = a_st + 32u + (uint32_t)(ks * 8), a_lo = a_hi + 16u;
            const uint32_t d = tb + (uint32_t)P.np;
            if (!TC_ABL(1)) tc_mma_tf32_ts_e(d, a_hi, b_hi, idesc1, acc);
            if (!TC_ABL(1) && !TC_ABL(2)) {
              tc_mma_tf32_ts_e(d, a_hi, b_lo, idesc1, 1u);
              tc_mma_tf32_ts_e(d, a_lo, b_hi, idesc1, 1u);
            }
          }
        }
        tc_commit_e(smem_u32(&empty_op[st]));  // implies fence::before_thread_sync
        if (++in_seg == P.seg_steps || s == n_steps - 1) {
          tc_commit_e(smem_u32(acc_full));
          in_seg = 0;
        }
        if (++st == TC_OP_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 10) {
    // ===================================================================== transform warps
    // thread = feature = TMEM lane: warps 2-5 own M block 0 (features 0..127), warps 6-9 M block 1 (features
    // 128..255; idle when n <= 128); warp w holds TMEM lanes 32 (w & 3) .. + 31.  Per step a thread reads column f
    // of the raw box (TC_BK values: consecutive lanes -> consecutive words) and writes them (A) to its TMEM lane and
    // (B) as row f of the K-major shared-memory operand, hi and lo each.  Every step is handled by every active warp,
    // in order.  (One set of four warps doing both blocks was latency-bound on its own chain — wait, 32 loads, split,
    // 16 stores, tcgen05.st, wait::st, fence, arrive: 1,390 cycles per step against 1,152 for the step's MMAs,
    // profiles/r2h_syrk_lab_a_in_tmem.txt.)
    const int m = (warp - 2) >> 2;
    if (m < P.mb) {
      const int f = m * 128 + (warp & 3) * 32 + lane;
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t boff = (uint32_t)(f >> 3) * TC_SBO + (uint32_t)(f & 7) * 16u;
      int rs = 0, os = 0;
      uint32_t rph = 0, oph = 0;
      for (int s = 0; s < (int)my_steps; ++s) {
        mbar_wait(smem_u32(&full_raw[rs]), rph);
        const float* raw = reinterpret_cast<const float*>(raw_base + (size_t)rs * P.raw_bytes);
        uint32_t vh[TC_BK], vl[TC_BK];
        if (f < P.np) {
#pragma unroll
          for (int r = 0; r < TC_BK; ++r) {
            const float x = raw[r * P.np + f];
            const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
            vh[r] = __float_as_uint(h);
            vl[r] = __float_as_uint(x - h);
          }
        } else {
#pragma unroll
          for (int r = 0; r < TC_BK; ++r) vh[r] = vl[r] = 0u;
        }
        // the raw box is in registers
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_raw[rs]));
        mbar_wait(smem_u32(&empty_op[os]), oph ^ 1);
        tc_fence_after();
        if (f < P.np && !TC_ABL(4)) {
          uint8_t* hi = op_base + (size_t)os * 2 * P.op_bytes;
          uint8_t* lo = hi + P.op_bytes;
#pragma unroll
          for (int kq = 0; kq < TC_BK / 4; ++kq) {
            *reinterpret_cast<uint4*>(hi + boff + (uint32_t)kq * (uint32_t)P.lbo) =
                make_uint4(vh[4 * kq], vh[4 * kq + 1], vh[4 * kq + 2], vh[4 * kq + 3]);
            *reinterpret_cast<uint4*>(lo + boff + (uint32_t)kq * (uint32_t)P.lbo) =
                make_uint4(vl[4 * kq], vl[4 * kq + 1], vl[4 * kq + 2], vl[4 * kq + 3]);
          }
        }
        // (A) lane = feature, 16 columns hi then 16 columns lo of this stage and block (warp-collective)
        const uint32_t taddr = tmem_base + lane_addr + (uint32_t)P.a_col0 + (uint32_t)(os * P.mb * 32 + m * 32);
        if (!TC_ABL(4)) {
          tmem_st16(taddr, vh);
          tmem_st16(taddr + 16, vl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&full_op[os]));
        if (++rs == TC_RAW_STAGES) { rs = 0; rph ^= 1u; }
        if (++os == TC_OP_STAGES) { os = 0; oph ^= 1u; }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int half = (warp - 10) >> 2;     // the two warps of a quarter take alternate 16-column chunks
    float* out = P.part + (size_t)blockIdx.x * P.n * P.np;
    const int64_t n_seg = (my_steps + P.seg_steps - 1) / P.seg_steps;
    for (int64_t seg = 0; seg < n_seg; ++seg) {
      mbar_wait(smem_u32(acc_full), (uint32_t)(seg & 1));
      tc_fence_after();
      for (int m = 0; m < P.mb; ++m) {
        const int row = m * 128 + quarter * 32 + lane;
        const int col_beg = m * 128;                       // block 1 starts at column 128
        const uint32_t tcol0 = m == 0 ? 0u : (uint32_t)P.np;
        const bool live = row < P.n;
        for (int c0 = col_beg + 16 * half; c0 < P.np; c0 += 32) {
          uint32_t v[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + tcol0 + (uint32_t)(c0 - col_beg);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                "=r"(v[14]), "=r"(v[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // pin the uses of v[] behind the wait (volatile asms keep their order)
          asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                            "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                            "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
          if (live && !TC_ABL(8)) {
            float* dst = out + (size_t)row * P.np + c0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "r"(v[4 * q]),
                           "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                           : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
  }
}

// C[i][j] (i <= j) = beta*C + alpha * sum_s part[s][i][j], mirrored; fixed order over s.
__global__ void syrk_tc_reduce_kernel(const float* __restrict__ part, int n_parts, int n, int np,
                                      float alpha, float beta, float* __restrict__ c, int64_t ldc) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * np) return;
  int i = e / np, j = e - i * np;
  if (j >= n || j < i) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += part[(size_t)p * n * np + e];
  float v = alpha * s;
  if (beta != 0.f) v = fmaf(beta, c[(int64_t)i * ldc + j], v);
  c[(int64_t)i * ldc + j] = v;
  if (i != j) c[(int64_t)j * ldc + i] = v;
}

// ------------------------------------------------------------------------------ host side
static int tc_grid(int64_t steps_total) {
  int64_t g = sm_count();
  if (g > steps_total) g = steps_total;
  if (g < 1) g = 1;
  return (int)g;
}

bool syrk_tcgen05_supported(int64_t k_rows, int64_t n) {
  return n >= 8 && n <= 256 && k_rows >= 1 && k_rows < ((int64_t)1 << 31) - 64;
}

size_t syrk_tcgen05_workspace_bytes(int64_t k_rows, int64_t n) {
  (void)k_rows;
  TcGeom g = tc_geom((int)n);
  return (size_t)sm_count() * n * g.np * sizeof(float);
}

int syrk_tcgen05_launch(const float* x, int64_t ldx, int64_t k_rows, int n, float alpha, float beta,
                        float* c, int64_t ldc, void* ws, cudaStream_t st) {
  TcGeom g = tc_geom(n);
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(LGNN_E_CUDA, "syrk_tcgen05: cuTensorMapEncodeTiled not available");
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)k_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ldx * 4};
  cuuint32_t box[2] = {(cuuint32_t)g.np, (cuuint32_t)TC_BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LGNN_E_CUDA, "syrk_tcgen05: cuTensorMapEncodeTiled failed (%d)", (int)r);

  TcParams P;
#ifdef LGNN_TC_ABLATE
  P.ablate = LGNN_TC_ABLATE_VALUE;
#endif
  P.n = n; P.np = g.np; P.mb = g.mb; P.lbo = g.lbo; P.op_bytes = g.op_bytes; P.raw_bytes = g.raw_bytes;
  P.tmem_cols = g.tmem_cols;
  P.a_col0 = g.a_col0;
  P.seg_steps = TC_SEG_STEPS_DEFAULT;
  if (const char* sg = getenv("LGNN_SYRK_SEG_STEPS")) {  // accuracy / speed experiments
    int v = atoi(sg);
    if (v >= 1 && v <= (1 << 20)) P.seg_steps = v;
  }
  P.steps_total = (k_rows + TC_BK - 1) / TC_BK;
  int grid = tc_grid(P.steps_total);
  P.steps_per_cta = (P.steps_total + grid - 1) / grid;
  P.part = reinterpret_cast<float*>(ws);
  LGNN_CUDA_TRY(cudaFuncSetAttribute(syrk_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)g.smem_bytes));
  LGNN_CUDA_TRY(cudaMemsetAsync(P.part, 0, (size_t)grid * n * g.np * sizeof(float), st));
  syrk_tcgen05_kernel<<<grid, TC_THREADS, g.smem_bytes, st>>>(tmap, P);
  LGNN_LAUNCH_CHECK("syrk_tcgen05_kernel");
  int elems = n * g.np;
  syrk_tc_reduce_kernel<<<(elems + 255) / 256, 256, 0, st>>>(P.part, grid, n, g.np, alpha, beta, c, ldc);
  LGNN_LAUNCH_CHECK("syrk_tc_reduce_kernel");
  return LGNN_OK;
}

}  // namespace lgnn
