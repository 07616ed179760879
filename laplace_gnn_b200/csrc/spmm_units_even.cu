// Unit-compacted slabs for column groups of g = 2, 6, 10 or 14 (g % 4 == 2): the same layout and the same
// staged SpMM as spmm_units.cu, for the groups a multiple-of-4 rule would pad by up to 33 %.
//
// Why: the column-parallel multi-GPU backward gives each of 8 ranks 6 (or 5) of the 47 Hessian-sqrt columns
// of the products shape.  Padded to 8 they cost every rank 8 columns' worth of gathers, GEMM rows and SYRK
// rows; at g = 6 the rank moves what it owns.
//
// What changes against g % 4 == 0: a slot is g*4 = 8*G2 bytes (G2 = g/2 odd), so the run of a 32-unit block
// would start on an 8-byte boundary whenever the units before it are an odd number — and the staged kernel
// copies runs with 16-byte cp.async.  The pack kernel therefore starts every block's run on an EVEN slot
// (hdr.first = sum of the previous blocks' live counts, each rounded up to even): at most one unused slot per
// block, and a row still fits its dense pitch because a block never holds more than 32 slots either way.  A run
// with an odd live count is copied with its last 16-byte piece half stale; no lane reads that half.
// The consumer lanes read their slot with 64-bit shared-memory loads: slot stride 2*G2 words with G2 odd maps
// the 16 lanes of a half-warp onto 16 distinct bank pairs — conflict-free without a swizzle.
//
// The arithmetic is the dense kernel's (one fmaf per (neighbour, element), neighbours in CSR order): the
// result is bit-identical to lgnn_spmm_f32 on the uncompacted slab, as for the other group sizes.
#include "common.cuh"
#include "spmm_internal.cuh"

namespace lgnn {

namespace {

constexpr int EVEN_THREADS = 256;

template <int G2>
__global__ void __launch_bounds__(1024) unit_pack_even_kernel(const UnitPackArgs A) {
  __shared__ uint32_t cnt[32];
  const int u = threadIdx.x, lane = u & 31, w = u >> 5;
  const int64_t row = blockIdx.x;
  const int64_t r0 = A.row_first ? __ldg(A.row_first + row) : 0;       // ragged: absolute (even) slot of the row
  const bool on = __ldg(A.act + row * A.lda + u) > 0.f;
  const uint32_t m = __ballot_sync(0xffffffffu, on);
  if (lane == 0) cnt[w] = ((uint32_t)__popc(m) + 1u) & ~1u;      // slots this block takes: live count rounded up to even
  float v[2 * G2];
  if (A.src) {
    const float* base = A.src + row * A.lds;
#pragma unroll
    for (int c = 0; c < 2 * G2; ++c) v[c] = base[(int64_t)c * A.h + u];
  }
  __syncthreads();                      // every dense value is in registers before the row is overwritten
  uint32_t first = 0;
  for (int b = 0; b < w; ++b) first += cnt[b];
  if (lane == 0 && A.hdr) A.hdr[row * (A.h >> 5) + w] = make_uint2(m, (uint32_t)r0 + first);
  if (on && A.src) {
    float* out = A.row_first ? A.dst + r0 * (2 * G2) : A.dst + row * A.ldd;
    float2* dst = reinterpret_cast<float2*>(out + (int64_t)(first + __popc(m & ((1u << lane) - 1u))) * (2 * G2));
#pragma unroll
    for (int t = 0; t < G2; ++t) dst[t] = make_float2(v[2 * t], v[2 * t + 1]);
  }
}

__device__ __forceinline__ void cp_async16_e(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_e() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_e() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void fma2(float2& a, float v, const float2& x) {
  a.x = fmaf(v, x.x, a.x);
  a.y = fmaf(v, x.y, a.y);
}

// The structure of spmm_units_staged_kernel (spmm_units.cu): a group of warps owns `rpg` consecutive rows,
// walks their (col, val) run in batches of 32 with the next batch's headers and the one after's (col, val) in
// flight, and copies every neighbour's run into a per-warp ring of U slots with cp.async.
//
// NB = unit blocks per warp (1 or 2).  The cost of one (warp, neighbour) iteration — shuffles, the copy issue,
// the wait, two warp syncs — does not depend on g, so narrow groups move few bytes per iteration: at g = 6 a
// warp's run is 16 live units x 24 bytes = 384 bytes and the kernel reached 0.67 of the copy rate against 0.88
// at g = 12 (profiles/r2a_units_lab.txt).  With NB = 2 a lane owns the units 32(2w) + L and 32(2w+1) + L: the
// runs of two consecutive blocks are adjacent in the row (block 2w+1 starts where block 2w's slots, rounded up
// to even, end), so the warp copies ONE run of twice the length per neighbour and reads both header words
// with one 16-byte load — g = 6 then moves what g = 12 moves per iteration.  (Four blocks per warp were no better: 84.7
// against 82.4 ms at g = 6, profiles/r2n_units_lab.txt — occupancy drops to 16 warps per SM.)
template <int G2, int NB, int U, int MINB>
__global__ void __launch_bounds__(EVEN_THREADS, MINB) spmm_units_even_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ slab, int64_t lds, const uint2* __restrict__ hdr,
    int nblk, float* __restrict__ y, int64_t ldy, int rpg) {
  extern __shared__ __align__(16) uint8_t units_even_smem_[];
  constexpr int SLOT = 256 * G2 * NB;                  // bytes: NB x 32 live units x g floats
  constexpr int PIECES = (16 * G2 * NB + 31) / 32;     // 16-byte pieces per lane for a full run
  const int lane = threadIdx.x & 31;
  const uint32_t ring = smem_addr(units_even_smem_) + (threadIdx.x >> 5) * (U * SLOT);
  const int64_t warp = ((int64_t)blockIdx.x * EVEN_THREADS + threadIdx.x) >> 5;
  const int wpr = nblk / NB;                           // warps per row group
  const int64_t grp = warp / wpr;
  const int64_t r0 = grp * rpg;
  if (r0 >= n_rows) return;
  const int w = (int)(warp - grp * wpr);
  const int nr = (int)((n_rows - r0) < rpg ? (n_rows - r0) : rpg);
  const uint32_t lt = (1u << lane) - 1u;
  const int h = nblk * 32;

  const int64_t my_rp64 = (lane <= nr) ? __ldg(rowptr + r0 + lane) : 0;      // rpg <= 31
  const int64_t k_beg = __shfl_sync(0xffffffffu, my_rp64, 0);
  const int my_rp = (int)(my_rp64 - k_beg);
  const int k_end = __shfl_sync(0xffffffffu, my_rp, nr);
  col += k_beg;
  val += k_beg;
  hdr += w * NB;
  float* yb = y + r0 * ldy + 32 * NB * w + lane;
  const float4* slab4 = reinterpret_cast<const float4*>(slab);

  float2 acc[NB][G2];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int t = 0; t < G2; ++t) acc[b][t] = make_float2(0.f, 0.f);

  auto load_cv = [&](int k0, int32_t& c, float& v) {
    const int k = k0 + lane;
    c = -1;
    v = 0.f;
    if (k < k_end) {
      c = __ldg(col + k);
      v = __ldg(val + k);
    }
  };
  // slab position of a neighbour's run in 16-byte units: the block's first slot is even, a slot is G2 half-pieces
  auto load_hdr = [&](int32_t c, uint32_t (&m)[NB], uint32_t& p) {
#pragma unroll
    for (int b = 0; b < NB; ++b) m[b] = 0u;
    p = 0u;
    if (c >= 0) {
      uint32_t first;
      if (NB >= 2) {
        const uint4* h4 = reinterpret_cast<const uint4*>(hdr + (int64_t)c * nblk);   // blocks NB w .. NB w + NB - 1
        const uint4 hd = __ldg(h4);
        m[0] = hd.x;
        first = hd.y;
        m[1 % NB] = hd.z;
        if (NB == 4) {
          const uint4 hd2 = __ldg(h4 + 1);
          m[2 % NB] = hd2.x;
          m[3 % NB] = hd2.z;
        }
      } else {
        const uint2 hd = __ldg(hdr + (int64_t)c * nblk);
        m[0] = hd.x;
        first = hd.y;
      }
      p = (uint32_t)(((int64_t)c * lds) >> 2) + (first >> 1) * G2;
    }
  };
  int row = 0;
  auto flush = [&]() {
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int t = 0; t < G2; ++t) {
        yb[32 * b + (2 * t + 0) * h] = acc[b][t].x;
        yb[32 * b + (2 * t + 1) * h] = acc[b][t].y;
        acc[b][t] = make_float2(0.f, 0.f);
      }
    yb += ldy;
    ++row;
  };

  int32_t c_a;                 // batch b+1: (col, val) loaded, header not yet
  float v_a;
  float my_v;                  // batch b: everything loaded
  uint32_t my_m[NB], my_p;
  load_cv(0, c_a, v_a);
  my_v = v_a;
  load_hdr(c_a, my_m, my_p);
  load_cv(32, c_a, v_a);
  int row_end = __shfl_sync(0xffffffffu, my_rp, 1);

  for (int k0 = 0; k0 < k_end; k0 += 32) {
    uint32_t m_n[NB], p_n;
    const float v_n = v_a;
    load_hdr(c_a, m_n, p_n);
    load_cv(k0 + 64, c_a, v_a);
    const int cnt = (k_end - k0) < 32 ? (k_end - k0) : 32;

    uint32_t slot_i = ring, slot_c = ring;             // ring slot of the next copy / of the next neighbour read
    auto issue = [&](int jj) {
      if (jj < cnt) {
        const uint32_t p = __shfl_sync(0xffffffffu, my_p, jj);
        int slots = __popc(__shfl_sync(0xffffffffu, my_m[0], jj));
#pragma unroll
        for (int b = 1; b < NB; ++b) slots = ((slots + 1) & ~1) + __popc(__shfl_sync(0xffffffffu, my_m[b], jj));
        const int n16 = (slots * G2 + 1) >> 1;         // 16-byte pieces covering the run (the last may be half stale)
#pragma unroll
        for (int t = 0; t < PIECES; ++t) {
          const int idx = lane + 32 * t;
          if (idx < n16) cp_async16_e(slot_i + 16u * idx, slab4 + (p + idx));
        }
      }
      cp_async_commit_e();
      slot_i = (slot_i + SLOT == ring + U * SLOT) ? ring : slot_i + SLOT;
    };
#pragma unroll
    for (int jj = 0; jj < U - 1; ++jj) issue(jj);
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      issue(j + U - 1);
      cp_async_wait_e<U - 1>();
      __syncwarp();
      while (k0 + j == row_end) {           // rows ending here (empty rows flush zeros)
        flush();
        row_end = __shfl_sync(0xffffffffu, my_rp, row + 1);
      }
      const float v = __shfl_sync(0xffffffffu, my_v, j);
      int base = 0;                          // first slot of block b inside the run (every block starts on an even slot)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint32_t mb = __shfl_sync(0xffffffffu, my_m[b], j);
        if ((mb >> lane) & 1u) {
          const uint32_t at = slot_c + 8u * (uint32_t)((base + __popc(mb & lt)) * G2);
#pragma unroll
          for (int t = 0; t < G2; ++t) fma2(acc[b][t], v, lds_f2(at + 8u * t));
        }
        base += (__popc(mb) + 1) & ~1;
      }
      __syncwarp();                         // the slot is rewritten by the copy issued next iteration
      slot_c = (slot_c + SLOT == ring + U * SLOT) ? ring : slot_c + SLOT;
    }
    my_v = v_n;
#pragma unroll
    for (int b = 0; b < NB; ++b) my_m[b] = m_n[b];
    my_p = p_n;
  }
  while (row < nr) flush();
}

template <int G2, int NB, int U, int MINB>
int launch_even(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val, const float* slab,
                int64_t lds, const uint2* hdr, int nblk, float* y, int64_t ldy, int rpg, cudaStream_t st) {
  const int64_t warps = (n_rows + rpg - 1) / rpg * (nblk / NB);
  const int64_t blocks = (warps + EVEN_THREADS / 32 - 1) / (EVEN_THREADS / 32);
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_units: grid too large");
  const int smem = (EVEN_THREADS / 32) * U * 256 * G2 * NB;
  LGNN_CUDA_TRY(cudaFuncSetAttribute(spmm_units_even_kernel<G2, NB, U, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  spmm_units_even_kernel<G2, NB, U, MINB><<<(unsigned)blocks, EVEN_THREADS, smem, st>>>(n_rows, rowptr, col, val, slab,
                                                                                       lds, hdr, nblk, y, ldy, rpg);
  LGNN_LAUNCH_CHECK("spmm_units_even_kernel");
  return LGNN_OK;
}

}  // namespace

int unit_pack_even(const UnitPackArgs& A, int64_t n_rows, int g, cudaStream_t st) {
  const unsigned grid = (unsigned)n_rows, block = (unsigned)A.h;
  switch (g) {
    case 2: unit_pack_even_kernel<1><<<grid, block, 0, st>>>(A); break;
    case 6: unit_pack_even_kernel<3><<<grid, block, 0, st>>>(A); break;
    case 10: unit_pack_even_kernel<5><<<grid, block, 0, st>>>(A); break;
    case 14: unit_pack_even_kernel<7><<<grid, block, 0, st>>>(A); break;
    default: return fail(LGNN_E_UNSUPPORTED, "unit_pack: g = %d is not 2, 6, 10 or 14", g);
  }
  LGNN_LAUNCH_CHECK("unit_pack_even_kernel");
  return LGNN_OK;
}

// variant: 0 / 8..11 = 4 rows per group, 12..15 = 8 rows per group; the low two bits pick the ring depth; bit 4
// (16..31) forces one unit block per warp where the default is two (g = 2, 6 with an even number of blocks)
int spmm_units_even(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val, const float* slab,
                    int64_t lds, const uint2* hdr, int g, int nblk, float* y, int64_t ldy, int variant,
                    cudaStream_t st) {
  const int rpg = 4 << ((variant >> 2) & 1);
  const int cfg = variant & 3;
  const bool pair = (nblk % 2 == 0) && !(variant & 16) && (reinterpret_cast<uintptr_t>(hdr) & 15) == 0;
#define LGNN_EVEN(G2_, NB_, U_, MINB_) launch_even<G2_, NB_, U_, MINB_>(n_rows, rowptr, col, val, slab, lds, hdr, nblk, y, ldy, rpg, st)
  switch (g) {
    case 2: return pair ? LGNN_EVEN(1, 2, 8, 3) : LGNN_EVEN(1, 1, 8, 3);
    case 6:
      if (pair) switch (cfg) {
        case 1: return LGNN_EVEN(3, 2, 4, 3);
        case 2: return LGNN_EVEN(3, 2, 8, 2);
        case 3: return LGNN_EVEN(3, 2, 5, 3);
        default: return LGNN_EVEN(3, 2, 6, 3);
      }
      switch (cfg) {
        case 1: return LGNN_EVEN(3, 1, 8, 4);
        case 2: return LGNN_EVEN(3, 1, 4, 4);
        default: return LGNN_EVEN(3, 1, 6, 4);
      }
    case 10:
      switch (cfg) {
        case 1: return LGNN_EVEN(5, 1, 6, 3);
        default: return LGNN_EVEN(5, 1, 4, 3);
      }
    case 14:
      switch (cfg) {
        case 1: return LGNN_EVEN(7, 1, 6, 2);
        default: return LGNN_EVEN(7, 1, 4, 3);
      }
    default: return fail(LGNN_E_UNSUPPORTED, "spmm_units: g = %d is not 2, 6, 10 or 14", g);
  }
#undef LGNN_EVEN
}

}  // namespace lgnn
