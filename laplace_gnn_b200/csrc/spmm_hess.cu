// Output-layer SpMM of the KFAC backward with the Hessian-sqrt right-hand sides generated on the fly.
//
// The first SpMM of every column group is gZ_L = Â^T delta_L with delta_L[j, c, :] = v_{j,c}, the c-th column of
// the softmax Hessian square root at train node j (curvlinops/kfac_utils.py:122-126 through kfac.py:637-661).
// lgnn_hess_rhs_f32 materialises those rows (g * C floats per node) and lgnn_spmm_f32 gathers them: 3 KB per edge
// at g = 16, C = 47.  But v_{j,c} is a closed form of the node's softmax alone, and a rank-2 one:
//
//     v_c[k] = sqrt(p_c) [ (d_ck - p_k)(1 + (f_c - fbar)/2) - p_k (f_k - fbar) ]                (fork)
//            = -( A_c P_k + S_c Q_k )   for k != c,          V_c   for k == c
//     P_k = p_k,  Q_k = p_k (f_k - fbar),  A_c = sqrt(p_c)(1 + (f_c - fbar)/2),  S_c = sqrt(p_c)
//     (textbook GGN mode: A_c = S_c = sqrt(p_c), Q = 0)
//
// so the SpMM only has to gather P, Q (2 C floats) and the group's A, S, V (3 g floats) per edge — 576 bytes
// instead of 3072 — and rebuild the g * C products in registers: 2 FMAs per (column, class) and neighbour.
// lgnn_hess_stats_f32 writes the five vectors once per fit ([N, 5 * Cp] floats, Cp = C rounded up to 4; zero rows
// for nodes outside the batch; A, S, V are accumulated atomically so that a node listed twice in the batch counts
// twice, like autograd's scatter and like lgnn_hess_rhs_f32).  The diagonal element V_c is stored as
// lgnn_hess_rhs_f32 computes it ((1 - p_c) first), not as the difference of two large terms.
//
// One warp per output row; lane = class (and class + 32), G columns x KH halves of accumulators in registers; the
// neighbour's a*A_c, a*S_c are broadcast to the warp through a double-buffered shared-memory line.  Rows are
// walked whole by one warp: meant for graphs without hub rows, like the unit-compacted SpMM.
#include "common.cuh"
#include "spmm_internal.cuh"

namespace lgnn {

namespace {

constexpr int SH_THREADS = 256;

__device__ __forceinline__ float sh_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sh_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// warp per train sample; the softmax arithmetic is that of hess_rhs_kernel (hess.cu)
__global__ void __launch_bounds__(SH_THREADS) hess_stats_kernel(
    const float* __restrict__ logits, int64_t ld, int C, int Cp, const int64_t* __restrict__ idx, int64_t m,
    int mode, float* __restrict__ stats, int64_t lds) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* p = sh + (size_t)w * C;
  int64_t s = ((int64_t)blockIdx.x * SH_THREADS + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * SH_THREADS) >> 5;
  for (; s < m; s += stride) {
    const int64_t node = idx[s];
    const float* f = logits + node * ld;
    float mx = -INFINITY;
    for (int k = lane; k < C; k += 32) mx = fmaxf(mx, f[k]);
    mx = sh_warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < C; k += 32) {
      float e = expf(f[k] - mx);
      p[k] = e;
      se += e;
    }
    se = sh_warp_sum(se);
    float fbar = 0.f;
    for (int k = lane; k < C; k += 32) {
      float pk = p[k] / se;
      p[k] = pk;
      fbar = fmaf(pk, f[k], fbar);
    }
    fbar = sh_warp_sum(fbar);
    float* out = stats + node * lds;
    for (int k = lane; k < C; k += 32) {
      const float pk = p[k];
      const float fck = f[k] - fbar;
      const float spc = sqrtf(pk);
      const float a = (mode == LGNN_HESS_REFERENCE) ? fmaf(0.5f, fck, 1.0f) : 1.0f;
      float v = (1.0f - pk) * a;                          // the k == c element, as hess_rhs_kernel computes it
      if (mode == LGNN_HESS_REFERENCE) v -= pk * fck;
      out[k] = pk;
      out[Cp + k] = (mode == LGNN_HESS_REFERENCE) ? pk * fck : 0.f;
      atomicAdd(out + 2 * Cp + k, spc * a);
      atomicAdd(out + 3 * Cp + k, spc);
      atomicAdd(out + 4 * Cp + k, spc * v);
    }
    __syncwarp();
  }
}

template <int G, int KH>
__global__ void __launch_bounds__(SH_THREADS, 3) spmm_hess_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ stats, int64_t lds, int Cp, int c0, int ncols,
    int width, float* __restrict__ y, int64_t ldy) {
  __shared__ __align__(16) float bc_[SH_THREADS / 32][2][2 * G];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t row = ((int64_t)blockIdx.x * SH_THREADS + threadIdx.x) >> 5;
  if (row >= n_rows) return;

  float acc[G][KH];
#pragma unroll
  for (int c = 0; c < G; ++c)
#pragma unroll
    for (int h = 0; h < KH; ++h) acc[c][h] = 0.f;
  float dacc = 0.f;                                   // lane c < ncols: sum_j a_ij V_{j, c0 + c}
  const bool col_lane = lane < ncols;                 // this lane fetches A, S, V of column c0 + lane
  int par = 0;

  const int64_t beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  for (int64_t k0 = beg; k0 < end; k0 += 32) {
    const int cnt = (int)((end - k0) < 32 ? (end - k0) : 32);
    float my_a = 0.f;
    int32_t my_c = 0;
    if (lane < cnt) {
      my_c = __ldg(col + k0 + lane);
      my_a = __ldg(val + k0 + lane);
    }
    // one neighbour's operands: a, P / Q at this lane's classes, a*A / a*S / a*V at this lane's column
    auto fetch = [&](int j, float& a, float (&P)[KH], float (&Q)[KH], float& al, float& sl, float& vd) {
      a = __shfl_sync(0xffffffffu, my_a, j & 31);
      const int32_t cj = __shfl_sync(0xffffffffu, my_c, j & 31);
      if (j >= cnt) a = 0.f;
#pragma unroll
      for (int h = 0; h < KH; ++h) P[h] = Q[h] = 0.f;
      al = sl = vd = 0.f;
      if (a != 0.f) {                                 // masked edges (sources outside the batch) pull nothing
        const float* sp = stats + (int64_t)cj * lds;
#pragma unroll
        for (int h = 0; h < KH; ++h) {
          const int k = lane + 32 * h;
          if (k < Cp) {
            P[h] = __ldg(sp + k);
            Q[h] = __ldg(sp + Cp + k);
          }
        }
        if (col_lane) {
          al = a * __ldg(sp + 2 * Cp + c0 + lane);
          sl = a * __ldg(sp + 3 * Cp + c0 + lane);
          vd = a * __ldg(sp + 4 * Cp + c0 + lane);
        }
      }
    };
    float a_c, P_c[KH], Q_c[KH], al_c, sl_c, vd_c;
    fetch(0, a_c, P_c, Q_c, al_c, sl_c, vd_c);
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      float a_n, P_n[KH], Q_n[KH], al_n, sl_n, vd_n;
      fetch(j + 1, a_n, P_n, Q_n, al_n, sl_n, vd_n);  // j + 1 == cnt fetches nothing
      if (a_c != 0.f) {                               // uniform across the warp: a_c is a broadcast value
        float* bc = bc_[wi][par];
        if (lane < G) {
          bc[lane] = al_c;
          bc[G + lane] = sl_c;
        }
        dacc += vd_c;
        __syncwarp();
#pragma unroll
        for (int c4 = 0; c4 < G / 4; ++c4) {
          const float4 A4 = *reinterpret_cast<const float4*>(bc + 4 * c4);
          const float4 S4 = *reinterpret_cast<const float4*>(bc + G + 4 * c4);
          const float Av[4] = {A4.x, A4.y, A4.z, A4.w};
          const float Sv[4] = {S4.x, S4.y, S4.z, S4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int h = 0; h < KH; ++h) {
              acc[4 * c4 + i][h] = fmaf(Av[i], P_c[h], acc[4 * c4 + i][h]);
              acc[4 * c4 + i][h] = fmaf(Sv[i], Q_c[h], acc[4 * c4 + i][h]);
            }
        }
        par ^= 1;                                     // the other line is free: everybody passed the last barrier
      }
      a_c = a_n; al_c = al_n; sl_c = sl_n; vd_c = vd_n;
#pragma unroll
      for (int h = 0; h < KH; ++h) { P_c[h] = P_n[h]; Q_c[h] = Q_n[h]; }
    }
  }

  float* yr = y + row * ldy;
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const float d = __shfl_sync(0xffffffffu, dacc, c);          // executed by all lanes, G <= 32
    if (c < width) {
#pragma unroll
      for (int h = 0; h < KH; ++h) {
        const int k = lane + 32 * h;
        if (k < Cp) {
          float o = 0.f;
          if (c < ncols) o = (k == c0 + c) ? d : -acc[c][h];
          yr[c * Cp + k] = o;
        }
      }
    }
  }
}

template <int G, int KH>
int launch_spmm_hess(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val, const float* stats,
                     int64_t lds, int Cp, int c0, int ncols, int width, float* y, int64_t ldy, cudaStream_t st) {
  const int64_t blocks = (n_rows + SH_THREADS / 32 - 1) / (SH_THREADS / 32);
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_hess: grid too large");
  spmm_hess_kernel<G, KH><<<(unsigned)blocks, SH_THREADS, 0, st>>>(n_rows, rowptr, col, val, stats, lds, Cp, c0, ncols,
                                                                  width, y, ldy);
  LGNN_LAUNCH_CHECK("spmm_hess_kernel");
  return LGNN_OK;
}

// ------------------------------------------------------------------------------------------
// Staged variant (flags bit 8).  The kernel above keeps ONE neighbour's operands in flight per warp (they travel
// through registers); at 24 warps per SM that is ~28 KB of gathers in flight, short of what the HBM latency asks
// for.  Here a neighbour's record — P | Q (2 Cp floats, contiguous in the stats row) and the group's slices of
// A, S, V (3 * G floats) — is copied by cp.async (16 bytes per lane) into a per-warp ring of U slots, U - 1
// neighbours in flight without holding registers, the pattern of spmm_units_staged_kernel.  The lanes then read
// P, Q at their classes, A and S as broadcast 128-bit loads, V at their column, and scale P, Q by the edge value.
// Needs c0 % 4 == 0 (16-byte aligned slices).  Columns >= ncols may read stale ring bytes: their accumulators are
// discarded at the store.
__device__ __forceinline__ void sh_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sh_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sh_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int G, int KH, int U>
__global__ void __launch_bounds__(SH_THREADS, 3) spmm_hess_staged_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ stats, int64_t lds, int Cp, int c0, int ncols,
    int width, float* __restrict__ y, int64_t ldy) {
  extern __shared__ __align__(16) uint8_t sh_ring_[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t row = ((int64_t)blockIdx.x * SH_THREADS + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int slot_floats = 2 * Cp + 3 * G;                       // multiple of 4
  float* ring = reinterpret_cast<float*>(sh_ring_) + (size_t)wi * U * slot_floats;
  const uint32_t ring_addr = smem_addr(ring);
  const int pq_pieces = Cp >> 1;                                // 2 Cp floats / 4
  constexpr int GP = G / 4;                                     // pieces per A / S / V slice

  float acc[G][KH];
#pragma unroll
  for (int c = 0; c < G; ++c)
#pragma unroll
    for (int h = 0; h < KH; ++h) acc[c][h] = 0.f;
  float dacc = 0.f;

  const int64_t beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  for (int64_t k0 = beg; k0 < end; k0 += 32) {
    const int cnt = (int)((end - k0) < 32 ? (end - k0) : 32);
    float my_a = 0.f;
    int32_t my_c = 0;
    if (lane < cnt) {
      my_c = __ldg(col + k0 + lane);
      my_a = __ldg(val + k0 + lane);
    }
    int slot_i = 0, slot_c = 0;
    auto issue = [&](int j) {
      if (j < cnt) {
        const float a = __shfl_sync(0xffffffffu, my_a, j);
        const int32_t cj = __shfl_sync(0xffffffffu, my_c, j);
        if (a != 0.f) {                                         // masked edges copy nothing
          const float* sp = stats + (int64_t)cj * lds;
          const uint32_t dst = ring_addr + (uint32_t)(slot_i * slot_floats) * 4u;
          if (lane < pq_pieces) sh_cp_async16(dst + 16u * lane, sp + 4 * lane);
          if (lane < 3 * GP) {
            const int sl = lane / GP, pc = lane - sl * GP;
            if (c0 + 4 * pc < Cp)
              sh_cp_async16(dst + 4u * (uint32_t)(2 * Cp + sl * G + 4 * pc), sp + (2 + sl) * Cp + c0 + 4 * pc);
          }
        }
      }
      sh_cp_async_commit();
      slot_i = (slot_i + 1 == U) ? 0 : slot_i + 1;
    };
#pragma unroll
    for (int j = 0; j < U - 1; ++j) issue(j);
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      issue(j + U - 1);
      sh_cp_async_wait<U - 1>();
      __syncwarp();
      const float a = __shfl_sync(0xffffffffu, my_a, j);
      if (a != 0.f) {                                           // uniform across the warp
        const float* rec = ring + slot_c * slot_floats;
        float Pa[KH], Qa[KH];
#pragma unroll
        for (int h = 0; h < KH; ++h) {
          const int k = lane + 32 * h;
          Pa[h] = (k < Cp) ? a * rec[k] : 0.f;
          Qa[h] = (k < Cp) ? a * rec[Cp + k] : 0.f;
        }
        if (lane < ncols) dacc = fmaf(a, rec[2 * Cp + 2 * G + lane], dacc);
#pragma unroll
        for (int c4 = 0; c4 < GP; ++c4) {
          const float4 A4 = *reinterpret_cast<const float4*>(rec + 2 * Cp + 4 * c4);
          const float4 S4 = *reinterpret_cast<const float4*>(rec + 2 * Cp + G + 4 * c4);
          const float Av[4] = {A4.x, A4.y, A4.z, A4.w};
          const float Sv[4] = {S4.x, S4.y, S4.z, S4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int h = 0; h < KH; ++h) {
              acc[4 * c4 + i][h] = fmaf(Av[i], Pa[h], acc[4 * c4 + i][h]);
              acc[4 * c4 + i][h] = fmaf(Sv[i], Qa[h], acc[4 * c4 + i][h]);
            }
        }
      }
      __syncwarp();                                             // the slot is rewritten by the copy issued next iteration
      slot_c = (slot_c + 1 == U) ? 0 : slot_c + 1;
    }
  }

  float* yr = y + row * ldy;
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const float d = __shfl_sync(0xffffffffu, dacc, c);
    if (c < width) {
#pragma unroll
      for (int h = 0; h < KH; ++h) {
        const int k = lane + 32 * h;
        if (k < Cp) {
          float o = 0.f;
          if (c < ncols) o = (k == c0 + c) ? d : -acc[c][h];
          yr[c * Cp + k] = o;
        }
      }
    }
  }
}

template <int G, int KH>
int launch_spmm_hess_staged(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                            const float* stats, int64_t lds, int Cp, int c0, int ncols, int width, float* y,
                            int64_t ldy, cudaStream_t st) {
  constexpr int U = 4;
  const int64_t blocks = (n_rows + SH_THREADS / 32 - 1) / (SH_THREADS / 32);
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_hess: grid too large");
  const int smem = (SH_THREADS / 32) * U * (2 * Cp + 3 * G) * 4;
  spmm_hess_staged_kernel<G, KH, U><<<(unsigned)blocks, SH_THREADS, smem, st>>>(n_rows, rowptr, col, val, stats, lds, Cp,
                                                                              c0, ncols, width, y, ldy);
  LGNN_LAUNCH_CHECK("spmm_hess_staged_kernel");
  return LGNN_OK;
}

bool spmm_hess_shape_ok(int64_t C, int64_t width) { return C >= 1 && C <= 64 && width >= 1 && width <= 16; }

}  // namespace

}  // namespace lgnn

using namespace lgnn;

extern "C" int lgnn_spmm_hess_supported(int64_t C, int64_t width) { return spmm_hess_shape_ok(C, width) ? 1 : 0; }

extern "C" int lgnn_hess_stats_f32(const float* logits, int64_t ld, int32_t C, const int64_t* idx, int64_t m, int mode,
                                   float* stats, int64_t ld_stats, lgnn_stream_t stream) {
  const int Cp = (C + 3) / 4 * 4;
  if (!logits || !stats || C < 1 || ld < C || m < 0 || ld_stats < 5 * (int64_t)Cp)
    return fail(LGNN_E_BADARG, "hess_stats: bad argument");
  if (mode != LGNN_HESS_REFERENCE && mode != LGNN_HESS_GGN) return fail(LGNN_E_BADARG, "hess_stats: unknown mode %d", mode);
  if (m == 0) return LGNN_OK;
  if (!idx) return fail(LGNN_E_BADARG, "hess_stats: null idx");
  const size_t smem = (size_t)(SH_THREADS / 32) * C * sizeof(float);
  if (smem > 48 * 1024) return fail(LGNN_E_UNSUPPORTED, "hess_stats: C=%d too large", C);
  int64_t blocks = (m + SH_THREADS / 32 - 1) / (SH_THREADS / 32);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  hess_stats_kernel<<<(unsigned)blocks, SH_THREADS, smem, as_stream(stream)>>>(logits, ld, C, Cp, idx, m, mode, stats,
                                                                             ld_stats);
  LGNN_LAUNCH_CHECK("hess_stats_kernel");
  return LGNN_OK;
}

extern "C" int lgnn_spmm_hess_f32(int64_t n_rows, int64_t nnz, const int64_t* rowptr, const int32_t* col,
                                  const float* val, const float* stats, int64_t ld_stats, int32_t C, int32_t c0,
                                  int32_t ncols, int32_t width, float* y, int64_t ldy, int flags,
                                  lgnn_stream_t stream) {
  if (n_rows < 0 || nnz < 0) return fail(LGNN_E_BADARG, "spmm_hess: negative size");
  if (!spmm_hess_shape_ok(C, width)) return fail(LGNN_E_UNSUPPORTED, "spmm_hess: C must be 1 .. 64 and the group at most 16 columns wide (C=%d width=%d)", (int)C, (int)width);
  const int Cp = (C + 3) / 4 * 4;
  if (c0 < 0 || ncols < 0 || ncols > width || c0 + ncols > C || ld_stats < 5 * (int64_t)Cp || ldy < (int64_t)width * Cp)
    return fail(LGNN_E_BADARG, "spmm_hess: bad column range or pitch");
  if (n_rows == 0) return LGNN_OK;
  if (!rowptr || !stats || !y) return fail(LGNN_E_BADARG, "spmm_hess: null pointer");
  if (nnz > 0 && (!col || !val)) return fail(LGNN_E_BADARG, "spmm_hess: null col / val");
  cudaStream_t st = as_stream(stream);
  const int g4 = (width + 3) / 4;
  // flags bit 8: the staged kernel (cp.async ring); needs 16-byte aligned rows and column slices
  if ((flags & 0x100) && c0 % 4 == 0 && ld_stats % 4 == 0 && (reinterpret_cast<uintptr_t>(stats) & 15) == 0) {
#define LGNN_SHS(G_) (Cp <= 32 ? launch_spmm_hess_staged<G_, 1>(n_rows, rowptr, col, val, stats, ld_stats, Cp, c0, ncols, width, y, ldy, st) \
                               : launch_spmm_hess_staged<G_, 2>(n_rows, rowptr, col, val, stats, ld_stats, Cp, c0, ncols, width, y, ldy, st))
    switch (g4) {
      case 1: return LGNN_SHS(4);
      case 2: return LGNN_SHS(8);
      case 3: return LGNN_SHS(12);
      default: return LGNN_SHS(16);
    }
#undef LGNN_SHS
  }
#define LGNN_SH(G_) (Cp <= 32 ? launch_spmm_hess<G_, 1>(n_rows, rowptr, col, val, stats, ld_stats, Cp, c0, ncols, width, y, ldy, st) \
                              : launch_spmm_hess<G_, 2>(n_rows, rowptr, col, val, stats, ld_stats, Cp, c0, ncols, width, y, ldy, st))
  switch (g4) {
    case 1: return LGNN_SH(4);
    case 2: return LGNN_SH(8);
    case 3: return LGNN_SH(12);
    default: return LGNN_SH(16);
  }
#undef LGNN_SH
}
