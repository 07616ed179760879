// Multi-RHS SpMM over UNIT-COMPACTED slabs.
//
// Below the output layer the right-hand sides of the KFAC backward are
//     delta[n, c, u] = (gZ W)[n, c, u] * 1[H[n, u] > 0]        (c: Hessian-sqrt column, u: hidden unit)
// — the relu' mask depends on (node, unit) only, so the g columns of a node share ONE zero pattern.
// The SpMM gZ' = Â^T delta that follows is a pure gather of those rows and moves 12 of the 13.5 TB
// of a products-shaped fit, so the rows are stored with their dead units squeezed out:
//
//   slab row n (in place, same pitch as the dense row [g][h]):
//       [slot S = 0 .. k_n-1][c = 0 .. g-1]      S-th ACTIVE unit of node n, its g values contiguous
//   header  hdr[n][w] = { mask: bit L <-> unit 32w+L active,  first: slot of the block's first active unit }
//
// lgnn_unit_pack_f32 rewrites a dense masked slab into that form in place (one CTA per node: read the
// whole row, barrier, write the compact row).  lgnn_spmm_units_f32 gives one warp the output tile
// (row i, units 32w .. 32w+31, all g columns): lane L owns unit 32w+L with its g accumulators in
// registers.  Per neighbour j the warp reads hdr[j][w] (one 8-byte word, fetched for 32 neighbours
// at once like col / val); the piece of row j the warp needs is ONE contiguous run of popc(mask)*g*4
// bytes — no per-element decode — and the eight warps of a CTA walk adjacent runs of the same row.
// Dead units contribute nothing, exactly the terms the dense kernel multiplies by zero, and the
// neighbour order is the dense kernel's — the result is bit-identical to lgnn_spmm_f32 on the
// uncompacted slab.  Two kernels: the simple one (lanes load their own slot straight from global
// memory) and the staged one (the run is copied to shared memory as a whole), see below.
//
// Algorithmic bytes per launch: nnz*(4+4) + (n_rows+1)*8 + nnz*nblk*8 [headers]
//                               + sum over edges of k_col*g*4 [live values] + n_rows*g*h*4 [dense output].
#include "common.cuh"
#include "spmm_internal.cuh"

namespace lgnn {

constexpr int UNITS_THREADS = 256;

// One CTA per node, one thread per hidden unit.  In place (src == dst, same pitch) the barrier between the loads
// and the stores is what makes the rewrite safe.  Ragged (row_first given): node n's slots start at the ABSOLUTE
// slot row_first[n] of dst — rows lie back to back, which is what travels between ranks in the row-partitioned
// backward — and the header words carry absolute slots (the SpMM is then launched with pitch 0).  src == nullptr
// writes the headers only, hdr == nullptr the values only.
template <int G4>
__global__ void __launch_bounds__(1024) unit_pack_kernel(const UnitPackArgs A) {
  __shared__ uint32_t cnt[32];
  const int u = threadIdx.x, lane = u & 31, w = u >> 5;
  const int64_t row = blockIdx.x;
  const int64_t r0 = A.row_first ? __ldg(A.row_first + row) : 0;
  const bool on = __ldg(A.act + row * A.lda + u) > 0.f;
  const uint32_t m = __ballot_sync(0xffffffffu, on);
  if (lane == 0) cnt[w] = __popc(m);
  float v[4 * G4];
  if (A.src) {
    const float* base = A.src + row * A.lds;
#pragma unroll
    for (int c = 0; c < 4 * G4; ++c) v[c] = base[(int64_t)c * A.h + u];
  }
  __syncthreads();                      // every dense value is in registers before the row is overwritten
  uint32_t first = 0;
  for (int b = 0; b < w; ++b) first += cnt[b];
  if (lane == 0 && A.hdr) A.hdr[row * (A.h >> 5) + w] = make_uint2(m, (uint32_t)r0 + first);
  if (on && A.src) {
    float* out = A.row_first ? A.dst + r0 * (4 * G4) : A.dst + row * A.ldd;
    float4* dst = reinterpret_cast<float4*>(out + (int64_t)(first + __popc(m & ((1u << lane) - 1u))) * (4 * G4));
#pragma unroll
    for (int t = 0; t < G4; ++t) dst[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
  }
}

// 128-bit read-only load under a lane predicate; lanes that do not load get zeros
__device__ __forceinline__ float4 ldg_f4_if(const float* p, uint32_t on) {
  float4 r;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.u32 q, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
      : "l"(p), "r"(on));
  return r;
}

// G4: float4 per slot (g = 4*G4 columns); U: neighbours in flight per warp.
template <int G4, int U, int MINB>
__global__ void __launch_bounds__(UNITS_THREADS, MINB) spmm_units_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ slab, int64_t lds, const uint2* __restrict__ hdr,
    int nblk, float* __restrict__ y, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * UNITS_THREADS + threadIdx.x) >> 5;
  const int64_t row = warp / nblk;
  if (row >= n_rows) return;
  const int w = (int)(warp - row * nblk);
  const uint32_t lt = (1u << lane) - 1u;

  float4 acc[G4];
#pragma unroll
  for (int t = 0; t < G4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int64_t beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  for (int64_t k0 = beg; k0 < end; k0 += 32) {
    const int cnt = (int)((end - k0) < 32 ? (end - k0) : 32);
    float my_v = 0.f;
    uint32_t my_m = 0u;
    const float* my_p = slab;               // first live value of this warp's unit block in row col[k]
    if (lane < cnt) {
      const int32_t c = __ldg(col + k0 + lane);
      my_v = __ldg(val + k0 + lane);
      const uint2 hd = __ldg(hdr + (int64_t)c * nblk + w);
      my_m = hd.x;
      my_p = slab + (int64_t)c * lds + (int64_t)hd.y * (4 * G4);
    }
    // lanes >= cnt hold an empty mask: the tail of a batch issues no loads
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float4 buf[U][G4];
      float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int srcl = (j + u) & 31;
        vv[u] = __shfl_sync(0xffffffffu, my_v, srcl);
        uint32_t m = __shfl_sync(0xffffffffu, my_m, srcl);
        const uint64_t p = __shfl_sync(0xffffffffu, (uint64_t)(uintptr_t)my_p, srcl);
        if (j + u >= 32) m = 0u;
        const uint32_t on = (m >> lane) & 1u;
        const float* src = reinterpret_cast<const float*>((uintptr_t)p) + __popc(m & lt) * (4 * G4);
#pragma unroll
        for (int t = 0; t < G4; ++t) buf[u][t] = ldg_f4_if(src + 4 * t, on);
      }
      // a dead unit holds zeros: the FMA is the dense kernel's multiplication by zero
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int t = 0; t < G4; ++t) fma4(acc[t], vv[u], buf[u][t]);
    }
  }
  float* yb = y + row * ldy + 32 * w + lane;
  const int64_t h = (int64_t)nblk * 32;
#pragma unroll
  for (int t = 0; t < G4; ++t) {
    yb[(4 * t + 0) * h] = acc[t].x;
    yb[(4 * t + 1) * h] = acc[t].y;
    yb[(4 * t + 2) * h] = acc[t].z;
    yb[(4 * t + 3) * h] = acc[t].w;
  }
}

// ------------------------------------------------------------------------------------------
// Staged variant (the default).  The simple kernel above issues, per neighbour, G4 128-bit loads whose
// lanes sit 16*G4 bytes apart: every instruction touches all ~7 cache lines of the run, and the L1
// wavefront pipe — not HBM — saturates (measured: 0.79 of the copy rate).  Here the warp copies the run
// as what it is, ONE contiguous piece: cp.async (LDGSTS, 16 bytes per lane, consecutive lanes ->
// consecutive addresses, L2 -> shared memory without passing through registers) into a per-warp ring
// of U slots, U-1 neighbours in flight; the owner lanes then read their g values back with
// conflict-free 128-bit shared-memory loads (slot stride 48 bytes at g = 12; an XOR swizzle of the
// 16-byte column for g = 8 / 16).
//
// A group of 8 warps (one per 32-unit block when h = 256) owns `rpg` CONSECUTIVE rows, whose non-zeros
// are one contiguous run of (col, val); it is walked in batches of 32 with the (col, val) of batch b+2
// and the header words of batch b+1 in flight behind the gathers of batch b, so the dependent chain
// rowptr -> col -> header -> values is paid once per row group.  A row boundary inside a batch only
// flushes the accumulators; the copy pipeline keeps running across it.
template <int G4>
__device__ __forceinline__ int unit_swz(int s) {
  return G4 == 4 ? ((s >> 1) & 3) : (G4 == 2 ? ((s >> 2) & 1) : 0);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}

template <int G4, int U, int MINB>
__global__ void __launch_bounds__(UNITS_THREADS, MINB) spmm_units_staged_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const float* __restrict__ slab, int64_t lds, const uint2* __restrict__ hdr,
    int nblk, float* __restrict__ y, int64_t ldy, int rpg) {
  extern __shared__ __align__(16) uint8_t units_smem_[];
  constexpr int SLOT = 512 * G4;                       // bytes: 32 live units x g floats
  const int lane = threadIdx.x & 31;
  const uint32_t ring = smem_addr(units_smem_) + (threadIdx.x >> 5) * (U * SLOT);
  const int64_t warp = ((int64_t)blockIdx.x * UNITS_THREADS + threadIdx.x) >> 5;
  const int64_t grp = warp / nblk;
  const int64_t r0 = grp * rpg;
  if (r0 >= n_rows) return;
  const int w = (int)(warp - grp * nblk);
  const int nr = (int)((n_rows - r0) < rpg ? (n_rows - r0) : rpg);
  const uint32_t lt = (1u << lane) - 1u;
  const int h = nblk * 32;

  // non-zeros of the row group, as 32-bit offsets from its first one
  const int64_t my_rp64 = (lane <= nr) ? __ldg(rowptr + r0 + lane) : 0;      // rpg <= 31
  const int64_t k_beg = __shfl_sync(0xffffffffu, my_rp64, 0);
  const int my_rp = (int)(my_rp64 - k_beg);
  const int k_end = __shfl_sync(0xffffffffu, my_rp, nr);
  col += k_beg;
  val += k_beg;
  hdr += w;
  float* yb = y + r0 * ldy + 32 * w + lane;
  const float4* slab4 = reinterpret_cast<const float4*>(slab);

  float4 acc[G4];
#pragma unroll
  for (int t = 0; t < G4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_cv = [&](int k0, int32_t& c, float& v) {
    const int k = k0 + lane;
    c = -1;
    v = 0.f;
    if (k < k_end) {
      c = __ldg(col + k);
      v = __ldg(val + k);
    }
  };
  // slab position of a neighbour's first live value, in float4 units (the host checked that it fits 32 bits)
  auto load_hdr = [&](int32_t c, uint32_t& m, uint32_t& p) {
    m = 0u;
    p = 0u;
    if (c >= 0) {
      const uint2 hd = __ldg(hdr + (int64_t)c * nblk);
      m = hd.x;
      p = (uint32_t)(((int64_t)c * lds) >> 2) + hd.y * G4;
    }
  };
  int row = 0;
  auto flush = [&]() {
#pragma unroll
    for (int t = 0; t < G4; ++t) {
      yb[(4 * t + 0) * h] = acc[t].x;
      yb[(4 * t + 1) * h] = acc[t].y;
      yb[(4 * t + 2) * h] = acc[t].z;
      yb[(4 * t + 3) * h] = acc[t].w;
      acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    yb += ldy;
    ++row;
  };

  int32_t c_a;                 // batch b+1: (col, val) loaded, header not yet
  float v_a;
  float my_v;                  // batch b: everything loaded
  uint32_t my_m, my_p;
  load_cv(0, c_a, v_a);
  my_v = v_a;
  load_hdr(c_a, my_m, my_p);
  load_cv(32, c_a, v_a);
  int row_end = __shfl_sync(0xffffffffu, my_rp, 1);

  for (int k0 = 0; k0 < k_end; k0 += 32) {
    uint32_t m_n, p_n;
    const float v_n = v_a;
    load_hdr(c_a, m_n, p_n);
    load_cv(k0 + 64, c_a, v_a);
    const int cnt = (k_end - k0) < 32 ? (k_end - k0) : 32;

    uint32_t slot_i = ring, slot_c = ring;             // ring slot of the next copy / of the next neighbour read
    auto issue = [&](int jj) {
      if (jj < cnt) {
        const uint32_t m = __shfl_sync(0xffffffffu, my_m, jj);
        const uint32_t p = __shfl_sync(0xffffffffu, my_p, jj);
        const int n4 = __popc(m) * G4;                 // float4s in the run
#pragma unroll
        for (int t = 0; t < G4; ++t) {
          const int idx = lane + 32 * t;
          if (idx < n4) {
            int d = idx;
            if (G4 == 4) d = (idx & ~3) | ((idx & 3) ^ unit_swz<G4>(idx >> 2));
            if (G4 == 2) d = (idx & ~1) | ((idx & 1) ^ unit_swz<G4>(idx >> 1));
            cp_async16(slot_i + 16u * d, slab4 + (p + idx));
          }
        }
      }
      cp_async_commit();
      slot_i = (slot_i + SLOT == ring + U * SLOT) ? ring : slot_i + SLOT;
    };
#pragma unroll
    for (int jj = 0; jj < U - 1; ++jj) issue(jj);
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      issue(j + U - 1);
      cp_async_wait<U - 1>();
      __syncwarp();
      while (k0 + j == row_end) {           // rows ending here (empty rows flush zeros)
        flush();
        row_end = __shfl_sync(0xffffffffu, my_rp, row + 1);
      }
      const uint32_t m = __shfl_sync(0xffffffffu, my_m, j);
      const float v = __shfl_sync(0xffffffffu, my_v, j);
      if ((m >> lane) & 1u) {
        const int sl = __popc(m & lt);
        const int sw = unit_swz<G4>(sl);
#pragma unroll
        for (int t = 0; t < G4; ++t) fma4(acc[t], v, lds_f4(slot_c + 16u * (sl * G4 + (t ^ sw))));
      }
      __syncwarp();                         // the slot is rewritten by the copy issued next iteration
      slot_c = (slot_c + SLOT == ring + U * SLOT) ? ring : slot_c + SLOT;
    }
    my_v = v_n;
    my_m = m_n;
    my_p = p_n;
  }
  while (row < nr) flush();
}

template <int G4, int U, int MINB>
static int launch_units_staged(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                               const float* slab, int64_t lds, const uint2* hdr, int nblk, float* y, int64_t ldy,
                               int rpg, cudaStream_t st) {
  const int64_t warps = (n_rows + rpg - 1) / rpg * nblk;
  const int64_t blocks = (warps + UNITS_THREADS / 32 - 1) / (UNITS_THREADS / 32);
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_units: grid too large");
  const int smem = (UNITS_THREADS / 32) * U * 512 * G4;
  LGNN_CUDA_TRY(cudaFuncSetAttribute(spmm_units_staged_kernel<G4, U, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  spmm_units_staged_kernel<G4, U, MINB><<<(unsigned)blocks, UNITS_THREADS, smem, st>>>(n_rows, rowptr, col, val, slab,
                                                                                     lds, hdr, nblk, y, ldy, rpg);
  LGNN_LAUNCH_CHECK("spmm_units_staged_kernel");
  return LGNN_OK;
}

template <int G4, int U, int MINB>
static int launch_units(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                        const float* slab, int64_t lds, const uint2* hdr, int nblk, float* y, int64_t ldy,
                        cudaStream_t st) {
  const int64_t warps = n_rows * nblk;
  const int64_t blocks = (warps + UNITS_THREADS / 32 - 1) / (UNITS_THREADS / 32);
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_units: grid too large");
  spmm_units_kernel<G4, U, MINB><<<(unsigned)blocks, UNITS_THREADS, 0, st>>>(n_rows, rowptr, col, val, slab, lds,
                                                                             hdr, nblk, y, ldy);
  LGNN_LAUNCH_CHECK("spmm_units_kernel");
  return LGNN_OK;
}

// g % 4 == 0: the kernels of this file; g % 4 == 2: spmm_units_even.cu (block runs start on even slots)
static bool units_shape_ok(int64_t g, int64_t h) {
  return g >= 2 && g <= 16 && g % 2 == 0 && h >= 32 && h <= 1024 && h % 32 == 0;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" int lgnn_unit_slabs_supported(int64_t g, int64_t h) { return units_shape_ok(g, h) ? 1 : 0; }

static int unit_pack_launch(const UnitPackArgs& A, int64_t n_rows, int64_t g, cudaStream_t st) {
  if (g % 4 != 0) return unit_pack_even(A, n_rows, (int)g, st);
  const unsigned grid = (unsigned)n_rows, block = (unsigned)A.h;
  switch (g / 4) {
    case 1: unit_pack_kernel<1><<<grid, block, 0, st>>>(A); break;
    case 2: unit_pack_kernel<2><<<grid, block, 0, st>>>(A); break;
    case 3: unit_pack_kernel<3><<<grid, block, 0, st>>>(A); break;
    default: unit_pack_kernel<4><<<grid, block, 0, st>>>(A); break;
  }
  LGNN_LAUNCH_CHECK("unit_pack_kernel");
  return LGNN_OK;
}

extern "C" int lgnn_unit_pack_f32(float* slab, int64_t lds, const float* act, int64_t lda, int64_t n_rows,
                                  int64_t g, int64_t h, void* hdr, lgnn_stream_t stream) {
  if (n_rows < 0) return fail(LGNN_E_BADARG, "unit_pack: negative row count");
  if (!units_shape_ok(g, h)) return fail(LGNN_E_UNSUPPORTED, "unit_pack: g must be even, 2 .. 16, and h a multiple of 32 up to 1024 (g=%lld h=%lld)", (long long)g, (long long)h);
  if (n_rows == 0) return LGNN_OK;
  if (!slab || !act || !hdr) return fail(LGNN_E_BADARG, "unit_pack: null pointer");
  if (lds < g * h || lda < h) return fail(LGNN_E_BADARG, "unit_pack: pitch smaller than the row");
  if ((reinterpret_cast<uintptr_t>(slab) & 15) || (lds % 4) || (reinterpret_cast<uintptr_t>(hdr) & 7))
    return fail(LGNN_E_ALIGN, "unit_pack: slab must be 16-byte aligned with a pitch that is a multiple of 4 floats, hdr 8-byte aligned");
  if (n_rows > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "unit_pack: too many rows");
  UnitPackArgs A{slab, lds, slab, lds, nullptr, act, lda, (int)h, reinterpret_cast<uint2*>(hdr)};
  return unit_pack_launch(A, n_rows, g, as_stream(stream));
}

extern "C" int lgnn_unit_pack_ragged_f32(const float* src, int64_t lds, const float* act, int64_t lda,
                                         int64_t n_rows, int64_t g, int64_t h, const int64_t* row_first,
                                         float* dst, void* hdr, lgnn_stream_t stream) {
  if (n_rows < 0) return fail(LGNN_E_BADARG, "unit_pack_ragged: negative row count");
  if (!units_shape_ok(g, h)) return fail(LGNN_E_UNSUPPORTED, "unit_pack_ragged: g must be even, 2 .. 16, and h a multiple of 32 up to 1024 (g=%lld h=%lld)", (long long)g, (long long)h);
  if (n_rows == 0) return LGNN_OK;
  if (!act || !row_first) return fail(LGNN_E_BADARG, "unit_pack_ragged: null act / row_first");
  if (!src && !hdr) return fail(LGNN_E_BADARG, "unit_pack_ragged: neither values (src) nor headers (hdr) asked for");
  if (src && !dst) return fail(LGNN_E_BADARG, "unit_pack_ragged: src without dst");
  if (lda < h || (src && lds < g * h)) return fail(LGNN_E_BADARG, "unit_pack_ragged: pitch smaller than the row");
  if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(hdr) & 7) || (reinterpret_cast<uintptr_t>(row_first) & 7))
    return fail(LGNN_E_ALIGN, "unit_pack_ragged: dst must be 16-byte aligned, hdr and row_first 8-byte aligned");
  if (n_rows > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "unit_pack_ragged: too many rows");
  UnitPackArgs A{src, lds, dst, 0, row_first, act, lda, (int)h, reinterpret_cast<uint2*>(hdr)};
  return unit_pack_launch(A, n_rows, g, as_stream(stream));
}

// slab_floats: extent of the slab (decides whether the 32-bit float4 offsets of the pipelined kernels reach it)
static int spmm_units_dispatch(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                               const float* slab, int64_t lds, int64_t slab_floats, const void* hdr, int64_t g,
                               int64_t h, float* y, int64_t ldy, int flags, lgnn_stream_t stream);

extern "C" int lgnn_spmm_units_f32(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* rowptr, const int32_t* col,
                                   const float* val, const float* slab, int64_t lds, const void* hdr,
                                   int64_t g, int64_t h, float* y, int64_t ldy, int flags,
                                   lgnn_stream_t stream) {
  if (n_rows < 0 || n_cols < 0 || nnz < 0) return fail(LGNN_E_BADARG, "spmm_units: negative size");
  if (!units_shape_ok(g, h)) return fail(LGNN_E_UNSUPPORTED, "spmm_units: g must be even, 2 .. 16, and h a multiple of 32 up to 1024 (g=%lld h=%lld)", (long long)g, (long long)h);
  if (n_rows == 0) return LGNN_OK;
  if (!rowptr || !slab || !hdr || !y) return fail(LGNN_E_BADARG, "spmm_units: null pointer");
  if (nnz > 0 && (!col || !val)) return fail(LGNN_E_BADARG, "spmm_units: null col / val");
  if (lds < g * h || ldy < g * h) return fail(LGNN_E_BADARG, "spmm_units: pitch smaller than the row");
  if ((reinterpret_cast<uintptr_t>(slab) & 15) || (lds % 4) || (reinterpret_cast<uintptr_t>(hdr) & 7))
    return fail(LGNN_E_ALIGN, "spmm_units: slab must be 16-byte aligned with a pitch that is a multiple of 4 floats, hdr 8-byte aligned");
  return spmm_units_dispatch(n_rows, rowptr, col, val, slab, lds, n_cols * lds, hdr, g, h, y, ldy, flags, stream);
}

// Ragged rows (lgnn_unit_pack_ragged_f32): the headers hold absolute slots, so a row is found through its header
// alone — the kernels are the same, run with pitch 0.
extern "C" int lgnn_spmm_units_ragged_f32(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* rowptr,
                                          const int32_t* col, const float* val, const float* slab,
                                          int64_t slab_floats, const void* hdr, int64_t g, int64_t h, float* y,
                                          int64_t ldy, int flags, lgnn_stream_t stream) {
  if (n_rows < 0 || n_cols < 0 || nnz < 0 || slab_floats < 0) return fail(LGNN_E_BADARG, "spmm_units_ragged: negative size");
  if (!units_shape_ok(g, h)) return fail(LGNN_E_UNSUPPORTED, "spmm_units_ragged: g must be even, 2 .. 16, and h a multiple of 32 up to 1024 (g=%lld h=%lld)", (long long)g, (long long)h);
  if (n_rows == 0) return LGNN_OK;
  if (!rowptr || !slab || !hdr || !y) return fail(LGNN_E_BADARG, "spmm_units_ragged: null pointer");
  if (nnz > 0 && (!col || !val)) return fail(LGNN_E_BADARG, "spmm_units_ragged: null col / val");
  if (ldy < g * h) return fail(LGNN_E_BADARG, "spmm_units_ragged: pitch smaller than the row");
  if ((reinterpret_cast<uintptr_t>(slab) & 15) || (reinterpret_cast<uintptr_t>(hdr) & 7))
    return fail(LGNN_E_ALIGN, "spmm_units_ragged: slab must be 16-byte aligned, hdr 8-byte aligned");
  if (slab_floats / g > 0xffffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_units_ragged: more than 2^32 slots");
  return spmm_units_dispatch(n_rows, rowptr, col, val, slab, 0, slab_floats, hdr, g, h, y, ldy, flags, stream);
}

static int spmm_units_dispatch(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                               const float* slab, int64_t lds, int64_t slab_floats, const void* hdr, int64_t g,
                               int64_t h, float* y, int64_t ldy, int flags, lgnn_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  const uint2* hd = reinterpret_cast<const uint2*>(hdr);
  const int nblk = (int)(h / 32);
  int variant = (flags >> 8) & 0xff;         // lab / test override of the unroll and occupancy choice
  // defaults measured on the products shape (profiles/r2a_units_lab.txt): 8 rows per warp group everywhere; g = 16:
  // ring of 4, 3 CTAs per SM (182.7 against 187.8 ms); g = 12: ring of 4, 4 CTAs per SM (139.1 against 141.7 ms)
  if (variant == 0 && g % 4 == 0) variant = g == 16 ? 14 : (g == 12 ? 13 : 12);
  // the pipelined kernel addresses the slab with 32-bit float4 offsets: slabs up to 64 GB
  const bool narrow = slab_floats <= ((int64_t)1 << 34);
  if (g % 4 != 0) {
    if (!narrow) return fail(LGNN_E_UNSUPPORTED, "spmm_units: g = %lld needs a slab of at most 64 GB", (long long)g);
    return spmm_units_even(n_rows, rowptr, col, val, slab, lds, hd, (int)g, nblk, y, ldy, variant, st);
  }
  if ((variant == 0 || variant >= 8) && narrow) {  // default: the staged kernel, 4 rows per group
    const int rpg = 4 << ((variant >> 2) & 1);     // 8..11: 4 rows per group, 12..15: 8 rows
    const int cfg = variant & 3;
#define LGNN_STAGED(G4_, U_, MINB_) launch_units_staged<G4_, U_, MINB_>(n_rows, rowptr, col, val, slab, lds, hd, nblk, y, ldy, rpg, st)
    switch (g / 4) {
      case 1: return LGNN_STAGED(1, 8, 3);
      case 2: return LGNN_STAGED(2, 6, 3);
      case 3:
        switch (cfg) {
          case 1: return LGNN_STAGED(3, 4, 4);
          case 2: return LGNN_STAGED(3, 6, 3);
          case 3: return LGNN_STAGED(3, 8, 2);
          default: return LGNN_STAGED(3, 4, 3);
        }
      default:
        switch (cfg) {
          case 1: return LGNN_STAGED(4, 3, 4);
          case 2: return LGNN_STAGED(4, 4, 3);
          default: return LGNN_STAGED(4, 6, 2);
        }
    }
#undef LGNN_STAGED
  }
  // variants 1..7, or a slab beyond 64 GB: the simple kernel (64-bit addressing)
  switch (g / 4) {
    case 1: return launch_units<1, 4, 3>(n_rows, rowptr, col, val, slab, lds, hd, nblk, y, ldy, st);
    case 2: return launch_units<2, 4, 3>(n_rows, rowptr, col, val, slab, lds, hd, nblk, y, ldy, st);
    case 3: return launch_units<3, 4, 3>(n_rows, rowptr, col, val, slab, lds, hd, nblk, y, ldy, st);
    default: return launch_units<4, 4, 2>(n_rows, rowptr, col, val, slab, lds, hd, nblk, y, ldy, st);
  }
}
