// Multi-RHS SpMM over ZERO-COMPRESSED slabs.
//
// The right-hand sides of the KFAC backward below the output layer are delta = (gZ W) ⊙ 1[H > 0]: the
// relu' mask leaves about half of every row exactly zero.  The SpMM that follows is a pure gather of
// those rows (13 TB of HBM reads per fit on the products shape), so the rows are stored compressed
// and the gather moves only their non-zeros:
//
//   packed row (fixed pitch, only the first len[r] bytes are ever read):
//     masks  [nblk][4] uint32   block b covers elements 128b .. 128b+127; bit L of word j <-> element 128b + 4L + j
//     prefix [nblk]    uint32   number of non-zeros before block b          (padded to 16 bytes)
//     values [nnz]     float    the non-zeros in natural order              (padded to 16 bytes)
//
// lgnn_pack_rows_f32 builds it (one warp per row, ballots + popcounts, coalesced 128-bit reads);
// lgnn_spmm_packed_f32 is the bulk-ring SpMM of spmm.cu with two changes: the producer copies
// len[col] bytes per neighbour (one cp.async.bulk, TMA 1-D), and a consumer thread finds the packed
// position of its float4 column with 4 popcounts (the mask block is warp-uniform: a broadcast load),
// reads its <= 4 values and FMAs exactly as the dense kernel does — same summation order, so the
// result is bit-identical to the dense SpMM of the uncompressed slab.
//
// Algorithmic bytes per launch: nnz*(4+4+4) [col, val, len gather] + (n_rows+1)*8 + sum over edges
// of len[col] + n_rows*d*4 [dense output].
#include "common.cuh"
#include "spmm_internal.cuh"

namespace lgnn {

constexpr int PK_BLOCK = 128;   // elements per mask block

struct PackGeom {
  int nblk;            // mask blocks per row
  int header_bytes;    // masks + prefix, multiple of 16
  int64_t pitch;       // bytes per packed row slot = header + d_pad * 4
};

__host__ __device__ inline PackGeom pack_geom(int64_t d) {
  PackGeom g;
  g.nblk = (int)((d + PK_BLOCK - 1) / PK_BLOCK);
  g.header_bytes = g.nblk * 16 + (g.nblk * 4 + 15) / 16 * 16;
  g.pitch = (int64_t)g.header_bytes + (int64_t)g.nblk * PK_BLOCK * 4;
  return g;
}

// one warp per row
__global__ void __launch_bounds__(256) pack_rows_kernel(int64_t n_rows, const float* __restrict__ x, int64_t ldx,
                                                        int d, uint8_t* __restrict__ packed, int64_t pitch,
                                                        int nblk, int header_bytes, int32_t* __restrict__ len) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const float* src = x + row * ldx;
  uint8_t* dst = packed + row * pitch;
  uint4* masks = reinterpret_cast<uint4*>(dst);
  uint32_t* prefix = reinterpret_cast<uint32_t*>(dst + (size_t)nblk * 16);
  float* values = reinterpret_cast<float*>(dst + header_bytes);
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t running = 0;
  for (int b = 0; b < nblk; ++b) {
    const int e0 = b * PK_BLOCK + 4 * lane;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e0 + 3 < d) {
      v = __ldg(reinterpret_cast<const float4*>(src + e0));
    } else {
      if (e0 + 0 < d) v.x = __ldg(src + e0 + 0);
      if (e0 + 1 < d) v.y = __ldg(src + e0 + 1);
      if (e0 + 2 < d) v.z = __ldg(src + e0 + 2);
    }
    const uint32_t w0 = __ballot_sync(0xffffffffu, v.x != 0.f);
    const uint32_t w1 = __ballot_sync(0xffffffffu, v.y != 0.f);
    const uint32_t w2 = __ballot_sync(0xffffffffu, v.z != 0.f);
    const uint32_t w3 = __ballot_sync(0xffffffffu, v.w != 0.f);
    if (lane == 0) {
      masks[b] = make_uint4(w0, w1, w2, w3);
      prefix[b] = running;
    }
    uint32_t off = running + __popc(w0 & lt) + __popc(w1 & lt) + __popc(w2 & lt) + __popc(w3 & lt);
    if (v.x != 0.f) values[off++] = v.x;
    if (v.y != 0.f) values[off++] = v.y;
    if (v.z != 0.f) values[off++] = v.z;
    if (v.w != 0.f) values[off++] = v.w;
    running += __popc(w0) + __popc(w1) + __popc(w2) + __popc(w3);
  }
  if (lane == 0) len[row] = header_bytes + (int32_t)((running * 4u + 15u) / 16u * 16u);
}

template <int VPT>  // float4 accumulators per consumer thread; d <= VPT*256 float4
__global__ void __launch_bounds__(BULK_THREADS, 1) spmm_packed_kernel(
    int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, const uint8_t* __restrict__ packed, int64_t pitch,
    const int32_t* __restrict__ len, int nblk, int header_bytes, float* __restrict__ y, int64_t ldy, int d4,
    int stages, int stage_bytes, int64_t nnz_per_cta, int64_t hub_len) {
  extern __shared__ uint8_t pk_smem_[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pk_smem_) + 127) & ~(uintptr_t)127);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)stages * stage_bytes);
  uint64_t* empty = full + BULK_MAX_STAGES;
  float* meta = reinterpret_cast<float*>(empty + BULK_MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nnz = __ldg(rowptr + n_rows);
  const bool last_cta = blockIdx.x == gridDim.x - 1;
  const int64_t t_beg = (int64_t)blockIdx.x * nnz_per_cta;
  if (blockIdx.x > 0 && t_beg >= nnz) return;
  const int64_t t_end = (last_cta || t_beg + nnz_per_cta >= nnz) ? nnz : t_beg + nnz_per_cta;
  const int64_t row_beg = blockIdx.x == 0 ? 0 : row_lower_bound(rowptr, n_rows, t_beg);
  const int64_t row_end = (last_cta || t_end >= nnz) ? n_rows : row_lower_bound(rowptr, n_rows, t_end);
  // hub rows are split at the budget boundaries exactly as in spmm_bulk_kernel
  const int64_t main_beg = __ldg(rowptr + row_beg);
  const bool has_prefix = blockIdx.x > 0 && row_beg > 0 && main_beg > t_beg &&
                          (main_beg - __ldg(rowptr + row_beg - 1)) > hub_len;
  const int64_t pre_end = main_beg < t_end ? main_beg : t_end;
  int64_t s_end = has_prefix ? pre_end : main_beg;
  if (row_end > row_beg) {
    s_end = __ldg(rowptr + row_end);
    const int64_t last_beg = __ldg(rowptr + row_end - 1);
    if (s_end - last_beg > hub_len && s_end > t_end) s_end = t_end;
  }
  const int64_t s_beg = has_prefix ? t_beg : main_beg;
  if (row_beg >= row_end && !has_prefix) return;

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      bulk_mbar_init(smem_addr(&full[i]), 1);
      bulk_mbar_init(smem_addr(&empty[i]), BULK_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == BULK_CONSUMERS / 32) {
    // ===================================================================== producer warp
    for (int64_t k0 = s_beg; k0 < s_end; k0 += 32) {
      const int64_t k = k0 + lane;
      const bool valid = k < s_end;
      int32_t c = 0;
      float v = 0.f;
      uint32_t bytes = 0;
      if (valid) {
        c = __ldg(col + k);
        v = __ldg(val + k);
        bytes = (uint32_t)__ldg(len + c);       // only the live part of the packed row travels
      }
      const int64_t it = k - s_beg;
      const int st = (int)(it % stages);
      const uint32_t ph = (uint32_t)((it / stages) & 1);
      for (int sub = 0; sub * stages < 32; ++sub) {
        if (valid && lane / stages == sub) {
          bulk_mbar_wait(smem_addr(&empty[st]), ph ^ 1u);
          meta[st] = v;
          const uint32_t bar = smem_addr(&full[st]);
          bulk_mbar_expect_tx(bar, bytes);
          bulk_g2s(smem_addr(ring + (size_t)st * stage_bytes), packed + (int64_t)c * pitch, bytes, bar);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== consumer warps
    const int t = threadIdx.x;
    bool act[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) act[j] = (t + j * BULK_CONSUMERS) < d4;
    const uint32_t lt = (1u << lane) - 1u;
    int64_t it = 0;
    auto segment = [&](int64_t row, int64_t cnt, bool atomic) {
      float4 acc[VPT];
#pragma unroll
      for (int j = 0; j < VPT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t e = 0; e < cnt; ++e, ++it) {
        const int st = (int)(it % stages);
        const uint32_t ph = (uint32_t)((it / stages) & 1);
        bulk_mbar_wait(smem_addr(&full[st]), ph);
        const float v = meta[st];
        const uint8_t* base = ring + (size_t)st * stage_bytes;
        const uint4* masks = reinterpret_cast<const uint4*>(base);
        const uint32_t* prefix = reinterpret_cast<const uint32_t*>(base + (size_t)nblk * 16);
        const float* values = reinterpret_cast<const float*>(base + header_bytes);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          if (!act[j]) continue;
          // float4 column q = j*256 + t lives in mask block q / 32 = j*8 + warp (warp-uniform), lane = q % 32
          const int b = j * (BULK_CONSUMERS / 32) + warp;
          const uint4 m = masks[b];
          uint32_t off = prefix[b] + __popc(m.x & lt) + __popc(m.y & lt) + __popc(m.z & lt) + __popc(m.w & lt);
          float4 x;
          const bool bx = (m.x >> lane) & 1u, by = (m.y >> lane) & 1u, bz = (m.z >> lane) & 1u,
                     bw = (m.w >> lane) & 1u;
          x.x = bx ? values[off] : 0.f;
          off += bx;
          x.y = by ? values[off] : 0.f;
          off += by;
          x.z = bz ? values[off] : 0.f;
          off += bz;
          x.w = bw ? values[off] : 0.f;
          fma4(acc[j], v, x);
        }
        __syncwarp();
        if (lane == 0) bulk_mbar_arrive(smem_addr(&empty[st]));
      }
      float4* dst = reinterpret_cast<float4*>(y + row * ldy) + t;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        if (!act[j]) continue;
        const float4 a = acc[j];
        if (atomic) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j * BULK_CONSUMERS), "f"(a.x),
                       "f"(a.y), "f"(a.z), "f"(a.w)
                       : "memory");
        } else {
          dst[j * BULK_CONSUMERS] = a;
        }
      }
    };
    if (has_prefix) segment(row_beg - 1, pre_end - t_beg, true);
    int64_t k_row = main_beg;
    for (int64_t row = row_beg; row < row_end; ++row) {
      const int64_t k_next = __ldg(rowptr + row + 1);
      const bool hub = (k_next - k_row) > hub_len;
      const int64_t k_stop = (hub && k_next > s_end) ? s_end : k_next;
      segment(row, k_stop - k_row, hub);
      k_row = k_next;
    }
  }
}

// hub rows receive their pieces by red.global.add: clear them first
__global__ void pk_zero_hub_rows_kernel(int64_t n_rows, const int64_t* __restrict__ rowptr, int64_t hub_len,
                                        float* __restrict__ y, int64_t ldy, int d4) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  if (__ldg(rowptr + row + 1) - __ldg(rowptr + row) <= hub_len) return;
  float4* dst = reinterpret_cast<float4*>(y + row * ldy);
  for (int c = lane; c < d4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int VPT>
static int launch_packed(int64_t n_rows, const int64_t* rowptr, const int32_t* col, const float* val,
                         const uint8_t* packed, const int32_t* len, const PackGeom& g, float* y, int64_t ldy,
                         int d4, int64_t nnz, int no_hubs, cudaStream_t st) {
  const int stage_bytes = (int)((g.pitch + 127) / 128 * 128);
  int stages = BULK_SMEM_RING / stage_bytes;
  if (stages > BULK_MAX_STAGES) stages = BULK_MAX_STAGES;
  if (stages < 2) return fail(LGNN_E_UNSUPPORTED, "spmm_packed: row too wide for the ring");
  const size_t smem = 128 + (size_t)stages * stage_bytes + 2 * BULK_MAX_STAGES * sizeof(uint64_t) +
                      BULK_MAX_STAGES * sizeof(float);
  LGNN_CUDA_TRY(cudaFuncSetAttribute(spmm_packed_kernel<VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t per_cta = BULK_NNZ_PER_CTA;
  int64_t blocks = (nnz + per_cta - 1) / per_cta;
  if (blocks < 1) blocks = 1;
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_packed: grid too large");
  int64_t hub_len = per_cta * 4;
  if (no_hubs) {
    hub_len = INT64_MAX;
  } else {
    const int64_t zb = (n_rows * 32 + 255) / 256;
    if (zb > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "spmm_packed: grid too large");
    pk_zero_hub_rows_kernel<<<(unsigned)zb, 256, 0, st>>>(n_rows, rowptr, hub_len, y, ldy, d4);
    LGNN_LAUNCH_CHECK("pk_zero_hub_rows_kernel");
  }
  spmm_packed_kernel<VPT><<<(unsigned)blocks, BULK_THREADS, smem, st>>>(
      n_rows, rowptr, col, val, packed, g.pitch, len, g.nblk, g.header_bytes, y, ldy, d4, stages, stage_bytes,
      per_cta, hub_len);
  LGNN_LAUNCH_CHECK("spmm_packed_kernel");
  return LGNN_OK;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" {

int64_t lgnn_pack_rows_pitch(int64_t d) { return d > 0 ? pack_geom(d).pitch : 0; }

int lgnn_pack_rows_f32(const float* x, int64_t ldx, int64_t n_rows, int64_t d, void* packed, int32_t* len,
                       lgnn_stream_t stream) {
  if (n_rows < 0 || d < 1 || d > 4096 || ldx < d) return fail(LGNN_E_BADARG, "pack_rows: bad shape (1 <= d <= 4096)");
  if (n_rows == 0) return LGNN_OK;
  if (!x || !packed || !len) return fail(LGNN_E_BADARG, "pack_rows: null pointer");
  if ((ldx % 4) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
    return fail(LGNN_E_ALIGN, "pack_rows: x and packed must be 16-byte aligned, ldx a multiple of 4");
  const PackGeom g = pack_geom(d);
  const int64_t blocks = (n_rows * 32 + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(LGNN_E_UNSUPPORTED, "pack_rows: grid too large");
  pack_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(n_rows, x, ldx, (int)d, static_cast<uint8_t*>(packed),
                                                                   g.pitch, g.nblk, g.header_bytes, len);
  LGNN_LAUNCH_CHECK("pack_rows_kernel");
  return LGNN_OK;
}

int lgnn_spmm_packed_f32(int64_t n_rows, int64_t nnz, const int64_t* rowptr, const int32_t* col, const float* val,
                         const void* packed, const int32_t* len, int64_t d, float* y, int64_t ldy, int flags,
                         lgnn_stream_t stream) {
  if (n_rows < 0 || nnz < 0 || d < 4 || d > 4096 || (d % 4) || ldy < d)
    return fail(LGNN_E_BADARG, "spmm_packed: bad shape (4 <= d <= 4096, d %% 4 == 0)");
  if (n_rows == 0) return LGNN_OK;
  if (!rowptr || !y || (nnz > 0 && (!col || !val || !packed || !len))) return fail(LGNN_E_BADARG, "spmm_packed: null pointer");
  if ((ldy % 4) || (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(packed) & 15))
    return fail(LGNN_E_ALIGN, "spmm_packed: y and packed must be 16-byte aligned, ldy a multiple of 4");
  const PackGeom g = pack_geom(d);
  const int d4 = (int)(d / 4);
  const int vpt = (d4 + BULK_CONSUMERS - 1) / BULK_CONSUMERS;
  const int no_hubs = (flags & LGNN_SPMM_NO_HUB_ROWS) ? 1 : 0;
  cudaStream_t st = as_stream(stream);
  const uint8_t* p = static_cast<const uint8_t*>(packed);
  switch (vpt) {
    case 1: return launch_packed<1>(n_rows, rowptr, col, val, p, len, g, y, ldy, d4, nnz, no_hubs, st);
    case 2: return launch_packed<2>(n_rows, rowptr, col, val, p, len, g, y, ldy, d4, nnz, no_hubs, st);
    case 3: return launch_packed<3>(n_rows, rowptr, col, val, p, len, g, y, ldy, d4, nnz, no_hubs, st);
    default: return launch_packed<4>(n_rows, rowptr, col, val, p, len, g, y, ldy, d4, nnz, no_hubs, st);
  }
}

}  // extern "C"
