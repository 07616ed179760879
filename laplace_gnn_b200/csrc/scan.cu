// Exclusive prefix sum int32 counts -> int64 offsets (three-phase reduce / scan / downsweep).
// Integer, HBM-bound, deterministic.  Used by the CSR build, transpose and halo kernels.
#include "common.cuh"

namespace lgnn {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total via *total
__device__ __forceinline__ int64_t block_excl_scan(int64_t v, int64_t* total) {
  __shared__ int64_t warp_tot[SCAN_THREADS / 32];
  __shared__ int64_t block_tot;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t incl = warp_incl_scan(v);
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    int64_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
    int64_t ti = warp_incl_scan(t);
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
    if (lane == SCAN_THREADS / 32 - 1) block_tot = ti;
  }
  __syncthreads();
  int64_t out = incl - v + warp_tot[w];
  *total = block_tot;
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int32_t* __restrict__ in,
                                                                   int64_t n,
                                                                   int64_t* __restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t j = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (j < n) s += in[j];
  }
  int64_t tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(int64_t* __restrict__ tile_sums,
                                                                  int64_t n_tiles) {
  // single block: serial over chunks of SCAN_THREADS tiles
  int64_t carry = 0;
  for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    int64_t j = base + threadIdx.x;
    int64_t v = j < n_tiles ? tile_sums[j] : 0;
    int64_t tot;
    int64_t ex = block_excl_scan(v, &tot);
    if (j < n_tiles) tile_sums[j] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const int32_t* __restrict__ in,
                                                                 int64_t n,
                                                                 const int64_t* __restrict__ tile_off,
                                                                 int64_t* __restrict__ out) {
  // thread t owns SCAN_ITEMS consecutive elements (blocked arrangement)
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t j = base + i;
    v[i] = j < n ? in[j] : 0;
    s += v[i];
  }
  int64_t tot;
  int64_t ex = block_excl_scan(s, &tot) + tile_off[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t j = base + i;
    if (j < n) out[j] = ex;
    ex += v[i];
    if (j == n - 1) out[n] = ex;
  }
}

__global__ void scan_empty_kernel(int64_t* out) { out[0] = 0; }

size_t scan_workspace_bytes(int64_t n) {
  int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  return align_up((size_t)(tiles > 0 ? tiles : 1) * sizeof(int64_t), 256);
}

int exclusive_scan_i32_to_i64(const int32_t* in, int64_t n, int64_t* out, void* ws, size_t ws_bytes,
                              cudaStream_t st) {
  if (n < 0) return fail(LGNN_E_BADARG, "scan: n < 0");
  if (n == 0) {
    scan_empty_kernel<<<1, 1, 0, st>>>(out);
    LGNN_LAUNCH_CHECK("scan_empty_kernel");
    return LGNN_OK;
  }
  if (ws_bytes < scan_workspace_bytes(n)) return fail(LGNN_E_NOMEM, "scan: workspace too small");
  int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int64_t* tile_sums = reinterpret_cast<int64_t*>(ws);
  scan_reduce_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, tile_sums);
  LGNN_LAUNCH_CHECK("scan_reduce_kernel");
  scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(tile_sums, tiles);
  LGNN_LAUNCH_CHECK("scan_tiles_kernel");
  scan_down_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, tile_sums, out);
  LGNN_LAUNCH_CHECK("scan_down_kernel");
  return LGNN_OK;
}

}  // namespace lgnn
