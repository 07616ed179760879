// C = beta*C + alpha * X^T X  — Kronecker-factor accumulation (A_l = H^T H / N, G_l = sum gZ^T gZ).
//
// This file holds the dispatcher, the deterministic split-K reduction shared by both
// implementations, and the CUDA-core fp32 implementation (used for shapes the tcgen05 kernel does
// not take, and as an independent cross-check of it in the tests).  The tensor-core
// implementation lives in syrk_tcgen05.cu.
//
// Split-K layout in the workspace:  part[slice][tile][TS*TS] for the upper-triangular tiles
// (ti <= tj); the reduce kernel sums slices in index order (bit-reproducible run to run), applies
// alpha / beta and mirrors the tile into both triangles of C.
#include "common.cuh"
#include "syrk_internal.cuh"

namespace lgnn {

constexpr int ST = 64;        // SIMT output tile
constexpr int SK = 16;        // k-depth per shared-memory stage
constexpr int SIMT_THREADS = 256;

__global__ void __launch_bounds__(SIMT_THREADS) syrk_simt_kernel(
    const float* __restrict__ x, int64_t ldx, int64_t k_rows, int n, int tiles_1d, int64_t rows_per_slice,
    float* __restrict__ part) {
  __shared__ float As[SK][ST + 4];
  __shared__ float Bs[SK][ST + 4];
  // decode upper-triangular tile index
  int t = blockIdx.x;
  int ti = 0;
  while (t >= tiles_1d - ti) { t -= tiles_1d - ti; ++ti; }
  int tj = ti + t;
  const int slice = blockIdx.y;
  const int64_t k_beg = (int64_t)slice * rows_per_slice;
  int64_t k_end = k_beg + rows_per_slice;
  if (k_end > k_rows) k_end = k_rows;

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lr = threadIdx.x >> 4;        // 0..15 : k within stage
  const int lc = (threadIdx.x & 15) * 4;  // 0..60 : column within tile
  for (int64_t k0 = k_beg; k0 < k_end; k0 += SK) {
    int64_t kr = k0 + lr;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int ca = ti * ST + lc + q, cb = tj * ST + lc + q;
      As[lr][lc + q] = (kr < k_end && ca < n) ? __ldg(x + kr * ldx + ca) : 0.f;
      Bs[lr][lc + q] = (kr < k_end && cb < n) ? __ldg(x + kr * ldx + cb) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n_tiles = tiles_1d * (tiles_1d + 1) / 2;
  float* out = part + ((int64_t)slice * n_tiles + blockIdx.x) * (ST * ST);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(ty * 4 + i) * ST + tx * 4 + j] = acc[i][j];
}

// part[slice][tile][TS*TS] -> C (both triangles), fixed summation order.
__global__ void syrk_reduce_kernel(const float* __restrict__ part, int n_slices, int tiles_1d, int ts,
                                   int n, float alpha, float beta, float* __restrict__ c, int64_t ldc) {
  int t = blockIdx.x;
  int ti = 0;
  while (t >= tiles_1d - ti) { t -= tiles_1d - ti; ++ti; }
  int tj = ti + t;
  const int n_tiles = tiles_1d * (tiles_1d + 1) / 2;
  const int64_t tile_elems = (int64_t)ts * ts;
  for (int e = threadIdx.x; e < tile_elems; e += blockDim.x) {
    int r = e / ts, q = e % ts;
    int gi = ti * ts + r, gj = tj * ts + q;
    if (gi >= n || gj >= n) continue;
    if (ti == tj && q < r) continue;  // lower part of a diagonal tile comes from its mirror
    float s = 0.f;
    for (int sl = 0; sl < n_slices; ++sl) s += part[((int64_t)sl * n_tiles + blockIdx.x) * tile_elems + e];
    float v = alpha * s;
    if (beta != 0.f) {
      float old = c[(int64_t)gi * ldc + gj];
      v = fmaf(beta, old, v);
    }
    c[(int64_t)gi * ldc + gj] = v;
    if (gi != gj) c[(int64_t)gj * ldc + gi] = v;
  }
}

struct SimtPlan {
  int tiles_1d, n_tiles, n_slices;
  int64_t rows_per_slice;
  size_t bytes;
};

static SimtPlan simt_plan(int64_t k_rows, int64_t n) {
  SimtPlan p;
  p.tiles_1d = (int)((n + ST - 1) / ST);
  p.n_tiles = p.tiles_1d * (p.tiles_1d + 1) / 2;
  int64_t want = ((int64_t)sm_count() * 4 + p.n_tiles - 1) / p.n_tiles;  // ~4 blocks per SM
  int64_t max_slices = (k_rows + 4 * SK - 1) / (4 * SK);                  // >= 64 rows per slice
  if (want > max_slices) want = max_slices;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  p.rows_per_slice = ((k_rows + want - 1) / want + SK - 1) / SK * SK;
  if (p.rows_per_slice < SK) p.rows_per_slice = SK;
  p.n_slices = (int)((k_rows + p.rows_per_slice - 1) / p.rows_per_slice);
  if (p.n_slices < 1) p.n_slices = 1;
  p.bytes = (size_t)p.n_slices * p.n_tiles * ST * ST * sizeof(float);
  return p;
}

int syrk_reduce_launch(const float* part, int n_slices, int tiles_1d, int ts, int n, float alpha,
                       float beta, float* c, int64_t ldc, cudaStream_t st) {
  int n_tiles = tiles_1d * (tiles_1d + 1) / 2;
  syrk_reduce_kernel<<<n_tiles, 256, 0, st>>>(part, n_slices, tiles_1d, ts, n, alpha, beta, c, ldc);
  LGNN_LAUNCH_CHECK("syrk_reduce_kernel");
  return LGNN_OK;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" {

size_t lgnn_syrk_workspace_bytes(int64_t k_rows, int64_t n, int impl) {
  if (k_rows < 0 || n <= 0) return 0;
  size_t simt = simt_plan(k_rows, n).bytes;
  size_t tc = 0;
  if (impl != LGNN_SYRK_SIMT && syrk_tcgen05_supported(k_rows, n)) tc = syrk_tcgen05_workspace_bytes(k_rows, n);
  size_t b = simt > tc ? simt : tc;
  return align_up(b, 256);
}

int lgnn_syrk_f32(const float* x, int64_t ldx, int64_t k_rows, int64_t n, float alpha, float beta,
                  float* c, int64_t ldc, void* ws, size_t ws_bytes, int impl, lgnn_stream_t stream) {
  if (!c || n <= 0 || k_rows < 0 || ldx < n || ldc < n) return fail(LGNN_E_BADARG, "syrk: bad argument");
  if (k_rows > 0 && !x) return fail(LGNN_E_BADARG, "syrk: null x");
  if (n > 46340) return fail(LGNN_E_UNSUPPORTED, "syrk: n too large");
  if (impl != LGNN_SYRK_AUTO && impl != LGNN_SYRK_SIMT && impl != LGNN_SYRK_TCGEN05)
    return fail(LGNN_E_BADARG, "syrk: unknown impl %d", impl);
  cudaStream_t st = as_stream(stream);
  bool tc_ok = syrk_tcgen05_supported(k_rows, n) && (ldx % 4 == 0) &&
               ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (impl == LGNN_SYRK_TCGEN05 && !tc_ok)
    return fail(LGNN_E_UNSUPPORTED, "syrk: tcgen05 path needs 8 <= n <= 256, ldx %% 4 == 0, 16-byte aligned x");
  if (impl != LGNN_SYRK_SIMT && tc_ok) {
    if (!ws || ws_bytes < syrk_tcgen05_workspace_bytes(k_rows, n)) return fail(LGNN_E_NOMEM, "syrk: workspace too small");
    return syrk_tcgen05_launch(x, ldx, k_rows, (int)n, alpha, beta, c, ldc, ws, st);
  }
  SimtPlan p = simt_plan(k_rows, n);
  if (!ws || ws_bytes < p.bytes) return fail(LGNN_E_NOMEM, "syrk: workspace %zu < %zu", ws_bytes, p.bytes);
  dim3 grid(p.n_tiles, p.n_slices);
  syrk_simt_kernel<<<grid, SIMT_THREADS, 0, st>>>(x, ldx, k_rows, (int)n, p.tiles_1d, p.rows_per_slice,
                                                  reinterpret_cast<float*>(ws));
  LGNN_LAUNCH_CHECK("syrk_simt_kernel");
  return syrk_reduce_launch(reinterpret_cast<const float*>(ws), p.n_slices, p.tiles_1d, ST, (int)n, alpha,
                            beta, c, ldc, st);
}

}  // extern "C"
