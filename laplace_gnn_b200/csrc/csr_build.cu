// Integer kernels: edge list -> CSR(A) with self loops, pattern transpose, degree normalisation,
// edge values, nnz-balanced row partition, halo marking, partition slicing.
//
// All results are order-independent (rows are bucketed with atomics, then every row is sorted
// and de-duplicated), so the output is bit-identical to the CPU oracle
// (oracle/gcn_kfac_oracle.py: coo_to_adj_csr, csr_transpose_pattern, normalize_adj_csr,
// row_partition, halo_columns).  HBM-bound integer work, run once per graph.
#include "common.cuh"

namespace lgnn {

constexpr int BUILD_THREADS = 256;
constexpr int WARP_SORT_CAP = 256;    // row length handled by one warp in shared memory
constexpr int BLOCK_SORT_CAP = 8192;  // row length handled by one block in shared memory

// ---------------------------------------------------------------- workspace layout (build)
struct BuildWs {
  int32_t* err;        // [2]  err[0] != 0: an edge endpoint was out of range; err[1]: #long rows
  int32_t* cnt;        // [n]  entries per row incl. duplicates (upper bound)
  int32_t* cursor;     // [n]
  int32_t* ucnt;       // [n]  unique entries per row
  int32_t* long_rows;  // [n]  rows longer than WARP_SORT_CAP
  int64_t* tmp_rowptr; // [n+1]
  int32_t* tmp_col;    // [total]
  void* scan_ws;
  size_t scan_bytes;
  size_t total_bytes;
};

static BuildWs carve_build(void* ws, int64_t n, int64_t total) {
  BuildWs w;
  char* p = reinterpret_cast<char*>(ws);
  char* p0 = p;
  w.err = carve<int32_t>(p, 64);
  w.cnt = carve<int32_t>(p, n);
  w.cursor = carve<int32_t>(p, n);
  w.ucnt = carve<int32_t>(p, n);
  w.long_rows = carve<int32_t>(p, n);
  w.tmp_rowptr = carve<int64_t>(p, n + 1);
  w.tmp_col = carve<int32_t>(p, total > 0 ? total : 1);
  w.scan_bytes = scan_workspace_bytes(n);
  w.scan_ws = p;
  p += w.scan_bytes;
  w.total_bytes = (size_t)(p - p0);
  return w;
}

// ---------------------------------------------------------------- kernels
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void count_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                   int64_t n_edges, int64_t n, int symmetric,
                                   int32_t* __restrict__ cnt, int32_t* __restrict__ err) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_edges; e += stride) {
    int64_t s = src[e], d = dst[e];
    if (s < 0 || s >= n || d < 0 || d >= n) {
      err[0] = 1;
      continue;
    }
    atomicAdd(&cnt[s], 1);
    if (symmetric) atomicAdd(&cnt[d], 1);
  }
}

__global__ void scatter_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                     int64_t n_edges, int64_t n, int symmetric,
                                     const int64_t* __restrict__ tmp_rowptr,
                                     int32_t* __restrict__ cursor, int32_t* __restrict__ tmp_col) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_edges + n; e += stride) {
    int64_t s, d;
    if (e < n_edges) {
      s = src[e];
      d = dst[e];
      if (s < 0 || s >= n || d < 0 || d >= n) continue;
    } else {
      s = d = e - n_edges;  // self loop
    }
    int32_t pos = atomicAdd(&cursor[s], 1);
    tmp_col[tmp_rowptr[s] + pos] = (int32_t)d;
    if (symmetric && e < n_edges) {
      int32_t pos2 = atomicAdd(&cursor[d], 1);
      tmp_col[tmp_rowptr[d] + pos2] = (int32_t)s;
    }
  }
}

// Normalised bitonic network on `len` keys (any len): every comparator puts the minimum at the
// lower index, so virtual +inf padding above `len` never moves and its comparators are skipped.
// `tid`/`nthreads` enumerate the cooperating threads; SYNC separates the stages.
template <typename SyncFn>
__device__ __forceinline__ void bitonic_sort_any(int32_t* a, int len, int tid, int nthreads,
                                                 SyncFn sync) {
  int p2 = 1;
  while (p2 < len) p2 <<= 1;
  for (int k = 2; k <= p2; k <<= 1) {
    // first stage of the merge: compare i with its mirror inside the k-block
    for (int t = tid; t < p2 / 2; t += nthreads) {
      int blk = t / (k / 2), off = t % (k / 2);
      int i = blk * k + off;
      int j = blk * k + (k - 1 - off);
      if (j < len) {
        int32_t x = a[i], y = a[j];
        if (x > y) { a[i] = y; a[j] = x; }
      }
    }
    sync();
    for (int s = k / 4; s >= 1; s >>= 1) {
      for (int t = tid; t < p2 / 2; t += nthreads) {
        int i = (t / s) * (2 * s) + (t % s);
        int j = i + s;
        if (j < len) {
          int32_t x = a[i], y = a[j];
          if (x > y) { a[i] = y; a[j] = x; }
        }
      }
      sync();
    }
  }
}

// warp per row: rows with len <= WARP_SORT_CAP are sorted in shared memory, longer rows are queued.
__global__ void __launch_bounds__(BUILD_THREADS) sort_rows_warp_kernel(
    const int64_t* __restrict__ rowptr, int32_t* __restrict__ colbuf, int64_t n,
    int32_t* __restrict__ ucnt, int32_t* __restrict__ long_rows, int32_t* __restrict__ n_long) {
  __shared__ int32_t sm[BUILD_THREADS / 32][WARP_SORT_CAP];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t row = (int64_t)blockIdx.x * (BUILD_THREADS / 32) + w;
  int64_t stride = (int64_t)gridDim.x * (BUILD_THREADS / 32);
  for (; row < n; row += stride) {
    int64_t b = rowptr[row], e = rowptr[row + 1];
    int64_t len64 = e - b;
    if (len64 > WARP_SORT_CAP) {
      if (lane == 0) long_rows[atomicAdd(n_long, 1)] = (int32_t)row;
      continue;
    }
    int len = (int)len64;
    int32_t* a = sm[w];
    for (int i = lane; i < len; i += 32) a[i] = colbuf[b + i];
    __syncwarp();
    if (len > 1) bitonic_sort_any(a, len, lane, 32, [] { __syncwarp(); });
    int uniq = 0;
    for (int i0 = 0; i0 < len; i0 += 32) {
      int i = i0 + lane;
      bool first = i < len && (i == 0 || a[i] != a[i - 1]);
      uniq += __popc(__ballot_sync(0xffffffffu, first));
      if (i < len) colbuf[b + i] = a[i];
    }
    if (ucnt != nullptr && lane == 0) ucnt[row] = uniq;
    __syncwarp();
  }
}

// block per queued long row: shared memory up to BLOCK_SORT_CAP, else in place in global memory.
__global__ void __launch_bounds__(BUILD_THREADS) sort_rows_block_kernel(
    const int64_t* __restrict__ rowptr, int32_t* __restrict__ colbuf,
    int32_t* __restrict__ ucnt, const int32_t* __restrict__ long_rows,
    const int32_t* __restrict__ n_long) {
  __shared__ int32_t sm[BLOCK_SORT_CAP];
  __shared__ int uniq_sh;
  int nl = *n_long;
  for (int q = blockIdx.x; q < nl; q += gridDim.x) {
    int64_t row = long_rows[q];
    int64_t b = rowptr[row], e = rowptr[row + 1];
    int len = (int)(e - b);  // rows are bounded by 2n+1 < 2^31 after bucketing by a 31-bit node id
    int32_t* a;
    if (len <= BLOCK_SORT_CAP) {
      for (int i = threadIdx.x; i < len; i += blockDim.x) sm[i] = colbuf[b + i];
      a = sm;
    } else {
      a = colbuf + b;
    }
    if (threadIdx.x == 0) uniq_sh = 0;
    __syncthreads();
    bitonic_sort_any(a, len, (int)threadIdx.x, (int)blockDim.x, [] { __syncthreads(); });
    int local = 0;
    for (int i = threadIdx.x; i < len; i += blockDim.x)
      local += (i == 0 || a[i] != a[i - 1]) ? 1 : 0;
    if (len <= BLOCK_SORT_CAP)
      for (int i = threadIdx.x; i < len; i += blockDim.x) colbuf[b + i] = a[i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&uniq_sh, local);
    __syncthreads();
    if (ucnt != nullptr && threadIdx.x == 0) ucnt[row] = uniq_sh;
    __syncthreads();
  }
}

// warp per row: copy the unique entries of the sorted row into the final CSR
__global__ void __launch_bounds__(BUILD_THREADS) compact_rows_kernel(
    const int64_t* __restrict__ tmp_rowptr, const int32_t* __restrict__ tmp_col, int64_t n,
    const int64_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; row < n; row += stride) {
    int64_t b = tmp_rowptr[row], e = tmp_rowptr[row + 1];
    int64_t out = rowptr[row];
    for (int64_t i0 = b; i0 < e; i0 += 32) {
      int64_t i = i0 + lane;
      int32_t v = i < e ? tmp_col[i] : 0;
      bool first = i < e && (i == b || v != tmp_col[i - 1]);
      unsigned m = __ballot_sync(0xffffffffu, first);
      if (first) col[out + __popc(m & ((1u << lane) - 1u))] = v;
      out += __popc(m);
    }
  }
}

// ---- transpose
__global__ void __launch_bounds__(BUILD_THREADS) count_cols_rows_kernel(
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
    int32_t* __restrict__ cnt) {
  // warp per row so that nnz (= rowptr[n_rows], a device value) is never needed on the host
  int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; row < n_rows; row += stride) {
    int64_t b = rowptr[row], e = rowptr[row + 1];
    for (int64_t k = b + lane; k < e; k += 32) atomicAdd(&cnt[col[k]], 1);
  }
}

__global__ void __launch_bounds__(BUILD_THREADS) scatter_transpose_kernel(
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
    const int64_t* __restrict__ t_rowptr, int32_t* __restrict__ cursor,
    int32_t* __restrict__ t_col) {
  int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; row < n_rows; row += stride) {
    int64_t b = rowptr[row], e = rowptr[row + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
      int32_t c = col[k];
      int32_t pos = atomicAdd(&cursor[c], 1);
      t_col[t_rowptr[c] + pos] = (int32_t)row;
    }
  }
}

// ---- normalisation
__global__ void degree_norm_kernel(int64_t n, const int64_t* __restrict__ rowptr,
                                   int64_t* __restrict__ deg, float* __restrict__ dis) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t d = rowptr[i + 1] - rowptr[i];
    if (deg != nullptr) deg[i] = d;
    // IEEE-rounded sqrt and divide: reproducible to the bit against numpy float32
    float f = d > 0 ? __fdiv_rn(1.0f, __fsqrt_rn(__ll2float_rn(d))) : 0.0f;
    dis[i] = f;
  }
}

__global__ void __launch_bounds__(BUILD_THREADS) edge_values_kernel(
    int64_t n_rows, int64_t row_offset, const int64_t* __restrict__ rowptr,
    const int32_t* __restrict__ col, const float* __restrict__ dis, float* __restrict__ val) {
  int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; row < n_rows; row += stride) {
    float di = dis[row_offset + row];
    int64_t b = rowptr[row], e = rowptr[row + 1];
    for (int64_t k = b + lane; k < e; k += 32) val[k] = __fmul_rn(di, dis[col[k]]);
  }
}

// ---- partition
__global__ void row_partition_kernel(const int64_t* __restrict__ rowptr, int64_t n, int32_t nparts,
                                     int64_t* __restrict__ bounds) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nparts) return;
  if (r == 0) { bounds[0] = 0; return; }
  if (r == nparts) { bounds[nparts] = n; return; }
  int64_t nnz = rowptr[n];
  // floor(r * nnz / nparts) without overflow for nnz < 2^56
  int64_t target = (int64_t)(((unsigned __int128)r * (unsigned __int128)nnz) / (unsigned)nparts);
  int64_t lo = 0, hi = n + 1;  // first i in [0, n] with rowptr[i] >= target
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] >= target) hi = mid; else lo = mid + 1;
  }
  bounds[r] = lo < n ? lo : n;
}

__global__ void halo_mark_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 int64_t lo, int64_t hi, uint8_t* __restrict__ flags) {
  int64_t b = rowptr[lo], e = rowptr[hi];
  int64_t k = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; k < e; k += stride) {
    int64_t c = col[k];
    if (c < lo || c >= hi) flags[c] = 1;
  }
}

__global__ void __launch_bounds__(BUILD_THREADS) csr_slice_remap_kernel(
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
    const float* __restrict__ val, int64_t lo, int64_t hi, const int64_t* __restrict__ bounds,
    int32_t nparts, int64_t pad, int64_t* __restrict__ out_rowptr, int32_t* __restrict__ out_col,
    float* __restrict__ out_val) {
  int lane = threadIdx.x & 31;
  int64_t base = rowptr[lo];
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; r <= hi - lo; r += stride) {
    if (lane == 0) out_rowptr[r] = rowptr[lo + r] - base;
    if (r == hi - lo) continue;
    int64_t b = rowptr[lo + r], e = rowptr[lo + r + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
      int64_t c = col[k];
      int q_lo = 0, q_hi = nparts;  // owner q: bounds[q] <= c < bounds[q+1]
      while (q_hi - q_lo > 1) {
        int mid = (q_lo + q_hi) >> 1;
        if (bounds[mid] <= c) q_lo = mid; else q_hi = mid;
      }
      out_col[k - base] = (int32_t)((int64_t)q_lo * pad + (c - bounds[q_lo]));
      if (val != nullptr) out_val[k - base] = val[k];
    }
  }
}

static inline unsigned grid_for(int64_t work, int threads, int per_sm = 8) {
  int64_t blocks = (work + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static int sort_rows(const int64_t* rowptr, int32_t* colbuf, int64_t n, int32_t* ucnt,
                     int32_t* long_rows, int32_t* n_long, cudaStream_t st) {
  LGNN_CUDA_TRY(cudaMemsetAsync(n_long, 0, sizeof(int32_t), st));
  sort_rows_warp_kernel<<<grid_for(n * 32, BUILD_THREADS, 8), BUILD_THREADS, 0, st>>>(
      rowptr, colbuf, n, ucnt, long_rows, n_long);
  LGNN_LAUNCH_CHECK("sort_rows_warp_kernel");
  sort_rows_block_kernel<<<sm_count() * 4, BUILD_THREADS, 0, st>>>(rowptr, colbuf, ucnt, long_rows,
                                                                   n_long);
  LGNN_LAUNCH_CHECK("sort_rows_block_kernel");
  return LGNN_OK;
}

}  // namespace lgnn

using namespace lgnn;

extern "C" {

size_t lgnn_csr_build_workspace_bytes(int64_t n, int64_t n_edges, int symmetric) {
  if (n < 0 || n_edges < 0) return 0;
  int64_t total = n_edges * (symmetric ? 2 : 1) + n;
  return carve_build(nullptr, n, total).total_bytes;
}

int lgnn_csr_build_count(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n,
                         int symmetric, void* ws, size_t ws_bytes, int64_t* rowptr,
                         lgnn_stream_t stream) {
  if (n < 0 || n_edges < 0 || n >= (int64_t)INT32_MAX) return fail(LGNN_E_BADARG, "csr_build: bad n / n_edges");
  if ((n_edges > 0 && (!src || !dst)) || !ws || !rowptr) return fail(LGNN_E_BADARG, "csr_build: null pointer");
  int64_t total = n_edges * (symmetric ? 2 : 1) + n;
  BuildWs w = carve_build(ws, n, total);
  if (ws_bytes < w.total_bytes) return fail(LGNN_E_NOMEM, "csr_build: workspace %zu < %zu", ws_bytes, w.total_bytes);
  cudaStream_t st = as_stream(stream);
  LGNN_CUDA_TRY(cudaMemsetAsync(w.err, 0, 64 * sizeof(int32_t), st));
  if (n == 0) {
    LGNN_CUDA_TRY(cudaMemsetAsync(rowptr, 0, sizeof(int64_t), st));
    return LGNN_OK;
  }
  fill_i32_kernel<<<grid_for(n, BUILD_THREADS), BUILD_THREADS, 0, st>>>(w.cnt, n, 1);  // self loops
  LGNN_LAUNCH_CHECK("fill_i32_kernel");
  LGNN_CUDA_TRY(cudaMemsetAsync(w.cursor, 0, (size_t)n * sizeof(int32_t), st));
  if (n_edges > 0) {
    count_edges_kernel<<<grid_for(n_edges, BUILD_THREADS), BUILD_THREADS, 0, st>>>(
        src, dst, n_edges, n, symmetric, w.cnt, w.err);
    LGNN_LAUNCH_CHECK("count_edges_kernel");
  }
  int rc = exclusive_scan_i32_to_i64(w.cnt, n, w.tmp_rowptr, w.scan_ws, w.scan_bytes, st);
  if (rc) return rc;
  scatter_edges_kernel<<<grid_for(n_edges + n, BUILD_THREADS), BUILD_THREADS, 0, st>>>(
      src, dst, n_edges, n, symmetric, w.tmp_rowptr, w.cursor, w.tmp_col);
  LGNN_LAUNCH_CHECK("scatter_edges_kernel");
  rc = sort_rows(w.tmp_rowptr, w.tmp_col, n, w.ucnt, w.long_rows, w.err + 1, st);
  if (rc) return rc;
  return exclusive_scan_i32_to_i64(w.ucnt, n, rowptr, w.scan_ws, w.scan_bytes, st);
}

int lgnn_csr_build_fill(const void* ws, size_t ws_bytes, int64_t n, int64_t n_edges, int symmetric,
                        const int64_t* rowptr, int32_t* col, lgnn_stream_t stream) {
  if (n < 0 || n_edges < 0 || !ws || !rowptr) return fail(LGNN_E_BADARG, "csr_build_fill: bad argument");
  if (n == 0) return LGNN_OK;
  if (!col) return fail(LGNN_E_BADARG, "csr_build_fill: null col");
  int64_t total = n_edges * (symmetric ? 2 : 1) + n;
  BuildWs w = carve_build(const_cast<void*>(ws), n, total);
  if (ws_bytes < w.total_bytes) return fail(LGNN_E_NOMEM, "csr_build_fill: workspace too small");
  cudaStream_t st = as_stream(stream);
  compact_rows_kernel<<<grid_for(n * 32, BUILD_THREADS, 8), BUILD_THREADS, 0, st>>>(
      w.tmp_rowptr, w.tmp_col, n, rowptr, col);
  LGNN_LAUNCH_CHECK("compact_rows_kernel");
  return LGNN_OK;
}

size_t lgnn_csr_transpose_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t nnz) {
  (void)n_rows; (void)nnz;
  if (n_cols < 0) return 0;
  // cnt, cursor, long_rows, n_long, scan
  return align_up((size_t)n_cols * 4, 256) * 3 + 256 + scan_workspace_bytes(n_cols);
}

int lgnn_csr_transpose(int64_t n_rows, int64_t n_cols, const int64_t* rowptr, const int32_t* col,
                       int64_t* t_rowptr, int32_t* t_col, void* ws, size_t ws_bytes,
                       lgnn_stream_t stream) {
  if (n_rows < 0 || n_cols < 0 || !rowptr || !t_rowptr || !ws) return fail(LGNN_E_BADARG, "csr_transpose: bad argument");
  if (ws_bytes < lgnn_csr_transpose_workspace_bytes(n_rows, n_cols, 0)) return fail(LGNN_E_NOMEM, "csr_transpose: workspace too small");
  cudaStream_t st = as_stream(stream);
  char* p = reinterpret_cast<char*>(ws);
  int32_t* n_long = carve<int32_t>(p, 64);
  int32_t* cnt = carve<int32_t>(p, n_cols);
  int32_t* cursor = carve<int32_t>(p, n_cols);
  int32_t* long_rows = carve<int32_t>(p, n_cols);
  void* scan_ws = p;
  if (n_cols == 0) {
    LGNN_CUDA_TRY(cudaMemsetAsync(t_rowptr, 0, sizeof(int64_t), st));
    return LGNN_OK;
  }
  LGNN_CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)n_cols * 4, st));
  LGNN_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)n_cols * 4, st));
  count_cols_rows_kernel<<<grid_for(n_rows * 32, BUILD_THREADS, 8), BUILD_THREADS, 0, st>>>(rowptr, col, n_rows, cnt);
  LGNN_LAUNCH_CHECK("count_cols_rows_kernel");
  int rc = exclusive_scan_i32_to_i64(cnt, n_cols, t_rowptr, scan_ws, scan_workspace_bytes(n_cols), st);
  if (rc) return rc;
  scatter_transpose_kernel<<<grid_for(n_rows * 32, BUILD_THREADS, 8), BUILD_THREADS, 0, st>>>(
      rowptr, col, n_rows, t_rowptr, cursor, t_col);
  LGNN_LAUNCH_CHECK("scatter_transpose_kernel");
  return sort_rows(t_rowptr, t_col, n_cols, nullptr, long_rows, n_long, st);
}

int lgnn_degree_norm(int64_t n, const int64_t* a_rowptr, int64_t* deg, float* dis,
                     lgnn_stream_t stream) {
  if (n < 0 || !a_rowptr || !dis) return fail(LGNN_E_BADARG, "degree_norm: bad argument");
  if (n == 0) return LGNN_OK;
  degree_norm_kernel<<<grid_for(n, BUILD_THREADS), BUILD_THREADS, 0, as_stream(stream)>>>(n, a_rowptr, deg, dis);
  LGNN_LAUNCH_CHECK("degree_norm_kernel");
  return LGNN_OK;
}

int lgnn_edge_values(int64_t n_rows, int64_t row_offset, const int64_t* rowptr, const int32_t* col,
                     const float* dis, float* val, lgnn_stream_t stream) {
  if (n_rows < 0 || !rowptr || !dis) return fail(LGNN_E_BADARG, "edge_values: bad argument");
  if (n_rows == 0) return LGNN_OK;
  edge_values_kernel<<<grid_for(n_rows * 32, BUILD_THREADS, 8), BUILD_THREADS, 0, as_stream(stream)>>>(
      n_rows, row_offset, rowptr, col, dis, val);
  LGNN_LAUNCH_CHECK("edge_values_kernel");
  return LGNN_OK;
}

int lgnn_row_partition(const int64_t* rowptr, int64_t n, int32_t nparts, int64_t* bounds,
                       lgnn_stream_t stream) {
  if (!rowptr || !bounds || n < 0 || nparts < 1) return fail(LGNN_E_BADARG, "row_partition: bad argument");
  row_partition_kernel<<<(nparts + 1 + 63) / 64, 64, 0, as_stream(stream)>>>(rowptr, n, nparts, bounds);
  LGNN_LAUNCH_CHECK("row_partition_kernel");
  return LGNN_OK;
}

int lgnn_halo_mark(const int64_t* rowptr, const int32_t* col, int64_t lo, int64_t hi,
                   uint8_t* flags, lgnn_stream_t stream) {
  if (!rowptr || !flags || lo < 0 || hi < lo) return fail(LGNN_E_BADARG, "halo_mark: bad argument");
  if (hi == lo) return LGNN_OK;
  halo_mark_kernel<<<sm_count() * 8, BUILD_THREADS, 0, as_stream(stream)>>>(rowptr, col, lo, hi, flags);
  LGNN_LAUNCH_CHECK("halo_mark_kernel");
  return LGNN_OK;
}

int lgnn_csr_slice_remap(const int64_t* rowptr, const int32_t* col, const float* val, int64_t lo,
                         int64_t hi, const int64_t* bounds, int32_t nparts, int64_t pad,
                         int64_t* out_rowptr, int32_t* out_col, float* out_val,
                         lgnn_stream_t stream) {
  if (!rowptr || !bounds || !out_rowptr || lo < 0 || hi < lo || nparts < 1 || pad < 0)
    return fail(LGNN_E_BADARG, "csr_slice_remap: bad argument");
  if ((int64_t)nparts * pad >= (int64_t)INT32_MAX) return fail(LGNN_E_UNSUPPORTED, "csr_slice_remap: padded index exceeds int32");
  csr_slice_remap_kernel<<<grid_for((hi - lo + 1) * 32, BUILD_THREADS, 8), BUILD_THREADS, 0, as_stream(stream)>>>(
      rowptr, col, val, lo, hi, bounds, nparts, pad, out_rowptr, out_col, out_val);
  LGNN_LAUNCH_CHECK("csr_slice_remap_kernel");
  return LGNN_OK;
}

}  // extern "C"
