// Shared helpers for the lgnn kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "lgnn.h"

namespace lgnn {

// thread-local last-error string (no global mutable state shared across threads)
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int cuda_fail(cudaError_t e, const char* what) {
  return fail(LGNN_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define LGNN_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::lgnn::cuda_fail(_e, #expr); \
  } while (0)

#define LGNN_LAUNCH_CHECK(name)                                  \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::lgnn::cuda_fail(_e, name);   \
  } while (0)

inline int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
inline T* carve(char*& p, size_t count) {
  T* out = reinterpret_cast<T*>(p);
  p += align_up(count * sizeof(T), 256);
  return out;
}

inline cudaStream_t as_stream(lgnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// exclusive prefix sum of int64 counts (device, in any stream); defined in scan.cu
//   out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 entries.  in may be int32 or int64.
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t n, int64_t* out, void* ws, size_t ws_bytes,
                              cudaStream_t st);

}  // namespace lgnn
