// Internal interface between the SYRK dispatcher (syrk.cu) and the tcgen05 kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace lgnn {

bool syrk_tcgen05_supported(int64_t k_rows, int64_t n);
size_t syrk_tcgen05_workspace_bytes(int64_t k_rows, int64_t n);
int syrk_tcgen05_launch(const float* x, int64_t ldx, int64_t k_rows, int n, float alpha, float beta,
                        float* c, int64_t ldc, void* ws, cudaStream_t st);

}  // namespace lgnn
