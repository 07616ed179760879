// LAB copy of the tcgen05 SYRK (syrk_tcgen05.cu compiled a second time with its ablation switches live), behind
// lgnn_syrk_lab_*: tools/syrk_lab.py times the same launch with single stages switched off.  Results of an ablated
// launch are WRONG by construction; nothing in the package calls these entry points.
#define LGNN_TC_ABLATE 1
#define LGNN_TC_ABLATE_VALUE ::lgnn::tc_lab_ablate()
namespace lgnn {
inline int& tc_lab_ablate() {
  static int v = 0;
  return v;
}
}  // namespace lgnn
#define TcGeom TcLabGeom
#define TcParams TcLabParams
#define syrk_tcgen05_kernel syrk_tcgen05_lab_kernel
#define syrk_tc_reduce_kernel syrk_tc_reduce_lab_kernel
#define syrk_tcgen05_supported syrk_tcgen05_lab_supported
#define syrk_tcgen05_workspace_bytes syrk_tcgen05_lab_workspace_bytes
#define syrk_tcgen05_launch syrk_tcgen05_lab_launch
namespace lgnn {
bool syrk_tcgen05_lab_supported(int64_t k_rows, int64_t n);
size_t syrk_tcgen05_lab_workspace_bytes(int64_t k_rows, int64_t n);
}
#include "syrk_tcgen05.cu"

using namespace lgnn;

extern "C" int lgnn_syrk_lab_set_ablate(int bits) {
  tc_lab_ablate() = bits;
  return LGNN_OK;
}
extern "C" size_t lgnn_syrk_lab_workspace_bytes(int64_t k_rows, int64_t n) {
  return syrk_tcgen05_lab_workspace_bytes(k_rows, n);
}
// C = X[:k_rows, :n]^T X[:k_rows, :n] through the lab copy (n <= 256, pitch a multiple of 4 floats, 16-byte aligned)
extern "C" int lgnn_syrk_lab_f32(const float* x, int64_t ldx, int64_t k_rows, int64_t n, float* c, int64_t ldc, void* ws,
                                 lgnn_stream_t stream) {
  if (!x || !c || !ws || !syrk_tcgen05_lab_supported(k_rows, n) || (ldx % 4) || (reinterpret_cast<uintptr_t>(x) & 15))
    return fail(LGNN_E_BADARG, "syrk_lab: bad argument");
  return syrk_tcgen05_lab_launch(x, ldx, k_rows, (int)n, 1.0f, 0.0f, c, ldc, ws, as_stream(stream));
}
