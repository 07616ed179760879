// LAB copy of the fused GEMM (gemm_mask.cu compiled a second time with its ablation switches live), behind
// lgnn_gemm_mask_lab_*: tools/gemm_lab.py times the same launch with single stages switched off to see which one
// paces a tile.  Results of an ablated launch are WRONG by construction; nothing in the package
// calls these entry points.
#define LGNN_GM_ABLATE 1
#define LGNN_GM_ABLATE_VALUE ::lgnn::gm_lab_ablate()
namespace lgnn {
inline int& gm_lab_ablate() {
  static int v = 0;
  return v;
}
}  // namespace lgnn
// every global symbol of gemm_mask.cu gets a lab name
#define GmParams GmLabParams
#define gemm_mask_kernel gemm_mask_lab_kernel
#define gemm_mask_prepare_kernel gemm_mask_lab_prepare_kernel
#define tmem_st_x16 tmem_st_x16_lab
#define mbar_arrive_remote mbar_arrive_remote_lab
#define lgnn_gemm_mask_supported lgnn_gemm_mask_lab_supported
#define lgnn_gemm_mask_kpad lgnn_gemm_mask_lab_kpad
#define lgnn_gemm_mask_prepare_f32 lgnn_gemm_mask_lab_prepare_f32
#define lgnn_gemm_mask_f32 lgnn_gemm_mask_lab_f32
#define lgnn_gemm_bias_f32 lgnn_gemm_bias_lab_f32
#include "gemm_mask.cu"

extern "C" int lgnn_gemm_mask_lab_set_ablate(int bits) {
  ::lgnn::gm_lab_ablate() = bits;
  return LGNN_OK;
}
