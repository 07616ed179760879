// ABI bookkeeping: version, last error, device info.
#include "common.cuh"

extern "C" {

int lgnn_abi_version(void) { return LGNN_ABI_VERSION; }

const char* lgnn_last_error(void) { return lgnn::err_buf(); }

int lgnn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  LGNN_CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp p;
  LGNN_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return LGNN_OK;
}

}  // extern "C"
