// ABI bookkeeping for libfsg_dense.so (include/fsg_dense.h).
#include "common.cuh"

extern "C" int fsg_abi_version(void) { return FSG_ABI_VERSION; }

extern "C" const char* fsg_status_string(int status) {
  switch (status) {
    case FSG_OK: return "ok";
    case FSG_ERR_INVALID_ARG: return "invalid argument (size, null pointer, alignment or unsupported combination)";
    case FSG_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case FSG_ERR_UNSUPPORTED: return "shape outside what the sm_100a kernels cover";
    default: break;
  }
  if (status >= FSG_ERR_CUDA) return cudaGetErrorString((cudaError_t)(status - FSG_ERR_CUDA));
  return "unknown status";
}
