// ABI bookkeeping entry points.
#include "common.cuh"

extern "C" int wf_abi_version(int* sm_arch) {
  if (sm_arch) *sm_arch = 100;
  return WF_ABI_VERSION;
}

extern "C" const char* wf_status_string(int status) {
  switch (status) {
    case WF_OK: return "ok";
    case WF_ERR_INVALID_ARG: return "invalid argument";
    case WF_ERR_UNSUPPORTED: return "unsupported configuration";
    case WF_ERR_NO_DEVICE: return "no CUDA device";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown waveflow_b200 status";
  }
}
