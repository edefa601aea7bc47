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

// Measurement aid for bench.py: a register-resident FFMA loop (8 independent chains per thread) that gives the FP32
// issue ceiling of the device the roofline of the FMA-bound live-path kernels is quoted against.
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, float seed, float* out) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed + (float)(threadIdx.x + k);
  const float m = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 12345.678f) out[0] = s;   // never true: keeps the loop alive
}

// Launches the probe; FLOPs executed = 2 * 8 * iters * blocks * 256.
extern "C" int wf_probe_fma(int iters, int blocks, float* out, void* stream) {
  if (iters <= 0 || blocks <= 0 || !out) return WF_ERR_INVALID_ARG;
  fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, out);
  WF_LAUNCH_CHECK();
  return WF_OK;
}
