// Rational-quadratic spline operator (reference: flows/bijections/neural_splines.py:11-184) for sm_100a.
#include <math.h>
#include "common.cuh"
#include "rqs_device.cuh"

using namespace wf;

// One thread per element.  Each thread streams its own K-float rows of unnormalised widths / heights with 128-bit
// loads (a row is one 128-byte line at K = 32), keeps them in registers for the softmax + left-to-right knot scan,
// and touches only the two derivative entries of the located bin.
template <int KMAX, bool VEC>
__global__ void __launch_bounds__(256)
rqs_kernel(const float* __restrict__ inputs, const float* __restrict__ uw, const float* __restrict__ uh,
           const float* __restrict__ ud, int64_t M, int K, float B, int inverse, float* __restrict__ outputs,
           float* __restrict__ logabsdet, int32_t* __restrict__ bin_idx) {
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float xv = ldg_stream(inputs + m);
    const bool inside = (xv >= -B) && (xv <= B);
    float out = xv, lad = 0.f;
    int bin = -1;
    if (inside) {
      float a[KMAX], b[KMAX];
      load_row<KMAX, VEC>(uw + m * K, K, a);
      load_row<KMAX, VEC>(uh + m * K, K, b);
      const float* udr = ud + m * (K - 1);
      rqs_eval<KMAX>(xv, a, b, K, B, inverse != 0, [&](int j) { return __ldg(udr + j); }, out, lad, bin);
    }
    stg_stream(outputs + m, out);
    stg_stream(logabsdet + m, lad);
    if (bin_idx) bin_idx[m] = bin;
  }
}

template <int KMAX>
static int launch_rqs(const float* inputs, const float* uw, const float* uh, const float* ud, int64_t M, int K,
                      float B, int inverse, float* outputs, float* logabsdet, int32_t* bin_idx, cudaStream_t s) {
  const int threads = 256;
  const int64_t want = (M + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  const int blocks = (int)(want < cap ? want : cap);
  const bool vec = (K % 4 == 0) && !(reinterpret_cast<uintptr_t>(uw) & 15) && !(reinterpret_cast<uintptr_t>(uh) & 15);
  if (vec)
    rqs_kernel<KMAX, true><<<blocks, threads, 0, s>>>(inputs, uw, uh, ud, M, K, B, inverse, outputs, logabsdet, bin_idx);
  else
    rqs_kernel<KMAX, false><<<blocks, threads, 0, s>>>(inputs, uw, uh, ud, M, K, B, inverse, outputs, logabsdet, bin_idx);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

extern "C" int wf_rqs_apply(const float* inputs, const float* uw, const float* uh, const float* ud, int64_t M, int K,
                            float tail_bound, int inverse, float* outputs, float* logabsdet, int32_t* bin_idx,
                            void* stream) {
  if (M == 0) return WF_OK;   // empty batch: nothing to do (pointers of empty buffers may be NULL)
  if (!inputs || !uw || !uh || !ud || !outputs || !logabsdet || M < 0 || K < 2 || !(tail_bound > 0.f)) return WF_ERR_INVALID_ARG;
  if (K > 64) return WF_ERR_UNSUPPORTED;
  // neural_splines.py:93-96
  if (1e-3 * K > 1.0) return WF_ERR_INVALID_ARG;
  if (M == 0) return WF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (K <= 8) return launch_rqs<8>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  if (K <= 16) return launch_rqs<16>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  if (K <= 32) return launch_rqs<32>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  return launch_rqs<64>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
}
