// tcgen05 GEMM kernels (see tc_gemm.cuh) and their C-ABI entry points.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>
#include "tc_gemm.cuh"

using namespace wf;
using namespace wf::tc;

namespace {

// ---------------------------------------------------------------------------------------------- epilogues
// mode 0: C = D (+ bias)            -> out_hi [M][N]
// mode 1: h = tanh(D + bias) split  -> out_hi, out_lo [M][N]  (TF32-exact planes: the next layer's A operand)
struct DenseEpi {
  float* out_hi; float* out_lo; const float* bias; int64_t M; int N; int mode;
  template <int NT>
  __device__ __forceinline__ void run(int m0, int n_tile, int row, uint32_t taddr) const {
    const int64_t r = (int64_t)m0 + row;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + (uint32_t)c0, v);
      if (r < M) {
        const int col = n_tile * NT + c0;
        float* oh = out_hi + r * N + col;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 x = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          if (bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + i)); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
          if (mode == 1) {
            x.x = tanhf(x.x); x.y = tanhf(x.y); x.z = tanhf(x.z); x.w = tanhf(x.w);
            const float4 h = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
            const float4 l = make_float4(tf32_rn(x.x - h.x), tf32_rn(x.y - h.y), tf32_rn(x.z - h.z), tf32_rn(x.w - h.w));
            *reinterpret_cast<float4*>(oh + i) = h;
            *reinterpret_cast<float4*>(out_lo + r * N + col + i) = l;
          } else {
            *reinterpret_cast<float4*>(oh + i) = x;
          }
        }
      }
    }
  }
};

template <int NT>
__global__ void __launch_bounds__(THREADS, 1) tc_dense_kernel(const __grid_constant__ Maps maps, int64_t M, int K, int n_tiles, DenseEpi e) {
  extern __shared__ unsigned char smem_raw[];
  auto epi = [&](int m0, int n_tile, int row, uint32_t taddr) { e.run<NT>(m0, n_tile, row, taddr); };
  gemm_mainloop<NT>(maps, M, K, n_tiles, smem_raw, epi);
}

template <int NT>
int launch_dense(const Maps& maps, int64_t M, int K, int N, const DenseEpi& e, cudaStream_t s) {
  const int n_tiles = N / NT;
  const int64_t tiles = ((M + TILE_M - 1) / TILE_M) * n_tiles;
  const int blocks = (int)(tiles < num_sms() ? tiles : num_sms());
  WF_CUDA(cudaFuncSetAttribute(tc_dense_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<NT>::TOTAL));
  tc_dense_kernel<NT><<<blocks, THREADS, Smem<NT>::TOTAL, s>>>(maps, M, K, n_tiles, e);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

// split kernel: x -> TF32-exact (hi, lo) planes
__global__ void tf32_split_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i], h = tf32_rn(v);
    hi[i] = h; lo[i] = tf32_rn(v - h);
  }
}

}  // namespace

extern "C" int wf_tf32_split(const float* x, int64_t n, float* hi, float* lo, void* stream) {
  if (n == 0) return WF_OK;
  if (!x || !hi || !lo || n < 0) return WF_ERR_INVALID_ARG;
  const int64_t want = (n + 255) / 256;
  tf32_split_kernel<<<(int)(want < 4096 ? want : 4096), 256, 0, (cudaStream_t)stream>>>(x, n, hi, lo);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

extern "C" int wf_tc_dense(const float* a_hi, const float* a_lo, int64_t M, int K, const float* w_hi, const float* w_lo, int N,
                           const float* bias, int mode, float* out_hi, float* out_lo, void* stream) {
  if (M == 0) return WF_OK;
  if (!a_hi || !a_lo || !w_hi || !w_lo || !out_hi || M < 0 || K < KB || (K % KB) || N < 16 || (mode != 0 && mode != 1)) return WF_ERR_INVALID_ARG;
  if (mode == 1 && !out_lo) return WF_ERR_INVALID_ARG;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(a_lo) | reinterpret_cast<uintptr_t>(w_hi) |
                       reinterpret_cast<uintptr_t>(w_lo) | reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo) |
                       reinterpret_cast<uintptr_t>(bias);
  if (al & 15) return WF_ERR_INVALID_ARG;
  const int nt = (N % 256 == 0) ? 256 : ((N % 192 == 0) ? 192 : ((N % 128 == 0) ? 128 : 0));
  if (!nt) return WF_ERR_UNSUPPORTED;
  Maps maps;
  const int st = make_maps(maps, a_hi, a_lo, M, K, w_hi, w_lo, N, nt);
  if (st != WF_OK) return st;
  DenseEpi e{out_hi, out_lo, bias, M, N, mode};
  cudaStream_t s = (cudaStream_t)stream;
  if (nt == 256) return launch_dense<256>(maps, M, K, N, e, s);
  if (nt == 192) return launch_dense<192>(maps, M, K, N, e, s);
  return launch_dense<128>(maps, M, K, N, e, s);
}
