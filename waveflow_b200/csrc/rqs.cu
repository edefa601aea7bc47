// Rational-quadratic spline operator (reference: flows/bijections/neural_splines.py:11-184) for sm_100a.
#include <math.h>
#include "common.cuh"
#include "rqs_device.cuh"

using namespace wf;

// One thread per element.  Each thread streams its own K-float rows of unnormalised widths / heights with 128-bit
// loads (a row is one 128-byte line at K = 32), keeps them in registers for the softmax + left-to-right knot scan,
// and touches only the two derivative entries of the located bin.
template <int KMAX, bool VEC, bool EXACT = false>
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
      if (EXACT) rqs_eval_exact<KMAX>(xv, a, b, K, B, inverse != 0, [&](int j) { return __ldg(udr + j); }, out, lad, bin);
      else rqs_eval<KMAX>(xv, a, b, K, B, inverse != 0, [&](int j) { return __ldg(udr + j); }, out, lad, bin);
    }
    stg_stream(outputs + m, out);
    stg_stream(logabsdet + m, lad);
    if (bin_idx) bin_idx[m] = bin;
  }
}

// Staged variant for K = 32 / 64: the K-float rows of a tile are brought in with fully coalesced 16-byte cp.async
// (LDGSTS) copies into an XOR-swizzled shared-memory tile, then each thread pulls ITS row into registers with
// conflict-free 128-bit loads.  Once the rows sit in registers the tile is free again, so the next tile's copies are
// issued before the (issue-bound) spline arithmetic of the current one: full load/compute overlap with a single buffer.
// The direct-load kernel above makes every warp request touch 32 different 128-byte lines (ncu: L1 wavefronts, not HBM,
// were the limit at 34 % of the measured bandwidth).
constexpr int RQS_TILE = 256;

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int K>
__global__ void __launch_bounds__(RQS_TILE, 2)
rqs_staged_kernel(const float* __restrict__ inputs, const float* __restrict__ uw, const float* __restrict__ uh,
                  const float* __restrict__ ud, int64_t M, float B, int inverse, float* __restrict__ outputs,
                  float* __restrict__ logabsdet, int32_t* __restrict__ bin_idx) {
  constexpr int KC = K / 4;                                  // 16-byte chunks per row
  extern __shared__ __align__(128) float4 tile[];            // [2 arrays][RQS_TILE rows][KC chunks], swizzled
  float4* tw = tile;
  float4* th = tile + RQS_TILE * KC;
  const int tid = threadIdx.x;
  const int64_t n_tiles = (M + RQS_TILE - 1) / RQS_TILE;

  auto issue = [&](int64_t t) {
    const int64_t row0 = t * RQS_TILE;
    const int rows = (int)((M - row0) < RQS_TILE ? (M - row0) : RQS_TILE);
    const float4* gw = reinterpret_cast<const float4*>(uw + row0 * K);
    const float4* gh = reinterpret_cast<const float4*>(uh + row0 * K);
    for (int g = tid; g < rows * KC; g += RQS_TILE) {
      const int r = g / KC, c = g % KC;
      const int dst = r * KC + (c ^ (r & (KC - 1) & 7));
      cp_async16(tw + dst, gw + g);
      cp_async16(th + dst, gh + g);
    }
    cp_async_commit();
  };

  int64_t t = blockIdx.x;
  if (t < n_tiles) issue(t);
  for (; t < n_tiles; t += gridDim.x) {
    const int64_t m = t * RQS_TILE + tid;
    const bool live = m < M;
    const float xv = live ? ldg_stream(inputs + m) : 0.f;
    cp_async_wait_all();
    __syncthreads();
    float a[K], b[K];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const int src = tid * KC + (c ^ (tid & (KC - 1) & 7));
      const float4 va = tw[src], vb = th[src];
      a[4 * c] = va.x; a[4 * c + 1] = va.y; a[4 * c + 2] = va.z; a[4 * c + 3] = va.w;
      b[4 * c] = vb.x; b[4 * c + 1] = vb.y; b[4 * c + 2] = vb.z; b[4 * c + 3] = vb.w;
    }
    __syncthreads();
    if (t + gridDim.x < n_tiles) issue(t + gridDim.x);
    if (live) {
      const bool inside = (xv >= -B) && (xv <= B);
      float out = xv, lad = 0.f;
      int bin = -1;
      if (inside) {
        const float* udr = ud + m * (K - 1);
        rqs_eval<K, true>(xv, a, b, K, B, inverse != 0, [&](int j) { return __ldg(udr + j); }, out, lad, bin);
      }
      stg_stream(outputs + m, out);
      stg_stream(logabsdet + m, lad);
      if (bin_idx) bin_idx[m] = bin;
    }
  }
  cp_async_wait_all();
}

template <int K>
static int launch_rqs_staged(const float* inputs, const float* uw, const float* uh, const float* ud, int64_t M, float B,
                             int inverse, float* outputs, float* logabsdet, int32_t* bin_idx, cudaStream_t s) {
  const size_t smem = (size_t)2 * RQS_TILE * K * sizeof(float);
  const int64_t n_tiles = (M + RQS_TILE - 1) / RQS_TILE;
  const int per_sm = (int)((227 * 1024) / (smem + 1024)) < 2 ? 1 : 2;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  const int blocks = (int)(n_tiles < cap ? n_tiles : cap);
  WF_CUDA(cudaFuncSetAttribute(rqs_staged_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rqs_staged_kernel<K><<<blocks, RQS_TILE, smem, s>>>(inputs, uw, uh, ud, M, B, inverse, outputs, logabsdet, bin_idx);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

template <int KMAX, bool EXACT = false>
static int launch_rqs(const float* inputs, const float* uw, const float* uh, const float* ud, int64_t M, int K,
                      float B, int inverse, float* outputs, float* logabsdet, int32_t* bin_idx, cudaStream_t s) {
  const int threads = 256;
  const int64_t want = (M + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  const int blocks = (int)(want < cap ? want : cap);
  const bool vec = (K % 4 == 0) && !(reinterpret_cast<uintptr_t>(uw) & 15) && !(reinterpret_cast<uintptr_t>(uh) & 15);
  if (vec)
    rqs_kernel<KMAX, true, EXACT><<<blocks, threads, 0, s>>>(inputs, uw, uh, ud, M, K, B, inverse, outputs, logabsdet, bin_idx);
  else
    rqs_kernel<KMAX, false, EXACT><<<blocks, threads, 0, s>>>(inputs, uw, uh, ud, M, K, B, inverse, outputs, logabsdet, bin_idx);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

extern "C" int wf_rqs_apply(const float* inputs, const float* uw, const float* uh, const float* ud, int64_t M, int K,
                            float tail_bound, int flags, float* outputs, float* logabsdet, int32_t* bin_idx,
                            void* stream) {
  if (flags & ~(WF_RQS_INVERSE | WF_RQS_EXACT_BINS)) return WF_ERR_INVALID_ARG;
  const int inverse = (flags & WF_RQS_INVERSE) ? 1 : 0;
  if (M == 0) return WF_OK;   // empty batch: nothing to do (pointers of empty buffers may be NULL)
  if (!inputs || !uw || !uh || !ud || !outputs || !logabsdet || M < 0 || K < 2 || !(tail_bound > 0.f)) return WF_ERR_INVALID_ARG;
  if (K > 64) return WF_ERR_UNSUPPORTED;
  // neural_splines.py:93-96
  if (1e-3 * K > 1.0) return WF_ERR_INVALID_ARG;
  if (M == 0) return WF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (flags & WF_RQS_EXACT_BINS) {
    if (K <= 8) return launch_rqs<8, true>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
    if (K <= 16) return launch_rqs<16, true>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
    if (K <= 32) return launch_rqs<32, true>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
    return launch_rqs<64, true>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  }
  const bool aligned = !(reinterpret_cast<uintptr_t>(uw) & 15) && !(reinterpret_cast<uintptr_t>(uh) & 15);
  if (aligned && K == 32 && M >= 4 * RQS_TILE) return launch_rqs_staged<32>(inputs, uw, uh, ud, M, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  if (aligned && K == 64 && M >= 4 * RQS_TILE) return launch_rqs_staged<64>(inputs, uw, uh, ud, M, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  if (K <= 8) return launch_rqs<8>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  if (K <= 16) return launch_rqs<16>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  if (K <= 32) return launch_rqs<32>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
  return launch_rqs<64>(inputs, uw, uh, ud, M, K, tail_bound, inverse, outputs, logabsdet, bin_idx, s);
}
