// Tensor-core coupling flow for BASELINE config 5: D = 64, K = 64 bins, width-512 FCNN conditioners
// (flows/bijections/neural_splines.py:187-188,244-296).
//
// Per half-update three tcgen05 GEMMs (see tc_gemm.cuh):
//   h1 = tanh(cond W1 + b1)   [M, 32] x [32, 512]      -> TF32-exact (hi, lo) planes, L2-resident scratch
//   h2 = tanh(h1  W2 + b2)    [M, 512] x [512, 512]    -> planes
//   raw = h2 W3 + b3          [M, 512] x [512, 32*192]    one 128 x 192 accumulator tile per (row block, target dimension),
// and the rational-quadratic spline is evaluated IN THE EPILOGUE of the third GEMM straight from tensor memory: the 191
// raw parameters of (sample, dimension) are read with tcgen05.ld by the thread that owns the accumulator row, so the
// [M, 6112] parameter tensor (24 KB per sample) never exists in HBM.
#include <math.h>
#include "tc_gemm.cuh"
#include "rqs_device.cuh"

using namespace wf;
using namespace wf::tc;

namespace {

constexpr int CD = 64, CHALF = 32, CHID = 512, CK = 64, CNT = 192;     // D, D/2, hidden, bins, padded params per dimension

// packed per conditioner (floats):  W1t_hi [512][32] | W1t_lo | b1 [512] | W2t_hi [512][512] | W2t_lo | b2 [512]
//                                   | W3t_hi [32*192][512] | W3t_lo | b3p [32*192]
constexpr int64_t OFF_W1H = 0, OFF_W1L = OFF_W1H + (int64_t)CHID * CHALF, OFF_B1 = OFF_W1L + (int64_t)CHID * CHALF;
constexpr int64_t OFF_W2H = OFF_B1 + CHID, OFF_W2L = OFF_W2H + (int64_t)CHID * CHID, OFF_B2 = OFF_W2L + (int64_t)CHID * CHID;
constexpr int64_t OFF_W3H = OFF_B2 + CHID, OFF_W3L = OFF_W3H + (int64_t)CHALF * CNT * CHID, OFF_B3 = OFF_W3L + (int64_t)CHALF * CNT * CHID;
constexpr int64_t NET_FLOATS = OFF_B3 + (int64_t)CHALF * CNT;

struct RqsEpi {
  float* x; float* x_hi; float* x_lo;     // [M][64] state (fp32 and its TF32-exact planes), target columns updated in place
  const float* b3p;                       // [32][192]
  float* lad;                             // [M][32] log|det| contributions of this half-update
  int64_t M; int tgt_off; float B; int inverse;

  __device__ __forceinline__ void run(int m0, int j, int row, uint32_t taddr) const {
    const int64_t r = (int64_t)m0 + row;
    const bool live = r < M;
    const float tj = live ? x[r * CD + tgt_off + j] : 0.f;
    const bool inside = (tj >= -B) && (tj <= B);
    const float tc_ = fminf(fmaxf(tj, -B), B);          // every lane runs the spline (tcgen05.ld is warp-collective)
    const float* bj = b3p + j * CNT;
    float a[CK], b[CK];
    {
      float v[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) a[c * 32 + i] = v[i] + __ldg(bj + c * 32 + i);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(taddr + (uint32_t)(CK + c * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) b[c * 32 + i] = v[i] + __ldg(bj + CK + c * 32 + i);
      }
    }
    const float mx_a = softmax_2b<CK, true>(a, CK, 2.f * B);   // neural_splines.py:260-261 (RQS repeats the softmax, quirk Q7)
    const float mx_b = softmax_2b<CK, true>(b, CK, 2.f * B);
    const RqsBin bin = rqs_locate<CK, true>(tc_, a, b, CK, B, inverse != 0, mx_a, mx_b);
    // the two derivative parameters of the located bin: columns 128 + idx - 1 and 128 + idx of this row
    float ud0 = 0.f, ud1 = 0.f;
    {
      float v[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(taddr + (uint32_t)(2 * CK + c * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int jd = c * 32 + i;
          const float raw = v[i] + __ldg(bj + 2 * CK + jd);
          ud0 = (jd == bin.idx - 1) ? raw : ud0;
          ud1 = (jd == bin.idx) ? raw : ud1;
        }
      }
    }
    float out, lad_v;
    rqs_finish(tc_, bin, CK, softplus_f(ud0), softplus_f(ud1), inverse != 0, out, lad_v);   // softplus once here, once in RQS
    if (live) {
      const float y = inside ? out : tj;
      const float h = tf32_rn(y);
      x[r * CD + tgt_off + j] = y;
      x_hi[r * CD + tgt_off + j] = h;
      x_lo[r * CD + tgt_off + j] = tf32_rn(y - h);
      lad[r * CHALF + j] = inside ? lad_v : 0.f;
    }
  }
};

__global__ void __launch_bounds__(THREADS, 1) tc_rqs_kernel(const __grid_constant__ Maps maps, int64_t M, RqsEpi e) {
  extern __shared__ unsigned char smem_raw[];
  auto epi = [&](int m0, int n_tile, int row, uint32_t taddr) { e.run(m0, n_tile, row, taddr); };
  gemm_mainloop<CNT>(maps, M, CHID, CHALF, smem_raw, epi);
}

// logdet[r] += sum_j lad[r][j]   (fixed summation order: deterministic)
__global__ void rowsum_kernel(const float* __restrict__ lad, int64_t M, float* __restrict__ logdet) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CHALF; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(lad + r * CHALF + j);
      s += v.x; s += v.y; s += v.z; s += v.w;
    }
    logdet[r] += s;
  }
}

__global__ void init_state_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ y, float* __restrict__ hi,
                                  float* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i], h = tf32_rn(v);
    y[i] = v; hi[i] = h; lo[i] = tf32_rn(v - h);
  }
}

struct DenseEpiT {     // tanh + split (same as tc_gemm.cu's mode 1, hidden width 512)
  float* out_hi; float* out_lo; const float* bias; int64_t M;
  __device__ __forceinline__ void run(int m0, int n_tile, int row, uint32_t taddr) const {
    const int64_t r = (int64_t)m0 + row;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + (uint32_t)c0, v);
      if (r < M) {
        const int col = n_tile * 256 + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col + i));
          const float4 t = make_float4(tanhf(v[i] + bb.x), tanhf(v[i + 1] + bb.y), tanhf(v[i + 2] + bb.z), tanhf(v[i + 3] + bb.w));
          const float4 h = make_float4(tf32_rn(t.x), tf32_rn(t.y), tf32_rn(t.z), tf32_rn(t.w));
          const float4 l = make_float4(tf32_rn(t.x - h.x), tf32_rn(t.y - h.y), tf32_rn(t.z - h.z), tf32_rn(t.w - h.w));
          *reinterpret_cast<float4*>(out_hi + r * CHID + col + i) = h;
          *reinterpret_cast<float4*>(out_lo + r * CHID + col + i) = l;
        }
      }
    }
  }
};

__global__ void __launch_bounds__(THREADS, 1) tc_hidden_kernel(const __grid_constant__ Maps maps, int64_t M, int K, DenseEpiT e) {
  extern __shared__ unsigned char smem_raw[];
  auto epi = [&](int m0, int n_tile, int row, uint32_t taddr) { e.run(m0, n_tile, row, taddr); };
  gemm_mainloop<256>(maps, M, K, CHID / 256, smem_raw, epi);
}

int grid_for(int64_t M, int n_tiles) {
  const int64_t tiles = ((M + TILE_M - 1) / TILE_M) * n_tiles;
  return (int)(tiles < num_sms() ? tiles : num_sms());
}

}  // namespace

extern "C" int64_t wf_rqs_coupling_tc_net_floats(void) { return NET_FLOATS; }

// workspace floats needed for a chunk of `rows` samples
extern "C" int64_t wf_rqs_coupling_tc_workspace_floats(int64_t rows) {
  return rows * (2 * CD + 4 * CHID + CHALF);      // x planes, h1/h2 planes, lad
}

extern "C" int wf_rqs_coupling_flow_tc(const float* weights, int n_layers, float tail_bound, int inverse, const float* x, int64_t N,
                                       float* y, float* logdet, float* workspace, int64_t workspace_floats, void* stream) {
  if (N == 0) return WF_OK;
  if (!weights || !x || !y || !logdet || !workspace || N < 0 || n_layers < 1 || !(tail_bound > 0.f)) return WF_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(workspace)) & 15) return WF_ERR_INVALID_ARG;
  const int64_t per_row = 2 * CD + 4 * CHID + CHALF;
  int64_t rows = workspace_floats / per_row;
  rows = (rows / TILE_M) * TILE_M;
  if (rows < TILE_M) return WF_ERR_INVALID_ARG;
  if (rows > N) rows = ((N + TILE_M - 1) / TILE_M) * TILE_M;
  cudaStream_t s = (cudaStream_t)stream;
  WF_CUDA(cudaFuncSetAttribute(tc_hidden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<256>::TOTAL));
  WF_CUDA(cudaFuncSetAttribute(tc_rqs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<CNT>::TOTAL));
  WF_CUDA(cudaMemsetAsync(logdet, 0, (size_t)N * sizeof(float), s));

  for (int64_t r0 = 0; r0 < N; r0 += rows) {
    const int64_t M = (N - r0) < rows ? (N - r0) : rows;
    float* xh = workspace;
    float* xl = xh + rows * CD;
    float* h1h = xl + rows * CD;
    float* h1l = h1h + rows * CHID;
    float* h2h = h1l + rows * CHID;
    float* h2l = h2h + rows * CHID;
    float* lad = h2l + rows * CHID;
    float* xs = y + r0 * CD;                       // the state lives in the output buffer
    init_state_kernel<<<1024, 256, 0, s>>>(x + r0 * CD, M * CD, xs, xh, xl);
    WF_LAUNCH_CHECK();
    for (int step = 0; step < 2 * n_layers; ++step) {
      // forward: f1(lower) -> upper, f2(upper) -> lower;  inverse: layers reversed, f2(upper) -> lower, f1(lower) -> upper
      const int layer = inverse ? n_layers - 1 - step / 2 : step / 2;
      const int which = inverse ? 1 - (step & 1) : (step & 1);
      const bool cond_is_lower = which == 0;
      const int cond_off = cond_is_lower ? 0 : CHALF, tgt_off = cond_is_lower ? CHALF : 0;
      const float* net = weights + (int64_t)(layer * 2 + which) * NET_FLOATS;
      Maps m1, m2, m3;
      int st;
      if ((st = make_map(&m1.a_hi, xh + cond_off, M, CHALF, TILE_M, CD)) != WF_OK) return st;
      if ((st = make_map(&m1.a_lo, xl + cond_off, M, CHALF, TILE_M, CD)) != WF_OK) return st;
      if ((st = make_map(&m1.b_hi, net + OFF_W1H, CHID, CHALF, 256)) != WF_OK) return st;
      if ((st = make_map(&m1.b_lo, net + OFF_W1L, CHID, CHALF, 256)) != WF_OK) return st;
      if ((st = make_maps(m2, h1h, h1l, M, CHID, net + OFF_W2H, net + OFF_W2L, CHID, 256)) != WF_OK) return st;
      if ((st = make_maps(m3, h2h, h2l, M, CHID, net + OFF_W3H, net + OFF_W3L, (int64_t)CHALF * CNT, CNT)) != WF_OK) return st;
      tc_hidden_kernel<<<grid_for(M, 2), THREADS, Smem<256>::TOTAL, s>>>(m1, M, CHALF, DenseEpiT{h1h, h1l, net + OFF_B1, M});
      WF_LAUNCH_CHECK();
      tc_hidden_kernel<<<grid_for(M, 2), THREADS, Smem<256>::TOTAL, s>>>(m2, M, CHID, DenseEpiT{h2h, h2l, net + OFF_B2, M});
      WF_LAUNCH_CHECK();
      tc_rqs_kernel<<<grid_for(M, CHALF), THREADS, Smem<CNT>::TOTAL, s>>>(m3, M, RqsEpi{xs, xh, xl, net + OFF_B3, lad, M, tgt_off, tail_bound, inverse ? 1 : 0});
      WF_LAUNCH_CHECK();
      rowsum_kernel<<<512, 256, 0, s>>>(lad, M, logdet + r0);
      WF_LAUNCH_CHECK();
    }
  }
  return WF_OK;
}
