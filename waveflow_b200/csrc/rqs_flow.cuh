// Fused Serial(NeuralSplineCoupling x L) -- reference: flows/bijections/neural_splines.py:187-188,244-296.
//
// One thread per sample carries its D coordinates through all 2L half-updates: FCNN conditioner (Dense-Tanh-Dense-Tanh-
// Dense) -> per target dimension (W, H, D) = (2B softmax, 2B softmax, softplus) -> unconstrained_RQS (which applies
// softmax / softplus AGAIN, quirk Q7) -> log-det.  The 3K-1 raw spline parameters of a dimension never leave registers
// (at D=64, K=64 they would be 24 KB per sample); conditioner weights stream through a shared-memory ring of
// chunks filled by cp.async.bulk (TMA): [W1 b1 W2 b2] then one [W3_j b3_j] block per target dimension.
#pragma once
#include <math.h>
#include "common.cuh"
#include "rqs_device.cuh"

namespace wf {
namespace cf {

constexpr int CF_THREADS = 256;
constexpr int CF_STAGES = 3;

// packed layout per conditioner (floats): head = W1[HALF][HD] | b1[HD] | W2[HD][HD] | b2[HD]   (padded to a multiple of 4)
//                                         then HALF blocks: W3_j[HD][KP3] | b3_j[KP3], KP3 = 3*KP (W | H | D, each padded to KP)
__host__ __device__ constexpr int cf_head_floats(int half, int hd) { return ((half * hd + hd + hd * hd + hd) + 3) & ~3; }
__host__ __device__ constexpr int cf_block_floats(int hd, int kp) { return hd * 3 * kp + 3 * kp; }

struct CfParams {
  const float* weights;
  const float* x;
  float* y;
  float* logdet;
  int64_t N;
  int n_layers, K, inverse;
  float B;
};

__device__ __forceinline__ float4 lds4c(const float* p) { return *reinterpret_cast<const float4*>(p); }

// h_out = tanh(W^T h_in + b): NIN inputs in registers, HD outputs, weights [NIN][HD] in shared memory
template <int NIN, int HD>
__device__ __forceinline__ void dense_tanh(const float* __restrict__ W, const float* __restrict__ b, const float (&in)[NIN],
                                           float (&out)[HD]) {
#pragma unroll
  for (int j0 = 0; j0 < HD; j0 += 4) {
    const float4 bb = lds4c(b + j0);
    float acc[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      const float4 w = lds4c(W + i * HD + j0);
      acc[0] = fmaf(in[i], w.x, acc[0]); acc[1] = fmaf(in[i], w.y, acc[1]);
      acc[2] = fmaf(in[i], w.z, acc[2]); acc[3] = fmaf(in[i], w.w, acc[3]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) out[j0 + t] = fast_tanh(acc[t]);
  }
}

// raw[0..KP) = h . W3[:, c0 .. c0+KP) + b3[c0 ..]  (rolled over blocks of 8 outputs; h statically indexed)
template <int HD, int KP>
__device__ __forceinline__ void dense_out(const float* __restrict__ W3, const float* __restrict__ b3, int c0, int ld,
                                          const float (&h)[HD], float (&raw)[KP]) {
#pragma unroll
  for (int q0 = 0; q0 < KP; q0 += 8) {
    const float4 ba = lds4c(b3 + c0 + q0), bb = lds4c(b3 + c0 + q0 + 4);
    float acc[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int i = 0; i < HD; ++i) {
      const float4 wa = lds4c(W3 + i * ld + c0 + q0), wb = lds4c(W3 + i * ld + c0 + q0 + 4);
      acc[0] = fmaf(h[i], wa.x, acc[0]); acc[1] = fmaf(h[i], wa.y, acc[1]);
      acc[2] = fmaf(h[i], wa.z, acc[2]); acc[3] = fmaf(h[i], wa.w, acc[3]);
      acc[4] = fmaf(h[i], wb.x, acc[4]); acc[5] = fmaf(h[i], wb.y, acc[5]);
      acc[6] = fmaf(h[i], wb.z, acc[6]); acc[7] = fmaf(h[i], wb.w, acc[7]);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) raw[q0 + t] = acc[t];
  }
}

template <int HALF, int HD, int KP, bool FULLK>
__global__ void __launch_bounds__(CF_THREADS) coupling_flow_kernel(const __grid_constant__ CfParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int HEADF = cf_head_floats(HALF, HD);
  constexpr int BLOCKF = cf_block_floats(HD, KP);
  constexpr int NETF = HEADF + HALF * BLOCKF;
  constexpr int SLOTF = HEADF > BLOCKF ? HEADF : BLOCKF;
  constexpr int D = 2 * HALF;
  constexpr int CHUNKS = 1 + HALF;                     // per conditioner
  float* slots = reinterpret_cast<float*>(smem_raw);   // [CF_STAGES][SLOTF]
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + CF_STAGES * SLOTF);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < CF_STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int64_t n_batches = (P.N + CF_THREADS - 1) / CF_THREADS;
  const int64_t my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nets_per_batch = 2 * P.n_layers;
  const int64_t g_total = my_batches * nets_per_batch * CHUNKS;
  // chunk g -> (conditioner index in application order, chunk within the conditioner)
  auto issue = [&](int64_t g) {
    const int c = (int)(g % CHUNKS);
    const int step = (int)((g / CHUNKS) % nets_per_batch);              // half-update index within a batch
    const int layer = P.inverse ? P.n_layers - 1 - step / 2 : step / 2;
    const int which = P.inverse ? 1 - (step & 1) : (step & 1);          // forward: f1 then f2; inverse: f2 then f1
    const float* net = P.weights + (size_t)(layer * 2 + which) * NETF;
    const float* src = c == 0 ? net : net + HEADF + (size_t)(c - 1) * BLOCKF;
    const uint32_t bytes = (uint32_t)((c == 0 ? HEADF : BLOCKF) * sizeof(float));
    const int s = (int)(g % CF_STAGES);
    mbar_expect_tx(&bars[s], bytes);
    bulk_g2s(slots + (size_t)s * SLOTF, src, bytes, &bars[s]);
  };
  int64_t g = 0;
  if (tid == 0)
    for (int k = 0; k < CF_STAGES - 1 && k < g_total; ++k) issue(k);
  auto next_chunk = [&]() -> const float* {
    __syncthreads();                                        // the slot of chunk g-1 is free again
    if (tid == 0 && g + CF_STAGES - 1 < g_total) issue(g + CF_STAGES - 1);
    const int s = (int)(g % CF_STAGES);
    mbar_wait(&bars[s], (uint32_t)((g / CF_STAGES) & 1));
    ++g;
    return slots + (size_t)s * SLOTF;
  };

  const int K = FULLK ? KP : P.K;
  const float B = P.B;
  for (int64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int64_t n_raw = batch * CF_THREADS + tid;
    const bool live = n_raw < P.N;
    const int64_t n = live ? n_raw : P.N - 1;
    float v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = __ldg(P.x + n * D + d);
    float logdet = 0.f;
#pragma unroll 1
    for (int step = 0; step < nets_per_batch; ++step) {
      // forward: step even -> upper <- RQS(upper | f1(lower)), odd -> lower <- RQS(lower | f2(upper))   (:254-272)
      // inverse: step even -> lower <- RQS^-1(lower | f2(upper)), odd -> upper <- RQS^-1(upper | f1(lower)) (:274-292)
      const bool cond_is_lower = P.inverse ? (step & 1) : !(step & 1);
      float cin[HALF];
#pragma unroll
      for (int i = 0; i < HALF; ++i) cin[i] = cond_is_lower ? v[i] : v[HALF + i];
      float h2[HD];
      {
        const float* head = next_chunk();
        const float* W1 = head;
        const float* b1 = W1 + HALF * HD;
        const float* W2 = b1 + HD;
        const float* b2 = W2 + HD * HD;
        float h1[HD];
        dense_tanh<HALF, HD>(W1, b1, cin, h1);
        dense_tanh<HD, HD>(W2, b2, h1, h2);
      }
#pragma unroll 1
      for (int j = 0; j < HALF; ++j) {
        const float* blk = next_chunk();
        const float* W3 = blk;
        const float* b3 = blk + HD * 3 * KP;
        float tj = cond_is_lower ? v[HALF] : v[0];
#pragma unroll
        for (int jj = 1; jj < HALF; ++jj) tj = (j == jj) ? (cond_is_lower ? v[HALF + jj] : v[jj]) : tj;
        float out = tj, lad = 0.f;
        if (tj >= -B && tj <= B) {
          float a[KP], b[KP];
          dense_out<HD, KP>(W3, b3, 0, 3 * KP, h2, a);
          dense_out<HD, KP>(W3, b3, KP, 3 * KP, h2, b);
          const float mx_a = softmax_2b<KP, FULLK>(a, K, 2.f * B);
          const float mx_b = softmax_2b<KP, FULLK>(b, K, 2.f * B);
          int bin;
          auto dget = [&](int jd) {      // softplus(raw derivative jd): one dot product with a per-thread column
            float acc = b3[2 * KP + jd];
#pragma unroll
            for (int i = 0; i < HD; ++i) acc = fmaf(h2[i], W3[i * 3 * KP + 2 * KP + jd], acc);
            return softplus_f(acc);
          };
          rqs_eval<KP, FULLK>(tj, a, b, K, B, P.inverse != 0, mx_a, mx_b, dget, out, lad, bin);
        }
        logdet += lad;
#pragma unroll
        for (int jj = 0; jj < HALF; ++jj) {
          if (cond_is_lower) v[HALF + jj] = (j == jj) ? out : v[HALF + jj];
          else v[jj] = (j == jj) ? out : v[jj];
        }
      }
    }
    if (live) {
#pragma unroll
      for (int d = 0; d < D; ++d) P.y[n * D + d] = v[d];
      P.logdet[n] = logdet;
    }
  }
}

template <int HALF, int HD, int KP, bool FULLK>
int launch_cf(const CfParams& P, cudaStream_t s) {
  constexpr int HEADF = cf_head_floats(HALF, HD);
  constexpr int BLOCKF = cf_block_floats(HD, KP);
  constexpr int SLOTF = HEADF > BLOCKF ? HEADF : BLOCKF;
  const size_t smem = (size_t)CF_STAGES * SLOTF * sizeof(float) + CF_STAGES * sizeof(uint64_t);
  if (smem > 227 * 1024) return WF_ERR_UNSUPPORTED;
  const int64_t n_batches = (P.N + CF_THREADS - 1) / CF_THREADS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  const int blocks = (int)(n_batches < cap ? n_batches : cap);
  WF_CUDA(cudaFuncSetAttribute(coupling_flow_kernel<HALF, HD, KP, FULLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coupling_flow_kernel<HALF, HD, KP, FULLK><<<blocks, CF_THREADS, smem, s>>>(P);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

// One translation unit per instantiation group, compiled in parallel: the fully unrolled width-64 conditioner products cost
// the compiler front end ~45 s per kernel.  K == KP (the common cases K = 32 and K = 8) runs the FULLK instantiation with
// every per-bin bound check folded away.
#define WF_DECL_CF(H, W, KP, F) int launch_cf_h##H##_w##W##_k##KP##_f##F(const CfParams& P, cudaStream_t s);
#define WF_DECL_CF_ALL(H, W) WF_DECL_CF(H, W, 8, 0) WF_DECL_CF(H, W, 8, 1) WF_DECL_CF(H, W, 32, 0) WF_DECL_CF(H, W, 32, 1)
WF_DECL_CF_ALL(1, 8) WF_DECL_CF_ALL(1, 64) WF_DECL_CF_ALL(2, 8) WF_DECL_CF_ALL(2, 64) WF_DECL_CF_ALL(4, 8) WF_DECL_CF_ALL(4, 64)
#undef WF_DECL_CF_ALL
#undef WF_DECL_CF
#define WF_DEF_CF(H, W, KP, F) \
  int launch_cf_h##H##_w##W##_k##KP##_f##F(const CfParams& P, cudaStream_t s) { return launch_cf<H, W, KP, F != 0>(P, s); }

}  // namespace cf
}  // namespace wf
