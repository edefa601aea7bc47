// Inverse of the live flow and the prior rejection sampler (one thread per sample).
//
// Reference: IMADE.inverse_fun (flows/bijections/made.py:85-100) with helpers.binary_search (utils/helpers.py:150-166),
// BoxTransformLayer.reverse_fun_{mean,first} (made.py:139-154,186-197), Serial.inverse_fun (bijections.py:462-463),
// Waveflow.sample / MFlow.sample (wavefunctions.py:74-107, flows/distributions.py:165-190) and the rejection samplers
// of bsplines_jax.py:144-171 / msplines_jax.py:129-154.
#pragma once
#include "live_device.cuh"

namespace wf {

struct InvParams {
  wf_live_model m;
  const float* weights;      // flow nets 0..L-1, then the prior net
  const float* rec_I; const int32_t* lo_I; const float* tab_I;
  const float* tab_P;        // dense prior tables (OB for the B prior, M tables for the M prior)
  const float* ob_to_b; const float* b_to_ob;
  const float* u_in;         // [N][D] prior-space points to invert (inverse entry point), or NULL when sampling
  float* x_out;              // [N][D]
  float* u_out;              // [N][D] (sampler, nullable)
  int64_t N;
  uint64_t seed;
  int exact;                 // 0: reference semantics (quirks Q1/Q2), 1: true inverse
  int do_sample;
  int n_knots_P;
  float wq_I[WF_MAX_P];
  float wq_P[WF_MAX_P];
};

// I-spline coefficients of one dimension from the conditioner output parked in S[0..31]:
//   c_q = w_q (s_q / S + reg) / Z   (made.py:67-72; see sigmoid_spline)  -> S[q];  exclusive prefix sums -> S[32 + q]
__device__ __forceinline__ void imade_coefficients(const Scratch& S, int P, const float* __restrict__ wq, float reg) {
  float ssum = 0.f;
  for (int q = 0; q < P; ++q) { const float s = fast_sigmoid(S[q]); S[q] = s; ssum += s; }
  float z = 0.f;
  for (int q = 0; q < P; ++q) { const float t = wq[q] * (S[q] / ssum + reg); S[q] = t; z += t; }
  float cum = 0.f;
  for (int q = 0; q < P; ++q) { const float c = S[q] / z; S[q] = c; S[WF_MAX_P + q] = cum; cum += c; }
}

// sum_q c_q I_q(x) from the compact records (same summation order as the dense sum: prefix, then the window)
__device__ __forceinline__ float ispline_value(const Scratch& S, int P, const float* __restrict__ rec, const int32_t* __restrict__ lo,
                                               const float* __restrict__ dense, int T, float x) {
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(x, T);
  const int lo_l = __ldg(lo + n.l), lo_r = __ldg(lo + n.r);
  const int sh = lo_r - lo_l;
  if (sh < 0 || sh > 1) {
    float a = 0.f;
    for (int q = 0; q < P; ++q)
      a = fmaf(S[q], lerp_tab(__ldg(dense + (size_t)n.l * 4 * WF_MAX_P + q), __ldg(dense + (size_t)n.r * 4 * WF_MAX_P + q), np_, n.dx), a);
    return a;
  }
  float a = S[WF_MAX_P + (lo_l < P ? lo_l : P - 1)];
#pragma unroll
  for (int t = 0; t < WF_WIN; ++t) {
    const int q = lo_l + t;
    if (q < P) {
      const int tr = t - sh;
      const float yl = __ldg(rec + (size_t)n.l * 4 * WF_WIN + t);
      const float yrr = __ldg(rec + (size_t)n.r * 4 * WF_WIN + (tr < 0 ? 0 : tr));
      a = fmaf(S[q], lerp_tab(yl, tr < 0 ? 1.f : yrr, np_, n.dx), a);
    }
  }
  return a;
}

// helpers.binary_search on f(x) = ispline(x) - y: returns the LOWER bracket
__device__ __forceinline__ float bisect(const Scratch& S, int P, const float* rec, const int32_t* lo, const float* dense, int T,
                                        float y, float tol) {
  float a = 0.f, b = 1.f;
  const float half_tol = tol / 2.f;
  for (int it = 0; it < 64; ++it) {
    const float mid = 0.5f * (a + b);
    if (!((a + half_tol < mid) && (mid < b - half_tol))) break;
    const bool upper = (ispline_value(S, P, rec, lo, dense, T, mid) - y) > 0.f;
    a = upper ? a : mid;
    b = upper ? mid : b;
  }
  return a;
}

// One IMADE layer inverted in place (v holds the layer OUTPUT on entry, the layer INPUT on return).
template <int D>
__device__ __forceinline__ void imade_inverse(const Ctx<D, false>& cx, const InvParams& P, const float* __restrict__ net,
                                              const Scratch& S, float (&v)[D]) {
  const wf_live_model& M = P.m;
  float h[WF_HIDDEN];
  float out[D];
#pragma unroll
  for (int d = 0; d < D; ++d) out[d] = 0.f;
  if (!P.exact) mlp_hidden<D, false>(cx, net, v, S, h);          // conditioned on the layer's inputs y (quirk Q1)
#pragma unroll 1
  for (int d = 0; d < D; ++d) {
    if (P.exact) mlp_hidden<D, false>(cx, net, out, S, h);       // conditioned on the already inverted prefix
    mlp_out<D, false>(cx, net, d, h, S);
    imade_coefficients(S, M.P_I, P.wq_I, M.reg);
    float yd = v[0];
#pragma unroll
    for (int dd = 1; dd < D; ++dd) yd = (d == dd) ? v[dd] : yd;
    const float xd = bisect(S, M.P_I, P.rec_I, P.lo_I, P.tab_I, M.T, yd, M.tol);
#pragma unroll
    for (int dd = 0; dd < D; ++dd) out[dd] = (d == dd) ? xd : out[dd];
  }
#pragma unroll
  for (int d = 0; d < D; ++d) v[d] = out[d];
}

template <int D>
__device__ __forceinline__ void box_inverse(const InvParams& P, float (&u)[D]) {
  const wf_live_model& M = P.m;
  const float L = M.box;
  float x[D];
  if (M.coord_mean) {
    if (!P.exact) {
      // made.py:186-197 as written (exact only for D = 2, quirk Q2)
      float pos = 0.f, sum = 0.f;
      x[0] = 0.f;
#pragma unroll
      for (int i = 1; i < D; ++i) { pos += u[i - 1]; x[i] = pos; sum += pos; }
      const float mean = sum / (float)D;
      const float w = x[D - 1];
      const float pm = u[D - 1] * (1.f - w) - (0.5f - mean);
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = (x[i] - mean + pm) * 2.f * L;
    } else {
      // true inverse of direct_fun_mean (made.py:156-183)
      float diff[D];
      float space = 2.f * L, w = 0.f;
#pragma unroll
      for (int i = 0; i < D - 1; ++i) { diff[i] = u[i] * (space + 1e-7f); space -= diff[i]; w += diff[i]; }
      x[0] = u[D - 1] * (2.f * L - w + 1e-7f) - L;
#pragma unroll
      for (int i = 1; i < D; ++i) x[i] = x[i - 1] + diff[i - 1];
    }
  } else {
    x[0] = (u[0] - 0.5f) * (2.f * L);
#pragma unroll
    for (int i = 1; i < D; ++i) x[i] = P.exact ? u[i] * (L - x[i - 1] + 1e-7f) + x[i - 1] : u[i] * (L - x[i - 1]) + x[i - 1];
  }
#pragma unroll
  for (int i = 0; i < D; ++i) u[i] = x[i];
}

// One column of the prior: rejection sampling of dimension `col` given the conditioner output in S[0..31].
template <int D>
__device__ __forceinline__ float sample_prior_column(const InvParams& P, const Scratch& S, const float* __restrict__ ob_s,
                                                     const float* __restrict__ bo_s, int col, int64_t sample) {
  const wf_live_model& M = P.m;
  const int PP = M.P_P, T = M.T;
  const float np_ = (float)(T - 1);
  float ymax;
  if (M.prior_kind == WF_KIND_B) {
    // w = o / sum o; ends masked; / ||w||; c = w @ ob_to_b; / ||c||   (wavefunctions.py:91-96, bsplines_jax.py:163-165)
    float osum = 0.f, n2 = 0.f;
    for (int q = 0; q < PP; ++q) osum += S[q];
    for (int q = 0; q < WF_MAX_P; ++q) { const float w = S[q] / osum * P.wq_P[q]; S[q] = w; n2 = fmaf(w, w, n2); }
    const float inv = 1.f / sqrtf(n2);
    float c2 = 0.f;
    for (int j = 0; j < PP; ++j) {
      float c = 0.f;
      for (int i = 0; i < PP; ++i) c = fmaf(S[i] * inv, ob_s[i * WF_MAX_P + j], c);
      S[WF_MAX_P + j] = c; c2 = fmaf(c, c, c2);
    }
    const float invc = 1.f / sqrtf(c2);
    ymax = 0.f;
    for (int j = 0; j < PP; ++j) S[WF_MAX_P + j] *= invc;
    for (int j = 0; j < PP; ++j) {            // ymax = max((c @ b_to_ob)^2): convex-hull bound in the local B basis
      float b = 0.f;
      for (int i = 0; i < PP; ++i) b = fmaf(S[WF_MAX_P + i], bo_s[i * WF_MAX_P + j], b);
      ymax = fmaxf(ymax, b * b);
    }
  } else {
    // sigmoid; / sum; remove_bias; boundary mask; / sum    (distributions.py:172-176), ymax = max(w) * n_knots
    float ssum = 0.f, z = 0.f;
    for (int q = 0; q < PP; ++q) { const float s = fast_sigmoid(S[q]); S[q] = s; ssum += s; }
    for (int q = 0; q < PP; ++q) { const float t = P.wq_P[q] * (S[q] / ssum); S[q] = t; z += t; }
    ymax = 0.f;
    for (int q = 0; q < PP; ++q) { const float c = S[q] / z; S[WF_MAX_P + q] = c; ymax = fmaxf(ymax, c); }
    ymax *= (float)P.n_knots_P;
  }
  float x = 0.f;
  for (uint32_t attempt = 0; attempt < 100000u; ++attempt) {
    float r0, r1;
    philox_uniform2(P.seed, (uint64_t)sample, (uint32_t)col, attempt, r0, r1);
    x = r0;
    const float y = r1 * ymax;
    const NodeIdx n = node_index(x, T);
    float f = 0.f;
    for (int j = 0; j < PP; ++j)
      f = fmaf(S[WF_MAX_P + j], lerp_tab(__ldg(P.tab_P + (size_t)n.l * 4 * WF_MAX_P + j), __ldg(P.tab_P + (size_t)n.r * 4 * WF_MAX_P + j), np_, n.dx), f);
    const float dens = M.prior_kind == WF_KIND_B ? f * f : f;
    if (y < dens) break;
  }
  return x;
}

// Shared memory: 2 x net ring | ob_to_b [32][32] | b_to_ob [32][32] | scratch [64][LIVE_THREADS] | 2 mbarriers
struct InvSmem {
  static __host__ __device__ size_t total(int D) {
    return 2 * (size_t)net_floats(D) * sizeof(float) + 2 * (size_t)WF_MAX_P * WF_MAX_P * sizeof(float) +
           (size_t)LIVE_SCRATCH * LIVE_THREADS * sizeof(float) + 2 * sizeof(uint64_t);
  }
};

template <int D>
__global__ void __launch_bounds__(LIVE_THREADS, 1) live_inverse_kernel(const __grid_constant__ InvParams P) {
  using C = Ctx<D, false>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NETF = net_floats(D);
  const wf_live_model& M = P.m;
  const int tid = threadIdx.x;
  float* nets_s = reinterpret_cast<float*>(smem_raw);
  float* ob_s = nets_s + 2 * NETF;
  float* bo_s = ob_s + WF_MAX_P * WF_MAX_P;
  float* scratch = bo_s + WF_MAX_P * WF_MAX_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + LIVE_SCRATCH * LIVE_THREADS);
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  if (P.do_sample && M.prior_kind == WF_KIND_B) {
    for (int i = tid; i < WF_MAX_P * WF_MAX_P; i += LIVE_THREADS) {
      const int r = i / WF_MAX_P, c = i % WF_MAX_P;
      const bool in = r < M.P_P && c < M.P_P;
      ob_s[i] = in ? P.ob_to_b[r * M.P_P + c] : 0.f;
      bo_s[i] = in ? P.b_to_ob[r * M.P_P + c] : 0.f;
    }
  }
  __syncthreads();
  C cx; cx.init(tid & 31);
  const Scratch S{scratch + tid};
  const int64_t n_batches = (P.N + LIVE_THREADS - 1) / LIVE_THREADS;
  // net sequence per batch: [prior (sampling only)], flow nets L-1 .. 0
  const int per_batch = M.n_layers + (P.do_sample ? 1 : 0);
  const uint32_t net_bytes = (uint32_t)(NETF * sizeof(float));
  auto net_index = [&](int64_t g) { const int k = (int)(g % per_batch); return P.do_sample ? (k == 0 ? M.n_layers : M.n_layers - k) : M.n_layers - 1 - k; };
  auto issue_net = [&](int64_t g) {
    const int b = (int)(g & 1);
    mbar_expect_tx(&bars[b], net_bytes);
    bulk_g2s(nets_s + (size_t)b * NETF, P.weights + (size_t)net_index(g) * NETF, net_bytes, &bars[b]);
  };
  const int64_t my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t g_total = my_batches * per_batch;
  int64_t g = 0;
  if (tid == 0 && g_total > 0) issue_net(0);
  auto next_net = [&]() -> const float* {
    __syncthreads();
    if (tid == 0 && g + 1 < g_total) issue_net(g + 1);
    mbar_wait(&bars[g & 1], (uint32_t)((g >> 1) & 1));
    const float* net = nets_s + (size_t)(g & 1) * NETF;
    ++g;
    return net;
  };

  for (int64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int64_t s_raw = batch * LIVE_THREADS + tid;
    const bool live = s_raw < P.N;
    const int64_t s = live ? s_raw : P.N - 1;
    float v[D];
    if (P.do_sample) {
      // autoregressive prior sampling (wavefunctions.py:88-101 / distributions.py:171-185)
      const float* net = next_net();
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = 0.f;
#pragma unroll 1
      for (int col = 0; col < D; ++col) {
        float h[WF_HIDDEN];
        mlp_hidden<D, false>(cx, net, v, S, h);
        mlp_out<D, false>(cx, net, col, h, S);
        const float xs = sample_prior_column<D>(P, S, ob_s, bo_s, col, s);
#pragma unroll
        for (int dd = 0; dd < D; ++dd) v[dd] = (col == dd) ? xs : v[dd];
      }
      if (live && P.u_out) {
#pragma unroll
        for (int d = 0; d < D; ++d) P.u_out[s * D + d] = v[d];
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = __ldg(P.u_in + s * D + d);
    }
    // Serial.inverse_fun: layers in reverse order, each (Reverse, IMADE) pair inverted as flip, then IMADE.inverse
#pragma unroll 1
    for (int layer = M.n_layers - 1; layer >= 0; --layer) {
      const float* net = next_net();
      float t[D];
#pragma unroll
      for (int d = 0; d < D; ++d) t[d] = v[D - 1 - d];
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = t[d];
      imade_inverse<D>(cx, P, net, S, v);
    }
    if (M.has_box) box_inverse<D>(P, v);
    if (live) {
#pragma unroll
      for (int d = 0; d < D; ++d) P.x_out[s * D + d] = v[d];
    }
  }
}

template <int D>
int launch_inverse(InvParams& P, cudaStream_t s) {
  const size_t smem = InvSmem::total(D);
  if (smem > 227 * 1024) return WF_ERR_UNSUPPORTED;
  const int64_t n_batches = (P.N + LIVE_THREADS - 1) / LIVE_THREADS;
  const int blocks = (int)(n_batches < num_sms() ? n_batches : num_sms());
  WF_CUDA(cudaFuncSetAttribute(live_inverse_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  live_inverse_kernel<D><<<blocks, LIVE_THREADS, smem, s>>>(P);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

int launch_inverse_d2(InvParams& P, cudaStream_t s);
int launch_inverse_d3(InvParams& P, cudaStream_t s);
int launch_inverse_d4(InvParams& P, cudaStream_t s);

}  // namespace wf
