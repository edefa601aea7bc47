// C-ABI entry points of the fused live path (wf_live_forward, wf_local_energy).
#include <stdlib.h>
#include <string.h>
#include "live_kernel.cuh"
#include "live_inverse.cuh"
#include "live_tc.cuh"

using namespace wf;

extern "C" int64_t wf_live_net_floats(int D) { return (D >= 2 && D <= WF_MAX_D) ? (int64_t)net_floats(D) : -1; }

namespace {

// remove_bias applied to a vector of ones gives the per-coefficient scale (sequential, in place: isplines_jax.py:196-199,
// msplines_jax.py:186-189); the {0: 0} / {0: 1} boundary constraints zero the end coefficients (isplines_jax.py:160-176).
void coeff_weights(int kind, int k, int P, int bc_bits, bool with_bias, float* w) {
  for (int q = 0; q < WF_MAX_P; ++q) w[q] = q < P ? 1.f : 0.f;
  if (with_bias) {
    for (int i = 0; i < k; ++i) {
      const int a = kind == WF_KIND_I ? i + 1 : i;
      const int b = kind == WF_KIND_I ? P - (i + 2) : P - (i + 1);
      if (a >= 0 && a < P) w[a] = w[a] * (float)(i + 1) / (float)k;
      if (b >= 0 && b < P) w[b] = w[b] * (float)(i + 1) / (float)k;
    }
  }
  if (bc_bits & 1) w[0] = 0.f;
  if (bc_bits & 2) w[P - 1] = 0.f;
}

int fill_params(const wf_live_model* m, const wf_live_tables* t, const float* weights, const float* x, int64_t N,
                LiveParams& P) {
  if (!m || !t || !t->dense_I || !x || N < 0) return WF_ERR_INVALID_ARG;
  if (m->D < 2 || m->D > WF_MAX_D || m->n_layers < 0 || m->n_layers > WF_MAX_LAYERS || m->T < 2) return WF_ERR_INVALID_ARG;
  if (m->P_I < 2 || m->P_I > WF_MAX_P || m->k_I < 0) return WF_ERR_INVALID_ARG;
  if (m->D > 4) return WF_ERR_UNSUPPORTED;
  if (m->weight_layout != WF_WEIGHTS_SIMT && m->weight_layout != WF_WEIGHTS_TC) return WF_ERR_INVALID_ARG;
  const bool pnet = m->prior_kind == WF_KIND_B || m->prior_kind == WF_KIND_M;
  if (!weights && (m->n_layers > 0 || pnet)) return WF_ERR_INVALID_ARG;
  if (m->n_layers > 0 && (!t->rec_I || !t->lo_I)) return WF_ERR_INVALID_ARG;
  if (m->prior_kind != -1 && !pnet) return WF_ERR_INVALID_ARG;
  if (pnet && (!t->dense_P || m->P_P < 2 || m->P_P > WF_MAX_P)) return WF_ERR_INVALID_ARG;
  if (m->prior_kind == WF_KIND_B && !t->ob_to_b) return WF_ERR_INVALID_ARG;
  if ((m->bc_P & 4) && (m->prior_kind != WF_KIND_B || m->P_P > WF_MAX_P - 1)) return WF_ERR_INVALID_ARG;   // folded prior layer
  if (m->prior_kind == WF_KIND_M && (!t->rec_P || !t->lo_P)) return WF_ERR_INVALID_ARG;
  if (m->has_box && !(m->box > 0.f)) return WF_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(weights) & 15) || (reinterpret_cast<uintptr_t>(t->dense_I) & 15) ||
      (t->dense_P && (reinterpret_cast<uintptr_t>(t->dense_P) & 15)))
    return WF_ERR_INVALID_ARG;
  memset(&P, 0, sizeof(P));
  P.m = *m;
  P.weights = weights; P.x = x; P.N = N;
  P.tab_I = t->dense_I; P.rec_I = t->rec_I; P.lo_I = t->lo_I;
  P.tab_P = t->dense_P; P.rec_P = t->rec_P; P.lo_P = t->lo_P; P.ob_to_b = t->ob_to_b;
  P.n_nets = m->n_layers + (pnet ? 1 : 0);
  coeff_weights(WF_KIND_I, m->k_I, m->P_I, m->bc_I, true, P.wq_I);
  if (pnet) coeff_weights(m->prior_kind, m->k_P, m->P_P, m->bc_P, m->prior_kind == WF_KIND_M, P.wq_P);
  P.rec_I_t = t->rec_I_t; P.rec_P_t = t->rec_P_t;
  for (int q = 0; q < m->P_I; ++q) P.wsum_I += P.wq_I[q];
  for (int q = 0; q < WF_MAX_P; ++q) P.cwq_I[q + 1] = P.cwq_I[q] + P.wq_I[q];
  if (pnet) for (int q = 0; q < m->P_P; ++q) P.wsum_P += P.wq_P[q];
  if (m->weight_layout == WF_WEIGHTS_TC) {
    if (m->n_layers > 0 && !t->rec_I_t) return WF_ERR_INVALID_ARG;
    if (m->prior_kind == WF_KIND_M && !t->rec_P_t) return WF_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(t->rec_I_t) & 15) || (reinterpret_cast<uintptr_t>(t->rec_P_t) & 15)) return WF_ERR_INVALID_ARG;
  }
  return WF_OK;
}

int dispatch_tc(LiveParams& P, bool lap, const ltc::TcExtra& X, cudaStream_t s) {
  // tensor-core image: B prior only with the pre-multiplied third layer (bit 2 of bc_P)
  if (P.m.prior_kind == WF_KIND_B && !(P.m.bc_P & 4)) return WF_ERR_INVALID_ARG;
  switch (P.m.D) {
    case 2: return lap ? launch_live_tc_d2_lap1(P, X, s) : launch_live_tc_d2_lap0(P, X, s);
    case 3: return lap ? launch_live_tc_d3_lap1(P, X, s) : launch_live_tc_d3_lap0(P, X, s);
    case 4: return lap ? launch_live_tc_d4_lap1(P, X, s) : launch_live_tc_d4_lap0(P, X, s);
    default: return WF_ERR_UNSUPPORTED;
  }
}

int dispatch(LiveParams& P, bool lap, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (P.m.weight_layout == WF_WEIGHTS_TC) {
    ltc::TcExtra X;
    memset(&X, 0, sizeof(X));
    return dispatch_tc(P, lap, X, s);
  }
  switch (P.m.D) {
    case 2: return lap ? launch_live_d2_lap1(P, s) : launch_live_d2_lap0(P, s);
    case 3: return lap ? launch_live_d3_lap1(P, s) : launch_live_d3_lap0(P, s);
    case 4: return lap ? launch_live_d4_lap1(P, s) : launch_live_d4_lap0(P, s);
    default: return WF_ERR_UNSUPPORTED;
  }
}

}  // namespace

extern "C" int wf_live_forward(const wf_live_model* model, const wf_live_tables* tables, const float* weights,
                               const float* x, int64_t N, float* u, float* logdet, float* logpdf, float* psi,
                               void* stream) {
  if (N == 0) return WF_OK;   // empty batch: nothing to do (pointers of empty buffers may be NULL)
  LiveParams P;
  const int st = fill_params(model, tables, weights, x, N, P);
  if (st != WF_OK) return st;
  if (psi && model->prior_kind != WF_KIND_B) return WF_ERR_INVALID_ARG;
  if (N == 0) return WF_OK;
  P.u = u; P.logdet = logdet; P.logpdf = logpdf; P.psi = psi;
  return dispatch(P, false, stream);
}

extern "C" int wf_local_energy(const wf_live_model* model, const wf_live_tables* tables, const float* weights,
                               const float* protons, int n_protons, const float* x, int64_t N, float* psi, float* hpsi,
                               float* eloc, float* grad, float* lap, double* sums, void* stream) {
  if (N == 0) return WF_OK;   // empty batch: nothing to do (pointers of empty buffers may be NULL)
  LiveParams P;
  const int st = fill_params(model, tables, weights, x, N, P);
  if (st != WF_OK) return st;
  if (model->prior_kind != WF_KIND_B) return WF_ERR_INVALID_ARG;
  if (n_protons < 0 || n_protons > WF_MAX_D || (n_protons > 0 && !protons)) return WF_ERR_INVALID_ARG;
  if (N == 0) return WF_OK;
  P.psi = psi; P.hpsi = hpsi; P.eloc = eloc; P.grad = grad; P.lap = lap; P.sums = sums;
  P.n_protons = n_protons;
  for (int i = 0; i < n_protons; ++i) P.protons[i] = protons[i];   // HOST array (a handful of floats)
  return dispatch(P, true, stream);
}

namespace {
int fill_inverse(const wf_live_model* m, const wf_live_tables* t, const float* weights, int64_t N, bool sample, InvParams& P) {
  LiveParams L;
  // same validation as the forward path (x is checked by the callers)
  const float dummy = 0.f;
  const int st = fill_params(m, t, weights, &dummy, N, L);
  if (st != WF_OK) return st;
  if (m->n_layers > 0 && !(m->tol > 0.f)) return WF_ERR_INVALID_ARG;
  if (m->bc_P & 4) return WF_ERR_INVALID_ARG;     // the sampler / inverse need the raw conditioner outputs (unfolded weights)
  if (m->weight_layout != WF_WEIGHTS_SIMT) return WF_ERR_INVALID_ARG;
  const bool pnet = m->prior_kind == WF_KIND_B || m->prior_kind == WF_KIND_M;
  if (sample && !pnet) return WF_ERR_INVALID_ARG;
  if (sample && m->prior_kind == WF_KIND_B && !t->b_to_ob) return WF_ERR_INVALID_ARG;
  if (sample && m->prior_kind == WF_KIND_M && m->n_knots_P <= 0) return WF_ERR_INVALID_ARG;
  memset(&P, 0, sizeof(P));
  P.m = *m; P.weights = weights; P.N = N;
  P.rec_I = t->rec_I; P.lo_I = t->lo_I; P.tab_I = t->dense_I; P.tab_P = t->dense_P;
  P.ob_to_b = t->ob_to_b; P.b_to_ob = t->b_to_ob;
  P.n_knots_P = m->n_knots_P;
  for (int q = 0; q < WF_MAX_P; ++q) { P.wq_I[q] = L.wq_I[q]; P.wq_P[q] = L.wq_P[q]; }
  return WF_OK;
}
int dispatch_inverse(InvParams& P, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  switch (P.m.D) {
    case 2: return launch_inverse_d2(P, s);
    case 3: return launch_inverse_d3(P, s);
    case 4: return launch_inverse_d4(P, s);
    default: return WF_ERR_UNSUPPORTED;
  }
}
}  // namespace

extern "C" int wf_live_inverse(const wf_live_model* model, const wf_live_tables* tables, const float* weights, const float* u,
                               int64_t N, int exact, float* x, void* stream) {
  if (N == 0) return WF_OK;
  if (!u || !x) return WF_ERR_INVALID_ARG;
  InvParams P;
  const int st = fill_inverse(model, tables, weights, N, false, P);
  if (st != WF_OK) return st;
  P.u_in = u; P.x_out = x; P.exact = exact ? 1 : 0; P.do_sample = 0;
  return dispatch_inverse(P, stream);
}

extern "C" int wf_live_sample(const wf_live_model* model, const wf_live_tables* tables, const float* weights, uint64_t seed,
                              int64_t N, int exact, float* x, float* u_out, void* stream) {
  if (N == 0) return WF_OK;
  if (!x) return WF_ERR_INVALID_ARG;
  InvParams P;
  const int st = fill_inverse(model, tables, weights, N, true, P);
  if (st != WF_OK) return st;
  P.x_out = x; P.u_out = u_out; P.seed = seed; P.exact = exact ? 1 : 0; P.do_sample = 1;
  return dispatch_inverse(P, stream);
}

// ---------------------------------------------------------------------------------------------- local energy + estimator exchange
namespace {
__global__ void __launch_bounds__(32) p2p_after_kernel(p2p::Args a, const double* __restrict__ local) {
  __shared__ double vals[p2p::MAX_WORLD * 4];
  p2p::allreduce_warp(a, local, vals, p2p::TIMEOUT_CYCLES);
}
}  // namespace

extern "C" int wf_local_energy_exchange(const wf_live_model* model, const wf_live_tables* tables, const float* weights,
                                        const float* protons, int n_protons, const float* x, int64_t N, float* psi, float* hpsi,
                                        float* eloc, float* grad, float* lap, double* sums, const uint64_t* peer_bufs_dev, int rank,
                                        int world, uint64_t step, double* sums_out, uint32_t* done_counter, void* stream) {
  if (!sums || !sums_out || !peer_bufs_dev || !done_counter || world < 1 || world > p2p::MAX_WORLD || rank < 0 || rank >= world || step == 0)
    return WF_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  p2p::Args a{reinterpret_cast<const unsigned long long*>(peer_bufs_dev), rank, world, (unsigned long long)step, sums_out};
  if (N > 0 && model && model->weight_layout == WF_WEIGHTS_TC) {
    LiveParams P;
    const int st = fill_params(model, tables, weights, x, N, P);
    if (st != WF_OK) return st;
    if (model->prior_kind != WF_KIND_B) return WF_ERR_INVALID_ARG;
    if (n_protons < 0 || n_protons > WF_MAX_D || (n_protons > 0 && !protons)) return WF_ERR_INVALID_ARG;
    P.psi = psi; P.hpsi = hpsi; P.eloc = eloc; P.grad = grad; P.lap = lap; P.sums = sums;
    P.n_protons = n_protons;
    for (int i = 0; i < n_protons; ++i) P.protons[i] = protons[i];
    ltc::TcExtra X;
    X.xchg = a; X.done_counter = done_counter;
    return dispatch_tc(P, true, X, s);        // the last CTA to retire runs the exchange: one launch per step
  }
  // CUDA-core kernel (or an empty shard): the exchange follows as its own 32-thread launch
  const int st = wf_local_energy(model, tables, weights, protons, n_protons, x, N, psi, hpsi, eloc, grad, lap, sums, stream);
  if (st != WF_OK) return st;
  p2p_after_kernel<<<1, 32, 0, s>>>(a, sums);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

// ---------------------------------------------------------------------------------------------- tensor-core weight image
namespace {
// one thread per float of the destination image (layout: live_tc.cuh, net_floats_tc)
__global__ void pack_tc_kernel(const float* __restrict__ src, float* __restrict__ dst, int D, int n_nets) {
  const int netf_s = net_floats(D), netf_t = ltc::net_floats_tc(D);
  const int64_t total = (int64_t)n_nets * netf_t;
  const int N3 = D * WF_MAX_P;
  const int w2p = WF_HIDDEN * WF_HIDDEN, w3p = N3 * WF_HIDDEN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int net = (int)(i / netf_t);
    int e = (int)(i % netf_t);
    const float* sn = src + (size_t)net * netf_s;
    const float* W1 = sn;
    const float* b1 = W1 + D * WF_HIDDEN;
    const float* W2 = b1 + WF_HIDDEN;
    const float* b2 = W2 + WF_HIDDEN * WF_HIDDEN;
    const float* W3 = b2 + WF_HIDDEN;
    const float* b3 = W3 + WF_HIDDEN * N3;
    float v;
    if (e < 2 * w2p + 2 * w3p) {
      // plane element: [k-block][row n][8 chunks of 4 floats], 16-byte chunk c stored at position c ^ (n % 8)
      const bool third = e >= 2 * w2p;
      if (third) e -= 2 * w2p;
      const int plane = third ? w3p : w2p, rows = third ? N3 : WF_HIDDEN;
      const bool lo = e >= plane;
      if (lo) e -= plane;
      const int kb = e / (rows * 32), r = (e / 32) % rows, pos = (e % 32) / 4, el = e % 4;
      const int k = kb * 32 + ((pos ^ (r & 7)) * 4) + el;
      const float w = third ? W3[(size_t)k * N3 + r] : W2[(size_t)k * WF_HIDDEN + r];
      const float hi = ltc::tf32_rn(w);
      v = lo ? ltc::tf32_rn(w - hi) : hi;
    } else {
      e -= 2 * w2p + 2 * w3p;
      if (e < D * WF_HIDDEN) v = W1[e];
      else if ((e -= D * WF_HIDDEN) < WF_HIDDEN) v = b1[e];
      else if ((e -= WF_HIDDEN) < WF_HIDDEN) v = b2[e];
      else v = b3[e - WF_HIDDEN];
    }
    dst[i] = v;
  }
}
}  // namespace

extern "C" int64_t wf_live_net_floats_tc(int D) { return (D >= 2 && D <= 4) ? (int64_t)ltc::net_floats_tc(D) : -1; }

extern "C" int wf_live_pack_tc(int D, int n_nets, const float* weights, float* weights_tc, void* stream) {
  if (D < 2 || D > 4) return WF_ERR_UNSUPPORTED;
  if (n_nets < 0 || (n_nets > 0 && (!weights || !weights_tc))) return WF_ERR_INVALID_ARG;
  if (n_nets == 0) return WF_OK;
  if (reinterpret_cast<uintptr_t>(weights_tc) & 15) return WF_ERR_INVALID_ARG;
  const int64_t total = (int64_t)n_nets * ltc::net_floats_tc(D);
  pack_tc_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weights, weights_tc, D, n_nets);
  WF_LAUNCH_CHECK();
  return WF_OK;
}
