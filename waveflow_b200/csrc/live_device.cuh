// Fused live path: BoxTransformLayer -> (IMADE, Reverse) x L -> B-/M-spline prior, with an optional forward-mode
// Laplacian carried through the very same code (template flag LAP).
//
// Thread mapping ("lane per component"): with LAP a walker occupies G = D + 2 consecutive lanes of a warp that hold,
// for every intermediate scalar s of the forward pass, the components
//     lane 0      : s                     (value)
//     lane 1..D   : ds/dx_j               (gradient w.r.t. the D walker coordinates)
//     lane D+1    : sum_j d2s/dx_j^2      (Laplacian)
// Linear layers act identically on every component (the bias only on the value lane), so the conditioner MLPs are plain
// SIMT code; nonlinearities mix components through warp shuffles.  Inside the per-dimension "glue" (sigmoid, the
// normalisations, the spline sums) gradient lanes additionally carry a pending second-derivative term `p`
// (d2s/dx_j^2 contributions not yet summed into the Laplacian lane) so that no cross-lane reduction is needed per
// coefficient; it is folded into the Laplacian lane once per dimension.
// Without LAP, G = 1 and every helper collapses to scalar code: one thread per sample.
//
// Code shape: the instruction cache, not the FMA pipe, limited the first (fully unrolled, ~0.5 MB of SASS) version of
// this kernel (ncu: 3.6 stall cycles per issue on "no instruction").  Hence: input activations of a layer live in
// registers (statically indexed, the reduction loop is unrolled), output blocks are produced by ROLLED loops and parked
// in a thread-private shared-memory scratch column, and all conditioner nets (flow layers and prior) run through the
// same rolled loop body.
//
// Reference: flows/bijections/made.py:66-81,108-183, model_factory.py:8-93, splines/isplines_jax.py:45-79,158-202,
// splines/bsplines_jax.py:127-137,173-198, wavefunctions.py:33-71, flows/distributions.py:139-163,
// utils/physics.py:50-93, vqmc.py:198-200.
#pragma once
#include <math.h>
#include "common.cuh"

namespace wf {

constexpr int LIVE_THREADS = 384;          // 12 warps per SM (170 registers per thread)
constexpr int LIVE_SCRATCH = 64;            // floats of thread-private scratch (one column per thread)
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ constexpr int net_floats(int D) {
  return D * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN * D * WF_MAX_P + D * WF_MAX_P;
}

struct LiveParams {
  wf_live_model m;
  const float* weights;   // n_nets conditioners, packed (see wf_live_net_floats)
  const float* tab_I;     // dense [T][4][32] I tables (fallback for arguments outside [0, 1])
  const float* rec_I;     // compact node records [T][4][8]
  const int32_t* lo_I;    // [T]
  const float* tab_P;     // dense [T][4][32]: OB tables (B prior) / M tables (M prior)
  const float* rec_P;     // compact records of the M tables (M prior)
  const int32_t* lo_P;
  const float* ob_to_b;   // [P_P][P_P]  (B prior)
  const float* x;         // [N][D]
  int64_t N;
  float* u; float* logdet; float* logpdf; float* psi;      // forward outputs (nullable)
  float* hpsi; float* eloc; float* grad; float* lap;       // local-energy outputs (nullable)
  double* sums;                                            // {sum E, sum E^2, n, sum psi^2} (nullable)
  const float* rec_I_t;   // compact records transposed to [T][8][4] (tensor-core kernels: one 128-bit load per window slot)
  const float* rec_P_t;
  float wsum_I, wsum_P;   // sum_q wq_I[q], sum_q wq_P[q] (sequential float32 sums)
  float cwq_I[WF_MAX_P + 1];   // cwq_I[j] = sum_{q < j} wq_I[q] (sequential float32 prefix sums)
  float wq_I[WF_MAX_P];   // remove_bias scale x boundary mask of the I-spline coefficients (0 beyond P_I)
  float wq_P[WF_MAX_P];   // same for the prior coefficients (M: remove_bias x mask, B: mask)
  float protons[WF_MAX_D];
  int n_protons;
  int n_nets;
  int warps_per_cta;      // working warps per CTA (<= LIVE_THREADS / 32), chosen by the launcher
};

// m: this lane's component; p: pending second derivative (gradient lanes only); v: the VALUE, known to every lane.
struct J { float m, p, v; };

template <int D, bool LAP>
struct Ctx {
  static constexpr int G = LAP ? D + 2 : 1;
  static constexpr int WPW = 32 / G;     // walkers per warp
  int gbase;                             // lane holding the value component of this walker
  int comp, slot;
  bool is_v, is_g, is_l;

  __device__ __forceinline__ void init(int lane) {
    if constexpr (LAP) {
      slot = lane / G;
      comp = lane - slot * G;
      if (slot >= WPW) { slot = WPW - 1; comp = G - 1; }   // idle tail lanes shadow the last walker's Laplacian lane
      gbase = slot * G;
      is_v = comp == 0; is_g = comp >= 1 && comp <= D; is_l = comp == D + 1;
    } else {
      gbase = lane; slot = lane; comp = 0; is_v = true; is_g = false; is_l = false;
    }
  }
  __device__ __forceinline__ float bv(float a) const {       // broadcast of the value component
    if constexpr (LAP) return __shfl_sync(FULL, a, gbase); else return a;
  }
  __device__ __forceinline__ float gsum(float t) const {     // sum over the gradient lanes (every lane receives it)
    if constexpr (LAP) {
      float s = 0.f;
#pragma unroll
      for (int d = 1; d <= D; ++d) s += __shfl_sync(FULL, t, gbase + d);
      return s;
    } else return 0.f;
  }
  __device__ __forceinline__ float fold(const J& a) const {  // -> 1-register bundle
    if constexpr (LAP) { const float s = gsum(a.p); return is_l ? a.m + s : a.m; } else return a.m;
  }
  __device__ __forceinline__ J constant(float c) const { return J{is_v ? c : 0.f, 0.f, c}; }
  // f(a) given f, f', f'' evaluated at a.v
  __device__ __forceinline__ J unary(const J& a, float f0, float f1, float f2) const {
    if constexpr (LAP) return J{is_v ? f0 : f1 * a.m, is_g ? fmaf(f2 * a.m, a.m, f1 * a.p) : 0.f, f0};
    else return J{f0, 0.f, f0};
  }
  __device__ __forceinline__ J mul(const J& a, const J& b) const {
    const float v = a.v * b.v;
    if constexpr (LAP)
      return J{is_v ? v : fmaf(a.v, b.m, b.v * a.m), is_g ? fmaf(2.f * a.m, b.m, fmaf(a.v, b.p, b.v * a.p)) : 0.f, v};
    else return J{v, 0.f, v};
  }
  __device__ __forceinline__ J add(const J& a, const J& b) const { return J{a.m + b.m, LAP ? a.p + b.p : 0.f, a.v + b.v}; }
  __device__ __forceinline__ J sub(const J& a, const J& b) const { return J{a.m - b.m, LAP ? a.p - b.p : 0.f, a.v - b.v}; }
  __device__ __forceinline__ J scale(const J& a, float s) const { return J{a.m * s, LAP ? a.p * s : 0.f, a.v * s}; }
  __device__ __forceinline__ J axpy(float s, const J& a, const J& y) const {    // s * a + y
    return J{fmaf(s, a.m, y.m), LAP ? fmaf(s, a.p, y.p) : 0.f, fmaf(s, a.v, y.v)};
  }
  __device__ __forceinline__ J addc(const J& a, float c) const { return J{is_v ? a.m + c : a.m, a.p, a.v + c}; }
  __device__ __forceinline__ J rsubc(float c, const J& a) const { return J{is_v ? c - a.m : -a.m, -a.p, c - a.v}; }
  __device__ __forceinline__ J recip(const J& a) const { const float r = 1.f / a.v; return unary(a, r, -r * r, 2.f * r * r * r); }
  __device__ __forceinline__ J log(const J& a) const { const float r = 1.f / a.v; return unary(a, logf(a.v), r, -r * r); }
  __device__ __forceinline__ J exp(const J& a) const { const float e = expf(a.v); return unary(a, e, e, e); }
  // a / b with the quotient VALUE rounded as a single fp32 division (as the reference computes it)
  __device__ __forceinline__ J div(const J& a, const J& b) const { J q = mul(a, recip(b)); q.v = a.v / b.v; if (is_v) q.m = q.v; return q; }
};

// tanh / sigmoid through ex2.approx + rcp.approx (about 1e-7 absolute error, measured in tests/test_gpu_live.py):
// the accurate libdevice versions cost ~3x the issue slots and the glue around the MLPs is issue bound.
// (fast_tanh / fast_sigmoid live in common.cuh)

// tanh on a 1-register bundle (MLP hidden layers): needs |grad|^2 on the Laplacian lane.
template <int D, bool LAP>
__device__ __forceinline__ float tanh_bundle(const Ctx<D, LAP>& cx, float a) {
  if constexpr (LAP) {
    const float th = fast_tanh(cx.bv(a));
    const float f1 = 1.f - th * th;
    const float gg = cx.gsum(a * a);
    float r = f1 * a;
    if (cx.is_l) r = fmaf(-2.f * th * f1, gg, r);
    return cx.is_v ? th : r;
  } else return fast_tanh(a);
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// thread-private scratch column in shared memory: slot-major, so a warp touches 32 consecutive words (conflict-free)
struct Scratch {
  float* col;
  __device__ __forceinline__ float& operator[](int slot) const { return col[slot * LIVE_THREADS]; }
};

// MADE connectivity (model_factory.py:8-19): hidden unit i has degree i % (D-1); a unit of degree c sees inputs of degree
// <= c, output dimension dd sees hidden units of degree <= dd-1.  The host packs the hidden units SORTED BY DEGREE
// (_live.pack_net), so "degree <= c" is the index prefix [0, deg_prefix<D>(c)) and whole segments of the (unrolled)
// reduction loops can be skipped with warp-uniform branches -- a third of the layer-2/3 FMAs at D = 4.
template <int D> __host__ __device__ constexpr int deg_prefix(int c) {      // number of hidden units with degree <= c
  int n = 0;
  for (int i = 0; i < WF_HIDDEN; ++i) n += ((i % (D - 1)) <= c) ? 1 : 0;
  return n;
}
template <int D> __host__ __device__ constexpr int unit_degree_sorted(int j) {   // degree of the j-th unit in sorted order
  int c = 0;
  while (deg_prefix<D>(c) <= j) ++c;
  return c;
}

// acc[0..8) += sum over the hidden units of degree C of h[i] * W[i * stride + 0..8)   (compile-time index range)
template <int D, int C>
__device__ __forceinline__ void fma_degree_segment(const float (&h)[WF_HIDDEN], const float* __restrict__ W, int stride,
                                                   float (&acc)[8]) {
  constexpr int LO = C == 0 ? 0 : deg_prefix<D>(C - 1);
  constexpr int HI = deg_prefix<D>(C);
#pragma unroll
  for (int i = LO; i < HI; ++i) {
    const float4 wa = lds4(W + i * stride), wb = lds4(W + i * stride + 4);
    acc[0] = fmaf(h[i], wa.x, acc[0]); acc[1] = fmaf(h[i], wa.y, acc[1]);
    acc[2] = fmaf(h[i], wa.z, acc[2]); acc[3] = fmaf(h[i], wa.w, acc[3]);
    acc[4] = fmaf(h[i], wb.x, acc[4]); acc[5] = fmaf(h[i], wb.y, acc[5]);
    acc[6] = fmaf(h[i], wb.z, acc[6]); acc[7] = fmaf(h[i], wb.w, acc[7]);
  }
}
// all segments of degree <= cmax (cmax is warp-uniform)
template <int D, int C = 0>
__device__ __forceinline__ void fma_degree_prefix(const float (&h)[WF_HIDDEN], const float* __restrict__ W, int stride,
                                                  int cmax, float (&acc)[8]) {
  if constexpr (C < D - 1) {
    if (C <= cmax) fma_degree_segment<D, C>(h, W, stride, acc);
    fma_degree_prefix<D, C + 1>(h, W, stride, cmax, acc);
  }
}
// degree of the last unit of the 8-wide output block starting at j0 (sorted order)
template <int D, int C = 1>
__device__ __forceinline__ int block_degree(int j0, int cur = 0) {
  if constexpr (C < D - 1) return block_degree<D, C + 1>(j0, (j0 + 7 >= deg_prefix<D>(C - 1)) ? C : cur);
  else return cur;
}

// Hidden layers of one conditioner: h = tanh(tanh(u W1 + b1) W2 + b2) (masked weights are stored as zeros).
// On return h[] holds the second hidden layer and the scratch column is free.
template <int D, bool LAP>
__device__ __forceinline__ void mlp_hidden(const Ctx<D, LAP>& cx, const float* __restrict__ net, const float (&u)[D],
                                           const Scratch& S, float (&h)[WF_HIDDEN]) {
  const float* W1 = net;
  const float* b1 = W1 + D * WF_HIDDEN;
  const float* W2 = b1 + WF_HIDDEN;
  const float* b2 = W2 + WF_HIDDEN * WF_HIDDEN;
#pragma unroll 1
  for (int j0 = 0; j0 < WF_HIDDEN; j0 += 4) {
    const float4 bb = lds4(b1 + j0);
    float acc[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t] = cx.is_v ? acc[t] : 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float4 w4 = lds4(W1 + d * WF_HIDDEN + j0);
      acc[0] = fmaf(u[d], w4.x, acc[0]); acc[1] = fmaf(u[d], w4.y, acc[1]);
      acc[2] = fmaf(u[d], w4.z, acc[2]); acc[3] = fmaf(u[d], w4.w, acc[3]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) S[j0 + t] = tanh_bundle<D, LAP>(cx, acc[t]);
  }
#pragma unroll
  for (int i = 0; i < WF_HIDDEN; ++i) h[i] = S[i];
#pragma unroll 1
  for (int j0 = 0; j0 < WF_HIDDEN; j0 += 8) {
    const float4 ba = lds4(b2 + j0), bb = lds4(b2 + j0 + 4);
    float acc[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = cx.is_v ? acc[t] : 0.f;
    // inputs needed by this block of 8 outputs: all units of degree <= degree(last output of the block)
    fma_degree_prefix<D>(h, W2 + j0, WF_HIDDEN, block_degree<D>(j0), acc);
#pragma unroll
    for (int t = 0; t < 8; ++t) S[j0 + t] = tanh_bundle<D, LAP>(cx, acc[t]);   // h (layer 1) is already in registers
  }
#pragma unroll
  for (int i = 0; i < WF_HIDDEN; ++i) h[i] = S[i];
}

// Output layer for dimension dd: S[q] = h . W3p[:, dd, q] + b3p[dd, q], q < 32 (padded columns carry zero weights).
// Dimension 0 of a MADE conditioner depends on the bias only (output degree -1, model_factory.py:15).
template <int D, bool LAP>
__device__ __forceinline__ void mlp_out(const Ctx<D, LAP>& cx, const float* __restrict__ net, int dd,
                                        const float (&h)[WF_HIDDEN], const Scratch& S) {
  const float* W3 = net + D * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN * WF_HIDDEN + WF_HIDDEN;
  const float* b3 = W3 + WF_HIDDEN * D * WF_MAX_P + dd * WF_MAX_P;
#pragma unroll 1
  for (int q0 = 0; q0 < WF_MAX_P; q0 += 8) {
    const float4 ba = lds4(b3 + q0), bb = lds4(b3 + q0 + 4);
    float acc[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = cx.is_v ? acc[t] : 0.f;
    // hidden units of degree c feed output dimension dd iff c <= dd - 1 (dimension 0: bias only)
    fma_degree_prefix<D>(h, W3 + dd * WF_MAX_P + q0, D * WF_MAX_P, dd - 1, acc);
#pragma unroll
    for (int t = 0; t < 8; ++t) S[q0 + t] = acc[t];
  }
}

// sum_q c_q (x) basis_q^{(k)}(x), assembled from the bundle sums A_k = sum_q f_k[q] c_q  (k, k+1, k+2).
// xd: this lane's derivative component of the spline argument (unused on the value lane).
template <int D, bool LAP>
__device__ __forceinline__ J spline_assemble(const Ctx<D, LAP>& cx, const J& A0, const J& A1, const J& A2, float xd) {
  if constexpr (LAP)
    return J{cx.is_v ? A0.m : fmaf(xd, A1.v, A0.m), cx.is_g ? fmaf(xd * xd, A2.v, fmaf(2.f * xd, A1.m, A0.p)) : 0.f, A0.v};
  else return A0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sigmoid-coefficient spline (IMADE layer / M prior).  Coefficients of the reference:
//   p = sigmoid(o); p /= sum p; p += reg; remove_bias; boundary mask; renormalise     (model_factory.py:61-69,
//   made.py:67-72, isplines_jax.py:158-202)  ==  c_q = w_q (s_q / S + reg) / sum_q' w_q' (s_q' / S + reg),
// with w_q = remove_bias scale x boundary mask (host-computed).  Everything downstream is bilinear in (s_q), 1/S, 1/Z,
// so only sums over q with lane-uniform coefficients are accumulated:
//   pass A (all q):   s_q, S = sum s_q, SW = sum w_q s_q, and the prefix sum over the bases below the local-support window
//                     (identically 1 for an I-spline value, 0 otherwise); s_q is parked in the scratch column;
//   pass B (8 window bases): the interpolated basis values from the compact node records.
// NOUT = 2: value and derivative spline (IMADE), NOUT = 1: value only (M prior).  PREFIX_ONE: I-spline tables.
// In: S[q] = conditioner output o_q (1-register bundle).  xd / xv: derivative component / value of the spline argument.
template <int D, bool LAP, int NOUT, bool PREFIX_ONE>
__device__ __forceinline__ void sigmoid_spline(const Ctx<D, LAP>& cx, const Scratch& S, int P, const float* __restrict__ wq,
                                               float reg, const float* __restrict__ rec, const int32_t* __restrict__ lo,
                                               const float* __restrict__ dense, int T, float xd, float xv, J& y, J& dy) {
  constexpr int NK = LAP ? NOUT + 2 : NOUT;
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(xv, T);
  const int lo_l = __ldg(lo + n.l), lo_r = __ldg(lo + n.r);
  const int sh = lo_r - lo_l;
  const bool local = (sh == 0) || (sh == 1);      // false only for arguments outside [0, 1] (wrapped gather rows)
  const int lo_w = local ? lo_l : 0;

  J Ssum = {0.f, 0.f, 0.f}, SW = {0.f, 0.f, 0.f}, PRE = {0.f, 0.f, 0.f};
  float Wsum = 0.f, Wpre = 0.f;
#pragma unroll 2
  for (int q = 0; q < P; ++q) {
    const float o = S[q];
    const float ov = cx.bv(o);
    const float s = fast_sigmoid(ov);
    const float d1 = s * (1.f - s);
    const J sq = cx.unary(J{o, 0.f, ov}, s, d1, d1 * (1.f - 2.f * s));
    const float w = wq[q];
    Ssum = cx.add(Ssum, sq);
    SW = cx.axpy(w, sq, SW);
    Wsum += w;
    if (PREFIX_ONE) {
      const float wp = (q < lo_w) ? w : 0.f;
      PRE = cx.axpy(wp, sq, PRE);
      Wpre += wp;
    }
    S[q] = sq.m;
    if (LAP) S[WF_MAX_P + q] = sq.p;
  }
  J Sk[NK];
  float Wk[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) { Sk[k] = J{0.f, 0.f, 0.f}; Wk[k] = 0.f; }
  if (PREFIX_ONE) { Sk[0] = PRE; Wk[0] = Wpre; }
  if (local) {
#pragma unroll 2
    for (int t = 0; t < WF_WIN; ++t) {
      const int q = lo_l + t;
      const int qc = q < P ? q : P - 1;
      const float w = q < P ? wq[qc] : 0.f;
      J sq;
      sq.m = S[qc];
      sq.p = LAP ? S[WF_MAX_P + qc] : 0.f;
      sq.v = cx.bv(sq.m);
      const int tr = t - sh;                     // slot of this basis in the right node's record
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int nd = k < 3 ? k : 3;
        const float yl = __ldg(rec + ((size_t)n.l * 4 + nd) * WF_WIN + t);
        const float yrr = __ldg(rec + ((size_t)n.r * 4 + nd) * WF_WIN + (tr < 0 ? 0 : tr));
        const float yr = tr < 0 ? ((PREFIX_ONE && nd == 0) ? 1.f : 0.f) : yrr;
        const float fw = lerp_tab(yl, yr, np_, n.dx) * w;
        Sk[k] = cx.axpy(fw, sq, Sk[k]);
        Wk[k] += fw;
      }
    }
  } else {
    // reference-exact dense evaluation (JAX gather wrap/clamp semantics)
    if (PREFIX_ONE) { Sk[0] = J{0.f, 0.f, 0.f}; Wk[0] = 0.f; }
    for (int q = 0; q < P; ++q) {
      J sq;
      sq.m = S[q];
      sq.p = LAP ? S[WF_MAX_P + q] : 0.f;
      sq.v = cx.bv(sq.m);
      const float w = wq[q];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int nd = k < 3 ? k : 3;
        const float yl = __ldg(dense + ((size_t)n.l * 4 + nd) * WF_MAX_P + q);
        const float yr = __ldg(dense + ((size_t)n.r * 4 + nd) * WF_MAX_P + q);
        const float fw = lerp_tab(yl, yr, np_, n.dx) * w;
        Sk[k] = cx.axpy(fw, sq, Sk[k]);
        Wk[k] += fw;
      }
    }
  }
  // r = 1/S;  Z = SW * r + reg * Wsum;  iz = 1/Z;  N_k = Sk * r + reg * Wk;  A_k = N_k * iz
  const J r = cx.recip(Ssum);
  const J iz = cx.recip(cx.addc(cx.mul(SW, r), reg * Wsum));
  J A[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) A[k] = cx.mul(cx.addc(cx.mul(Sk[k], r), reg * Wk[k]), iz);
  if constexpr (LAP) {
    y = spline_assemble<D, LAP>(cx, A[0], A[1], A[2], xd);
    if (NOUT == 2) dy = spline_assemble<D, LAP>(cx, A[1], A[2], A[NK - 1], xd);
  } else {
    y = A[0];
    if (NOUT == 2) dy = A[NOUT - 1];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// B prior factor (wavefunctions.py:58-65, bsplines_jax.py:127-137,173-198):
//   w = o / sum o; boundary mask; w /= ||w||; c = w @ ob_to_b; c /= ||c||; phi = sum_j c_j OB_j(clip(u)).
// Both normalisations are positive rescalings, so phi = sign(sum o) * (sum_j c'_j OB_j) / ||c'||, c' = (mask o) @ ob_to_b.
// ob_s: ob_to_b zero-padded to [32][32] in shared memory.  In: S[q] = o_q.
template <int D, bool LAP>
__device__ __forceinline__ J bprior_factor(const Ctx<D, LAP>& cx, const Scratch& S, int P, const float* __restrict__ wq,
                                           const float* __restrict__ ob_s, const float* __restrict__ tab, int T,
                                           float xd_in, float xv, bool folded) {
  constexpr int NK = LAP ? 3 : 1;
  // clip(u, 0, 1): derivative 1 strictly inside, 0 outside
  const float xc = fminf(fmaxf(xv, 0.f), 1.f);
  const float xd = ((xv > 0.f) && (xv < 1.f)) ? xd_in : 0.f;
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(xc, T);
  // folded (bit 2 of bc_P): the host multiplied mask @ ob_to_b into the third layer, S[q] already holds c'_q and S[31] the
  // sum of the raw outputs (the per-walker 32 x 32 product below is skipped)
  float osum = 0.f;
  if (folded) {
    osum = S[WF_MAX_P - 1];
  } else {
#pragma unroll 4
    for (int q = 0; q < WF_MAX_P; ++q) {
      const float o = S[q];
      if (q < P) osum += o;
      S[q] = o * wq[q];                     // wq is 0 beyond P and on the constrained ends
    }
  }
  const float sgn = cx.bv(osum) < 0.f ? -1.f : 1.f;
  J A[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) A[k] = J{0.f, 0.f, 0.f};
  J Q = {0.f, 0.f, 0.f};                  // Q = sum_j c'_j^2
#pragma unroll 1
  for (int j0 = 0; j0 < P; j0 += 4) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (folded) {
#pragma unroll
      for (int t = 0; t < 4; ++t) c[t] = (j0 + t < P) ? S[j0 + t] : 0.f;
    } else {
#pragma unroll 8
      for (int i = 0; i < WF_MAX_P; ++i) {
        const float ow = S[i];
        const float4 w4 = lds4(ob_s + i * WF_MAX_P + j0);
        c[0] = fmaf(ow, w4.x, c[0]); c[1] = fmaf(ow, w4.y, c[1]);
        c[2] = fmaf(ow, w4.z, c[2]); c[3] = fmaf(ow, w4.w, c[3]);
      }
    }
    float f[NK][4];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)n.l * 4 + k) * WF_MAX_P + j0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)n.r * 4 + k) * WF_MAX_P + j0));
      f[k][0] = lerp_tab(a.x, b.x, np_, n.dx); f[k][1] = lerp_tab(a.y, b.y, np_, n.dx);
      f[k][2] = lerp_tab(a.z, b.z, np_, n.dx); f[k][3] = lerp_tab(a.w, b.w, np_, n.dx);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float cv = cx.bv(c[t]);
#pragma unroll
      for (int k = 0; k < NK; ++k) { A[k].m = fmaf(f[k][t], c[t], A[k].m); A[k].v = fmaf(f[k][t], cv, A[k].v); }
      Q.v = fmaf(cv, cv, Q.v);
      if constexpr (LAP) {
        Q.m = cx.is_v ? Q.v : fmaf(2.f * cv, c[t], Q.m);
        Q.p = cx.is_g ? fmaf(2.f * c[t], c[t], Q.p) : 0.f;
      } else Q.m = Q.v;
    }
  }
  J num;
  if constexpr (LAP) num = spline_assemble<D, LAP>(cx, A[0], A[1], A[2], xd); else num = A[0];
  // phi = sgn * num * Q^{-1/2}
  const float isq = 1.f / sqrtf(Q.v);
  const J iq = cx.unary(Q, isq, -0.5f * isq / Q.v, 0.75f * isq / (Q.v * Q.v));
  return cx.scale(cx.mul(num, iq), sgn);
}

// BoxTransformLayer forward (made.py:118-137 'first', :156-183 'mean') on jets: xs -> us (1-register bundles), ld -= log-dets.
// Without a box the inputs pass through unchanged.
template <int D, bool LAP>
__device__ __forceinline__ void box_transform(const Ctx<D, LAP>& cx, const wf_live_model& M, const float (&xs)[D], float (&us)[D], J& ld) {
  J X[D];
#pragma unroll
  for (int d = 0; d < D; ++d) X[d] = J{cx.is_v ? xs[d] : (cx.comp == d + 1 ? 1.f : 0.f), 0.f, xs[d]};
  if (M.has_box) {
    const float L = M.box, tolr = 1e-7f;
    J U[D];
    if (M.coord_mean) {
      J sum = X[0];
#pragma unroll
      for (int d = 1; d < D; ++d) sum = cx.add(sum, X[d]);
      J mean = cx.scale(sum, 1.f / (float)D);
      mean.v = sum.v / (float)D; if (cx.is_v) mean.m = mean.v;
      const J l = cx.sub(mean, X[0]);
      const J wd = cx.sub(X[D - 1], X[0]);
      J space = cx.constant(2.f * L);
#pragma unroll
      for (int i = 0; i < D - 1; ++i) {
        const J diff = cx.sub(X[i + 1], X[i]);
        const J den = cx.addc(space, tolr);
        U[i] = cx.div(diff, den);
        ld = cx.sub(ld, cx.log(den));
        space = cx.sub(space, diff);
      }
      const J den = cx.addc(cx.rsubc(2.f * L, wd), tolr);
      U[D - 1] = cx.div(cx.sub(cx.addc(mean, L), l), den);
      ld = cx.sub(ld, cx.log(den));
    } else {
      U[0] = cx.scale(cx.addc(X[0], L), 1.f / (2.f * L));
      U[0].v = (xs[0] + L) / (2.f * L); if (cx.is_v) U[0].m = U[0].v;
      ld = cx.addc(ld, -logf(2.f * L));
#pragma unroll
      for (int i = 1; i < D; ++i) {
        const J den = cx.addc(cx.rsubc(L, X[i - 1]), tolr);
        U[i] = cx.div(cx.sub(X[i], X[i - 1]), den);
        ld = cx.sub(ld, cx.log(den));
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) us[d] = cx.fold(U[d]);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) us[d] = X[d].m;
  }
}

}  // namespace wf
