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
// Reference: flows/bijections/made.py:66-81,108-183, model_factory.py:8-93, splines/isplines_jax.py:45-79,158-202,
// splines/bsplines_jax.py:127-137,173-198, wavefunctions.py:33-71, flows/distributions.py:139-163,
// utils/physics.py:50-93, vqmc.py:198-200.
#pragma once
#include <math.h>
#include "common.cuh"

namespace wf {

constexpr int LIVE_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ constexpr int net_floats(int D) {
  return D * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN * D * WF_MAX_P + D * WF_MAX_P;
}

struct LiveParams {
  wf_live_model m;
  const float* weights;   // (n_layers + has_prior_net) nets, packed (see wf_live_net_floats)
  const float* tab_I;     // [T][4][32]
  const float* tab_P;     // [T][4][32]  (OB tables for the B prior, M tables for the M prior)
  const float* ob_to_b;   // [P_P][P_P]  (B prior)
  const float* x;         // [N][D]
  int64_t N;
  float* u; float* logdet; float* logpdf; float* psi;      // forward outputs (nullable)
  float* hpsi; float* eloc; float* grad; float* lap;       // local-energy outputs (nullable)
  double* sums;                                            // {sum E, sum E^2, n, sum psi^2} (nullable)
  float wq_I[WF_MAX_P];   // remove_bias scale x boundary mask of the I-spline coefficients (0 beyond P_I)
  float wq_P[WF_MAX_P];   // same for the prior coefficients (M: remove_bias x mask, B: mask)
  float protons[WF_MAX_D];
  int n_protons;
  int nets_resident;      // all conditioner nets fit in shared memory at once
  int n_nets;
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
  __device__ __forceinline__ J from1(float a) const { return J{a, 0.f, bv(a)}; }
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
  __device__ __forceinline__ J addc(const J& a, float c) const { return J{is_v ? a.m + c : a.m, a.p, a.v + c}; }
  __device__ __forceinline__ J rsubc(float c, const J& a) const { return J{is_v ? c - a.m : -a.m, -a.p, c - a.v}; }
  __device__ __forceinline__ J recip(const J& a) const { const float r = 1.f / a.v; return unary(a, r, -r * r, 2.f * r * r * r); }
  __device__ __forceinline__ J log(const J& a) const { const float r = 1.f / a.v; return unary(a, logf(a.v), r, -r * r); }
  __device__ __forceinline__ J exp(const J& a) const { const float e = expf(a.v); return unary(a, e, e, e); }
  // a / b with the quotient VALUE rounded as a single fp32 division (as the reference computes it)
  __device__ __forceinline__ J div(const J& a, const J& b) const { J q = mul(a, recip(b)); q.v = a.v / b.v; if (is_v) q.m = q.v; return q; }
};

// tanh on a 1-register bundle (MLP hidden layers): needs |grad|^2 on the Laplacian lane.
template <int D, bool LAP>
__device__ __forceinline__ float tanh_bundle(const Ctx<D, LAP>& cx, float a) {
  if constexpr (LAP) {
    const float th = tanhf(cx.bv(a));
    const float f1 = 1.f - th * th;
    const float gg = cx.gsum(a * a);
    float r = f1 * a;
    if (cx.is_l) r = fmaf(-2.f * th * f1, gg, r);
    return cx.is_v ? th : r;
  } else return tanhf(a);
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// MADE degrees (model_factory.py:8-19): input d -> d, hidden i -> i % (D-1), output d -> d - 1.
template <int D> __host__ __device__ constexpr bool m1_on(int d, int j) { return (j % (D - 1)) >= d; }
template <int D> __host__ __device__ constexpr bool m2_on(int i, int j) { return (j % (D - 1)) >= (i % (D - 1)); }
template <int D> __host__ __device__ constexpr bool m3_on(int i, int d) { return (d - 1) >= (i % (D - 1)); }

// Two hidden layers of a conditioner: h2 = tanh(tanh(u W1 + b1) W2 + b2), masks compiled in.
template <int D, bool LAP>
__device__ __forceinline__ void mlp_hidden(const Ctx<D, LAP>& cx, const float* __restrict__ net, const float (&u)[D],
                                           float (&h2)[WF_HIDDEN]) {
  const float* W1 = net;
  const float* b1 = W1 + D * WF_HIDDEN;
  const float* W2 = b1 + WF_HIDDEN;
  const float* b2 = W2 + WF_HIDDEN * WF_HIDDEN;
  float h1[WF_HIDDEN];
#pragma unroll
  for (int j0 = 0; j0 < WF_HIDDEN; j0 += 4) {
    const float4 bb = lds4(b1 + j0);
    float acc[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t] = cx.is_v ? acc[t] : 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float4 w4 = lds4(W1 + d * WF_HIDDEN + j0);
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (m1_on<D>(d, j0 + t)) acc[t] = fmaf(u[d], w[t], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) h1[j0 + t] = tanh_bundle<D, LAP>(cx, acc[t]);
  }
#pragma unroll
  for (int j0 = 0; j0 < WF_HIDDEN; j0 += 8) {
    const float4 ba = lds4(b2 + j0), bb = lds4(b2 + j0 + 4);
    float acc[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = cx.is_v ? acc[t] : 0.f;
#pragma unroll
    for (int i = 0; i < WF_HIDDEN; ++i) {
      const float4 wa = lds4(W2 + i * WF_HIDDEN + j0), wb = lds4(W2 + i * WF_HIDDEN + j0 + 4);
      const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (m2_on<D>(i, j0 + t)) acc[t] = fmaf(h1[i], w[t], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) h2[j0 + t] = tanh_bundle<D, LAP>(cx, acc[t]);
  }
}

// Output layer for dimension dd: o[q] = h2 . W3p[:, dd, q] + b3p[dd, q], q < 32 (padded columns carry zero weights).
template <int D, bool LAP>
__device__ __forceinline__ void mlp_out(const Ctx<D, LAP>& cx, const float* __restrict__ net, const int dd,
                                        const float (&h2)[WF_HIDDEN], float (&o)[WF_MAX_P]) {
  const float* W3 = net + D * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN * WF_HIDDEN + WF_HIDDEN;
  const float* b3 = W3 + WF_HIDDEN * D * WF_MAX_P;
#pragma unroll
  for (int q0 = 0; q0 < WF_MAX_P; q0 += 8) {
    const float4 ba = lds4(b3 + dd * WF_MAX_P + q0), bb = lds4(b3 + dd * WF_MAX_P + q0 + 4);
    float acc[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = cx.is_v ? acc[t] : 0.f;
#pragma unroll
    for (int i = 0; i < WF_HIDDEN; ++i) {
      if (m3_on<D>(i, dd)) {
        const float* wr = W3 + (i * D + dd) * WF_MAX_P + q0;
        const float4 wa = lds4(wr), wb = lds4(wr + 4);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = fmaf(h2[i], w[t], acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) o[q0 + t] = acc[t];
  }
}

// Interpolated basis values f[k][t] = basis_{q0+t}^{(k)}(x) for k < NK, t < 4, from the dense [T][4][32] table.
template <int NK>
__device__ __forceinline__ void table_chunk(const float* __restrict__ tab, const NodeIdx& n, float np_, int q0,
                                            float (&f)[NK][4]) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int nd = k < 3 ? k : 3;
    const float4 a = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)n.l * 4 + nd) * WF_MAX_P + q0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)n.r * 4 + nd) * WF_MAX_P + q0));
    f[k][0] = lerp_tab(a.x, b.x, np_, n.dx);
    f[k][1] = lerp_tab(a.y, b.y, np_, n.dx);
    f[k][2] = lerp_tab(a.z, b.z, np_, n.dx);
    f[k][3] = lerp_tab(a.w, b.w, np_, n.dx);
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
// so only sums over q with lane-uniform coefficients are accumulated.
// NOUT = 2: value and derivative spline (IMADE), NOUT = 1: value only (M prior).
// xd: derivative component of the spline argument on this lane; xv: its value (table lookup position).
template <int D, bool LAP, int NOUT>
__device__ __forceinline__ void sigmoid_spline(const Ctx<D, LAP>& cx, const float (&o)[WF_MAX_P], int P,
                                               const float* __restrict__ wq, float reg, const float* __restrict__ tab,
                                               int T, float xd, float xv, J& y, J& dy) {
  constexpr int NK = LAP ? NOUT + 2 : NOUT;
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(xv, T);
  J S = {0.f, 0.f, 0.f}, SW = {0.f, 0.f, 0.f};
  J Sk[NK];
  float Wk[NK], Wsum = 0.f;
#pragma unroll
  for (int k = 0; k < NK; ++k) { Sk[k] = J{0.f, 0.f, 0.f}; Wk[k] = 0.f; }
#pragma unroll
  for (int q0 = 0; q0 < WF_MAX_P; q0 += 4) {
    if (q0 < P) {
      float f[NK][4];
      table_chunk<NK>(tab, n, np_, q0, f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int q = q0 + t;
        if (q < P) {
          const float ov = cx.bv(o[q]);
          const float s = 1.f / (1.f + expf(-ov));
          const float d1 = s * (1.f - s);
          const J sq = cx.unary(J{o[q], 0.f, ov}, s, d1, d1 * (1.f - 2.f * s));
          const float w = wq[q];
          S = cx.add(S, sq);
          SW.m = fmaf(w, sq.m, SW.m); SW.v = fmaf(w, s, SW.v);
          if (LAP) SW.p = fmaf(w, sq.p, SW.p);
          Wsum += w;
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            const float fw = f[k][t] * w;
            Sk[k].m = fmaf(fw, sq.m, Sk[k].m); Sk[k].v = fmaf(fw, s, Sk[k].v);
            if (LAP) Sk[k].p = fmaf(fw, sq.p, Sk[k].p);
            Wk[k] += fw;
          }
        }
      }
    }
  }
  // r = 1/S;  Z = SW * r + reg * Wsum;  iz = 1/Z;  N_k = Sk * r + reg * Wk;  A_k = N_k * iz
  const J r = cx.recip(S);
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
// ob_s: ob_to_b zero-padded to [32][32] in shared memory.
template <int D, bool LAP>
__device__ __forceinline__ J bprior_factor(const Ctx<D, LAP>& cx, const float (&o)[WF_MAX_P], int P,
                                           const float* __restrict__ wq, const float* __restrict__ ob_s,
                                           const float* __restrict__ tab, int T, float xd_in, float xv) {
  constexpr int NK = LAP ? 3 : 1;
  // clip(u, 0, 1): derivative 1 strictly inside, 0 outside
  const float xc = fminf(fmaxf(xv, 0.f), 1.f);
  const float xd = ((xv > 0.f) && (xv < 1.f)) ? xd_in : 0.f;
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(xc, T);
  float ow[WF_MAX_P];
  float osum = 0.f;
#pragma unroll
  for (int q = 0; q < WF_MAX_P; ++q) {
    if (q < P) osum += o[q];
    ow[q] = o[q] * wq[q];                 // wq is 0 beyond P and on the constrained ends
  }
  const float sgn = cx.bv(osum) < 0.f ? -1.f : 1.f;
  J A[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) A[k] = J{0.f, 0.f, 0.f};
  J Q = {0.f, 0.f, 0.f};                  // Q = sum_j c'_j^2
#pragma unroll
  for (int j0 = 0; j0 < WF_MAX_P; j0 += 4) {
    if (j0 < P) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < WF_MAX_P; ++i) {
        const float4 w4 = lds4(ob_s + i * WF_MAX_P + j0);
        c[0] = fmaf(ow[i], w4.x, c[0]); c[1] = fmaf(ow[i], w4.y, c[1]);
        c[2] = fmaf(ow[i], w4.z, c[2]); c[3] = fmaf(ow[i], w4.w, c[3]);
      }
      float f[NK][4];
      table_chunk<NK>(tab, n, np_, j0, f);
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
  }
  J num;
  if constexpr (LAP) num = spline_assemble<D, LAP>(cx, A[0], A[1], A[2], xd); else num = A[0];
  // phi = sgn * num * Q^{-1/2}
  const float isq = 1.f / sqrtf(Q.v);
  const J iq = cx.unary(Q, isq, -0.5f * isq / Q.v, 0.75f * isq / (Q.v * Q.v));
  return cx.scale(cx.mul(num, iq), sgn);
}

}  // namespace wf
