// Second-order jets (value, gradient w.r.t. the D walker coordinates, Laplacian) and their adjoints: the algebra of the
// reverse pass through the forward-mode Laplacian (value_and_grad(loss_fn_efficient), vqmc.py:193-221, over
// physics.py:50-52).  A jet s = (v, g[D], l); every *_bwd accumulates  in_bar += (d out / d in)^T out_bar.
#pragma once
#include "common.cuh"

namespace wf {
namespace train {

template <int D>
struct Jet {
  float v, g[D], l;
};

template <int D>
__device__ __forceinline__ Jet<D> jzero() {
  Jet<D> j;
  j.v = 0.f; j.l = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) j.g[i] = 0.f;
  return j;
}
template <int D>
__device__ __forceinline__ float gdot(const Jet<D>& a, const Jet<D>& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) s = fmaf(a.g[i], b.g[i], s);
  return s;
}
template <int D>
__device__ __forceinline__ void jacc(Jet<D>& a, const Jet<D>& b) {
  a.v += b.v; a.l += b.l;
#pragma unroll
  for (int i = 0; i < D; ++i) a.g[i] += b.g[i];
}
template <int D>
__device__ __forceinline__ void jaxpy(Jet<D>& a, float c, const Jet<D>& b) {
  a.v = fmaf(c, b.v, a.v); a.l = fmaf(c, b.l, a.l);
#pragma unroll
  for (int i = 0; i < D; ++i) a.g[i] = fmaf(c, b.g[i], a.g[i]);
}
template <int D>
__device__ __forceinline__ Jet<D> jscale(const Jet<D>& a, float c) {
  Jet<D> o;
  o.v = a.v * c; o.l = a.l * c;
#pragma unroll
  for (int i = 0; i < D; ++i) o.g[i] = a.g[i] * c;
  return o;
}
template <int D>
__device__ __forceinline__ Jet<D> jsub(const Jet<D>& a, const Jet<D>& b) {
  Jet<D> o;
  o.v = a.v - b.v; o.l = a.l - b.l;
#pragma unroll
  for (int i = 0; i < D; ++i) o.g[i] = a.g[i] - b.g[i];
  return o;
}

// out = a * b
template <int D>
__device__ __forceinline__ Jet<D> jmul(const Jet<D>& a, const Jet<D>& b) {
  Jet<D> o;
  o.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < D; ++i) o.g[i] = fmaf(a.v, b.g[i], b.v * a.g[i]);
  o.l = fmaf(a.v, b.l, fmaf(b.v, a.l, 2.f * gdot(a, b)));
  return o;
}
// a_bar += (d(a*b)/da)^T ob   (call again with the roles swapped for b_bar)
template <int D>
__device__ __forceinline__ void jmul_bwd(const Jet<D>& b, const Jet<D>& ob, Jet<D>& ab) {
  ab.v += fmaf(ob.v, b.v, fmaf(ob.l, b.l, gdot(ob, b)));
#pragma unroll
  for (int i = 0; i < D; ++i) ab.g[i] += fmaf(ob.g[i], b.v, 2.f * ob.l * b.g[i]);
  ab.l = fmaf(ob.l, b.v, ab.l);
}

// out = f(a) with f0 = f(a.v), f1 = f', f2 = f''
template <int D>
__device__ __forceinline__ Jet<D> junary(const Jet<D>& a, float f0, float f1, float f2) {
  Jet<D> o;
  o.v = f0;
#pragma unroll
  for (int i = 0; i < D; ++i) o.g[i] = f1 * a.g[i];
  o.l = fmaf(f1, a.l, f2 * gdot(a, a));
  return o;
}
// a_bar += (d f(a)/da)^T ob;  needs f', f'', f'''
template <int D>
__device__ __forceinline__ void junary_bwd(const Jet<D>& a, float f1, float f2, float f3, const Jet<D>& ob, Jet<D>& ab) {
  const float s2 = gdot(a, a);
  ab.v += fmaf(ob.v, f1, fmaf(f2, gdot(ob, a), ob.l * fmaf(f2, a.l, f3 * s2)));
#pragma unroll
  for (int i = 0; i < D; ++i) ab.g[i] += fmaf(f1, ob.g[i], 2.f * f2 * ob.l * a.g[i]);
  ab.l = fmaf(f1, ob.l, ab.l);
}

template <int D>
__device__ __forceinline__ Jet<D> jrecip(const Jet<D>& a) {
  const float r = 1.f / a.v;
  return junary(a, r, -r * r, 2.f * r * r * r);
}
template <int D>
__device__ __forceinline__ void jrecip_bwd(const Jet<D>& a, const Jet<D>& ob, Jet<D>& ab) {
  const float r = 1.f / a.v, r2 = r * r;
  junary_bwd(a, -r2, 2.f * r2 * r, -6.f * r2 * r2, ob, ab);
}
template <int D>
__device__ __forceinline__ Jet<D> jlog(const Jet<D>& a) {
  const float r = 1.f / a.v;
  return junary(a, logf(a.v), r, -r * r);
}
template <int D>
__device__ __forceinline__ void jlog_bwd(const Jet<D>& a, const Jet<D>& ob, Jet<D>& ab) {
  const float r = 1.f / a.v, r2 = r * r;
  junary_bwd(a, r, -r2, 2.f * r2 * r, ob, ab);
}
template <int D>
__device__ __forceinline__ Jet<D> jrsqrt(const Jet<D>& a) {
  const float s = rsqrtf(a.v), r = 1.f / a.v;
  return junary(a, s, -0.5f * s * r, 0.75f * s * r * r);
}
template <int D>
__device__ __forceinline__ void jrsqrt_bwd(const Jet<D>& a, const Jet<D>& ob, Jet<D>& ab) {
  const float s = rsqrtf(a.v), r = 1.f / a.v;
  junary_bwd(a, -0.5f * s * r, 0.75f * s * r * r, -1.875f * s * r * r * r, ob, ab);
}
struct Sig { float s, d1, d2, d3; };
__device__ __forceinline__ Sig sigmoid_derivs(float x) {
  Sig o;
  o.s = 1.f / (1.f + expf(-x));
  o.d1 = o.s * (1.f - o.s);
  o.d2 = o.d1 * (1.f - 2.f * o.s);
  o.d3 = fmaf(o.d2, 1.f - 2.f * o.s, -2.f * o.d1 * o.d1);
  return o;
}

// ---- jets stored as G = D + 2 consecutive rows of a row-major [rows][width] array
template <int D>
__device__ __forceinline__ Jet<D> jload(const float* __restrict__ base, int64_t n, int width, int col) {
  const float* p = base + n * (D + 2) * (int64_t)width + col;
  Jet<D> j;
  j.v = p[0];
#pragma unroll
  for (int i = 0; i < D; ++i) j.g[i] = p[(int64_t)(1 + i) * width];
  j.l = p[(int64_t)(D + 1) * width];
  return j;
}
template <int D>
__device__ __forceinline__ void jstore(float* __restrict__ base, int64_t n, int width, int col, const Jet<D>& j) {
  float* p = base + n * (D + 2) * (int64_t)width + col;
  p[0] = j.v;
#pragma unroll
  for (int i = 0; i < D; ++i) p[(int64_t)(1 + i) * width] = j.g[i];
  p[(int64_t)(D + 1) * width] = j.l;
}

// ---- table lookups (isplines_jax.py:45-66): basis q, derivative orders 0..3 at x, from the dense [T][4][32] layout
struct Basis4 { float f[4]; };
__device__ __forceinline__ Basis4 basis4(const float* __restrict__ tab, const NodeIdx& ni, float np_, int q) {
  Basis4 b;
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) {
    const float yl = __ldg(tab + ((int64_t)ni.l * 4 + nd) * WF_MAX_P + q);
    const float yr = __ldg(tab + ((int64_t)ni.r * 4 + nd) * WF_MAX_P + q);
    b.f[nd] = lerp_tab(yl, yr, np_, ni.dx);
  }
  return b;
}

}  // namespace train
}  // namespace wf
