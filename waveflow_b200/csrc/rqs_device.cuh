// Device-side rational-quadratic spline evaluation shared by the operator kernel (rqs.cu) and the fused coupling flow.
// Follows flows/bijections/neural_splines.py:74-184 step by step.
#pragma once
#include "common.cuh"

namespace wf {

constexpr float RQS_MIN_BIN = 1e-3f;   // DEFAULT_MIN_BIN_WIDTH / HEIGHT (neural_splines.py:6-7)
constexpr float RQS_MIN_DER = 1e-3f;   // DEFAULT_MIN_DERIVATIVE (neural_splines.py:8)
constexpr float RQS_EPS = 1e-6f;       // searchsorted eps (neural_splines.py:11)

__device__ __forceinline__ float softplus_f(float x) {
  // jax.nn.softplus = logaddexp(x, 0)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

template <int KMAX, bool VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ row, int K, float (&a)[KMAX]) {
  if (VEC) {
#pragma unroll
    for (int j = 0; j < KMAX; j += 4) {
      if (j < K) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + j));
        a[j] = v.x; a[j + 1] = v.y; a[j + 2] = v.z; a[j + 3] = v.w;
      } else {
        a[j] = a[j + 1] = a[j + 2] = a[j + 3] = -INFINITY;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) a[j] = j < K ? __ldg(row + j) : -INFINITY;
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// In place: unnormalised -> right knots.  a[j] <- knot_{j+1} (neural_splines.py:98-107 / :111-120); knot_0 = lo.
// Also returns count = #{j in 0..K : x >= knot_j (+eps on the last)} (neural_splines.py:11-13).
// mx = max_j a[j] (callers that already know it pass it in).  FULL: K == KMAX, every bound check folds at compile time.
template <int KMAX, bool FULL>
__device__ __forceinline__ int knots_impl(float (&a)[KMAX], int K_in, float lo, float hi, float x, float mx) {
  const int K = FULL ? KMAX : K_in;
  // softmax through ex2.approx on pre-scaled arguments and one reciprocal: ~1e-7 relative on the widths, far inside the
  // 1e-5 parity budget, and 3x fewer issue slots than expf + IEEE division per bin (this loop is the kernel's hot spot);
  // two interleaved partial sums halve the dependent-add chain
  constexpr float LOG2E = 1.4426950408889634f;
  const float mxs = mx * LOG2E;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) {
      a[j] = ex2_approx(fmaf(a[j], LOG2E, -mxs));
      if (j & 1) s1 += a[j]; else s0 += a[j];
    }
  const float sum = s0 + s1;
  const float scale = (1.f - RQS_MIN_BIN * (float)K) / sum;
  const float span = hi - lo;
  float c = 0.f;
  int count = (x >= lo) ? 1 : 0;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) {
      const float w = fmaf(scale, a[j], RQS_MIN_BIN);
      c += w;
      float kn = fmaf(span, c, lo);
      if (j == K - 1) kn = hi;
      a[j] = kn;
      const float cmp = (j == K - 1) ? kn + RQS_EPS : kn;
      count += (x >= cmp) ? 1 : 0;
    }
  return count;
}
// Exact-bin variant (WF_RQS_EXACT_BINS): the knot positions are computed with exactly the float32 operation sequence of the
// float32 restatement of neural_splines.py:98-107 (oracle/rqs.py: correctly rounded exp, SEQUENTIAL sums, one IEEE division
// per bin, no fused multiply-adds), so the located bin is bit-identical to that arithmetic for every input -- including
// inputs within an ulp of a knot, where the ex2.approx softmax above may land in the neighbouring bin.
//   e_j = fl32(exp(a_j - max))  [exp evaluated in float64 and rounded once],  s = ((e_0 + e_1) + e_2) + ...,
//   w_j = min + (1 - min K) * (e_j / s),  c_j = c_{j-1} + w_j,  knot_{j+1} = (hi - lo) * c_j + lo,  knot_K = hi.
template <int KMAX>
__device__ __forceinline__ int knots_exact(float (&a)[KMAX], int K, float lo, float hi, float x) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) mx = fmaxf(mx, a[j]);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) {
      a[j] = (float)exp((double)__fsub_rn(a[j], mx));
      sum = __fadd_rn(sum, a[j]);
    }
  const float c1 = __fsub_rn(1.f, __fmul_rn(RQS_MIN_BIN, (float)K));
  const float span = (float)((double)hi - (double)lo);
  float c = 0.f;
  int count = (x >= lo) ? 1 : 0;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) {
      const float w = __fadd_rn(RQS_MIN_BIN, __fmul_rn(c1, __fdiv_rn(a[j], sum)));
      c = __fadd_rn(c, w);
      float kn = __fadd_rn(__fmul_rn(span, c), lo);
      if (j == K - 1) kn = hi;
      a[j] = kn;
      const float cmp = (j == K - 1) ? __fadd_rn(kn, RQS_EPS) : kn;
      count += (x >= cmp) ? 1 : 0;
    }
  return count;
}
// FULL = true asserts K == KMAX (every bound check folds at compile time); callers that know K statically say so.
template <int KMAX, bool FULL = false>
__device__ __forceinline__ int knots_inplace(float (&a)[KMAX], int K, float lo, float hi, float x, float mx) {
  return knots_impl<KMAX, FULL>(a, K, lo, hi, x, mx);
}
template <int KMAX, bool FULL = false>
__device__ __forceinline__ int knots_inplace(float (&a)[KMAX], int K_in, float lo, float hi, float x) {
  const int K = FULL ? KMAX : K_in;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) mx = fmaxf(mx, a[j]);
  return knots_impl<KMAX, FULL>(a, K, lo, hi, x, mx);
}

template <int KMAX>
__device__ __forceinline__ void knot_pair(const float (&a)[KMAX], int K, float lo, int idx, float& kl, float& kr) {
  kl = lo; kr = lo;
#pragma unroll
  for (int j = 0; j < KMAX; ++j)
    if (j < K) {
      if (j == idx - 1) kl = a[j];
      if (j == idx) kr = a[j];
    }
}

// Located bin of an element: knots of the bin in both directions
struct RqsBin { int idx; float cwl, in_w, chl, in_h; };

// a/b: unnormalised widths/heights (destroyed: they become the right knots).
template <int KMAX>
__device__ __forceinline__ RqsBin rqs_locate_counts(float x, const float (&a)[KMAX], const float (&b)[KMAX], int K, float B, bool inverse,
                                                    int cnt_w, int cnt_h) {
  RqsBin r;
  r.idx = min(max((inverse ? cnt_h : cnt_w) - 1, 0), K - 1);
  float cwr, chr;
  knot_pair<KMAX>(a, K, -B, r.idx, r.cwl, cwr);
  knot_pair<KMAX>(b, K, -B, r.idx, r.chl, chr);
  r.in_w = cwr - r.cwl; r.in_h = chr - r.chl;
  return r;
}
template <int KMAX>
__device__ __forceinline__ RqsBin rqs_locate_exact(float x, float (&a)[KMAX], float (&b)[KMAX], int K, float B, bool inverse) {
  const int cnt_w = knots_exact<KMAX>(a, K, -B, B, x);
  const int cnt_h = knots_exact<KMAX>(b, K, -B, B, x);
  return rqs_locate_counts<KMAX>(x, a, b, K, B, inverse, cnt_w, cnt_h);
}
template <int KMAX, bool FULL = false>
__device__ __forceinline__ RqsBin rqs_locate(float x, float (&a)[KMAX], float (&b)[KMAX], int K, float B, bool inverse) {
  const int cnt_w = knots_inplace<KMAX, FULL>(a, K, -B, B, x);
  const int cnt_h = knots_inplace<KMAX, FULL>(b, K, -B, B, x);
  return rqs_locate_counts<KMAX>(x, a, b, FULL ? KMAX : K, B, inverse, cnt_w, cnt_h);
}
// same, for rows whose maxima are already known (the coupling layer's own 2B * softmax puts the maximum at exactly 2B / sum)
template <int KMAX, bool FULL = false>
__device__ __forceinline__ RqsBin rqs_locate(float x, float (&a)[KMAX], float (&b)[KMAX], int K, float B, bool inverse, float mx_a,
                                             float mx_b) {
  const int cnt_w = knots_inplace<KMAX, FULL>(a, K, -B, B, x, mx_a);
  const int cnt_h = knots_inplace<KMAX, FULL>(b, K, -B, B, x, mx_b);
  return rqs_locate_counts<KMAX>(x, a, b, FULL ? KMAX : K, B, inverse, cnt_w, cnt_h);
}

// ud0 / ud1: unnormalised derivatives at the two knots of the bin (interior entries idx-1 / idx; ignored at the ends,
// where the reference pads with log(exp(1 - min_derivative) - 1), neural_splines.py:33-42)
__device__ __forceinline__ void rqs_finish(float x, const RqsBin& r, int K, float ud0_in, float ud1_in, bool inverse,
                                           float& out, float& lad) {
  const float in_w = r.in_w, in_h = r.in_h, cwl = r.cwl, chl = r.chl;
  const float delta = in_h / in_w;
  const float cpad = logf(expf(1.f - RQS_MIN_DER) - 1.f);
  const float ud0 = (r.idx == 0) ? cpad : ud0_in;
  const float ud1 = (r.idx == K - 1) ? cpad : ud1_in;
  const float d0 = RQS_MIN_DER + softplus_f(ud0);
  const float d1 = RQS_MIN_DER + softplus_f(ud1);
  const float s = d0 + d1 - 2.f * delta;
  float theta;
  if (inverse) {
    const float dy = x - chl;
    const float qa = dy * s + in_h * (delta - d0);
    const float qb = in_h * d0 - dy * s;
    const float qc = -delta * dy;
    const float disc = qb * qb - 4.f * qa * qc;
    theta = (2.f * qc) / (-qb - sqrtf(disc));
    out = theta * in_w + cwl;
  } else {
    theta = (x - cwl) / in_w;
  }
  const float t1mt = theta * (1.f - theta);
  const float den = delta + s * t1mt;
  if (!inverse) {
    const float numer = in_h * (delta * theta * theta + d0 * t1mt);
    out = chl + numer / den;
  }
  const float omt = 1.f - theta;
  const float num = delta * delta * (d1 * theta * theta + 2.f * delta * t1mt + d0 * omt * omt);
  const float l = logf(num) - 2.f * logf(den);
  lad = inverse ? -l : l;
}

// dget(j), j in [0, K-2]: unnormalised interior derivative j.
template <int KMAX, bool FULL = false, class DGet>
__device__ __forceinline__ void rqs_eval(float x, float (&a)[KMAX], float (&b)[KMAX], int K, float B, bool inverse,
                                         DGet dget, float& out, float& lad, int& bin) {
  const RqsBin r = rqs_locate<KMAX, FULL>(x, a, b, K, B, inverse);
  bin = r.idx;
  const float ud0 = (r.idx == 0) ? 0.f : dget(r.idx - 1);
  const float ud1 = (r.idx == K - 1) ? 0.f : dget(r.idx);
  rqs_finish(x, r, K, ud0, ud1, inverse, out, lad);
}

template <int KMAX, class DGet>
__device__ __forceinline__ void rqs_eval_exact(float x, float (&a)[KMAX], float (&b)[KMAX], int K, float B, bool inverse,
                                               DGet dget, float& out, float& lad, int& bin) {
  const RqsBin r = rqs_locate_exact<KMAX>(x, a, b, K, B, inverse);
  bin = r.idx;
  const float ud0 = (r.idx == 0) ? 0.f : dget(r.idx - 1);
  const float ud1 = (r.idx == K - 1) ? 0.f : dget(r.idx);
  rqs_finish(x, r, K, ud0, ud1, inverse, out, lad);
}

template <int KMAX, bool FULL = false, class DGet>
__device__ __forceinline__ void rqs_eval(float x, float (&a)[KMAX], float (&b)[KMAX], int K, float B, bool inverse, float mx_a,
                                         float mx_b, DGet dget, float& out, float& lad, int& bin) {
  const RqsBin r = rqs_locate<KMAX, FULL>(x, a, b, K, B, inverse, mx_a, mx_b);
  bin = r.idx;
  const float ud0 = (r.idx == 0) ? 0.f : dget(r.idx - 1);
  const float ud1 = (r.idx == K - 1) ? 0.f : dget(r.idx);
  rqs_finish(x, r, K, ud0, ud1, inverse, out, lad);
}

// 2B * softmax(raw[0..K))  (neural_splines.py:260-261: the coupling layer's own normalisation, before RQS repeats it), in place
// Returns the maximum of the result: the largest input maps to exp(0) = 1 exactly, so it is exactly 2B / sum.
template <int KP, bool FULL>
__device__ __forceinline__ float softmax_2b_impl(float (&a)[KP], int K_in, float twoB) {
  const int K = FULL ? KP : K_in;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < KP; ++j)
    if (j < K) mx = fmaxf(mx, a[j]);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < KP; ++j)
    if (j < K) {
      a[j] = __expf(a[j] - mx);
      if (j & 1) s1 += a[j]; else s0 += a[j];
    }
  const float s = twoB / (s0 + s1);
#pragma unroll
  for (int j = 0; j < KP; ++j)
    if (j < K) a[j] *= s;
  return s;
}
template <int KP, bool FULL = false>
__device__ __forceinline__ float softmax_2b(float (&a)[KP], int K, float twoB) {
  return softmax_2b_impl<KP, FULL>(a, K, twoB);
}

}  // namespace wf
