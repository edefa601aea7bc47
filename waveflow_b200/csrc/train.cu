// Parameter gradient of the VQMC loss (value_and_grad(loss_fn_efficient), vqmc.py:193-221): a reverse pass through the
// forward-mode Laplacian of psi.  Layer-wise formulation: every activation is a second-order jet (value, d/dx_1..D,
// Laplacian) stored as G = D + 2 consecutive rows, so the conditioner layers are plain [N*G, K] x [K, N_out] products
// (bias on the value rows only) and their weight gradients are X^T dY over all N*G rows; the element-wise pieces
// (tanh, coefficient normalisation, table splines, psi assembly) carry the jet algebra of train_jets.cuh.
//
//   forward : box -> { linear, tanh, linear, tanh, linear, spline head } x (L flow nets + prior) -> psi, H psi, E_loc
//   seed    : psi_bar = (a + V b) / N, lap_bar = -b / (2N),  a = 2 (E - avg) / psi - H psi / psi^2, b = 1 / psi
//   backward: the same chain in reverse; table derivatives are the next table, order 4 clamps to 3.
#include <math.h>
#include "train_jets.cuh"
#include "train_tc.cuh"

using namespace wf;
using namespace wf::train;

namespace {

constexpr int HID = 64;            // conditioner width (model_factory.py:40)
constexpr int LIN_THREADS = 256;
constexpr int MAX_W = 128;         // widest layer (D * P)
constexpr int HEAD_THREADS = 128;

// ---------------------------------------------------------------------------------------------- MADE masks
// model_factory.py:8-19: degrees in = arange(D), hidden = h % (D-1), out = d - 1; mask[in][out] = deg_out >= deg_in.
__host__ __device__ inline bool made_mask(int layer, int D, int k, int n) {
  if (layer == 0) return true;                    // no mask (folded prior layer: the mask is applied when unfolding)
  if (layer == 1) return (n % (D - 1)) >= k;
  if (layer == 2) return (n % (D - 1)) >= (k % (D - 1));
  return ((n % D) - 1) >= (k % (D - 1));
}

// all layers of all conditioners in one launch: blockIdx.y = 3 * net + (layer - 1)
struct MaskJobs {
  int64_t src[3 * (WF_MAX_LAYERS + 1)], dst[3 * (WF_MAX_LAYERS + 1)];
  int K[3 * (WF_MAX_LAYERS + 1)], N[3 * (WF_MAX_LAYERS + 1)];
};
__global__ void mask_weights_kernel(const float* __restrict__ params, float* __restrict__ Wm, const __grid_constant__ MaskJobs jobs, int D) {
  const int j = blockIdx.y, layer = j % 3 + 1;
  const int K = jobs.K[j], N = jobs.N[j];
  const float* W = params + jobs.src[j];
  float* out = Wm + jobs.dst[j];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
    const int k = i / N, n = i % N;
    out[i] = made_mask(layer, D, k, n) ? W[i] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------- folded prior layer
// Everything between the prior conditioner's third layer and the un-normalised B-spline coefficients is linear
// (boundary mask, ob_to_b: bsplines_jax.py:132-134,173-198), so it is multiplied into the layer once per step:
//   Wf[h][q*D + d] = sum_p W3m[h][p*D + d] * F[p][q],  F = diag(mask) @ ob_to_b;   Wf[h][D*P + d] = sum_p W3m[h][p*D + d]
// (row h = 64 is the bias).  The last D columns carry sum_p o_p, whose sign the conditioner's normalisation leaves behind.
__global__ void fold_prior_kernel(const float* __restrict__ W3m, const float* __restrict__ b3, const float* __restrict__ ob_to_b,
                                  int D, int P, float* __restrict__ Wf, float* __restrict__ bf) {
  const int NcF = D * P + D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (HID + 1) * NcF) return;
  const int h = i / NcF, c = i % NcF;
  const float* src = h < HID ? W3m + (size_t)h * D * P : b3;
  float acc = 0.f;
  if (c < D * P) {
    const int q = c / D, d = c % D;
    for (int p = 1; p < P - 1; ++p) acc = fmaf(src[p * D + d], __ldg(ob_to_b + p * P + q), acc);
  } else {
    const int d = c - D * P;
    for (int p = 0; p < P; ++p) acc += src[p * D + d];
  }
  if (h < HID) Wf[(size_t)h * NcF + c] = acc; else bf[c] = acc;
}
// gW3[h][p*D + d] += mask3 * sum_q gWf[h][q*D + d] * F[p][q]   (the sum columns have zero gradient: a sign is piecewise constant)
__global__ void unfold_prior_grad_kernel(const float* __restrict__ gWf, const float* __restrict__ gbf, const float* __restrict__ ob_to_b,
                                         int D, int P, float* __restrict__ gW3, float* __restrict__ gb3) {
  const int NcF = D * P + D, DP = D * P;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (HID + 1) * DP) return;
  const int h = i / DP, c = i % DP;
  const int p = c / D, d = c % D;
  if (p == 0 || p == P - 1) return;                                  // masked end coefficients: no gradient
  const float* src = h < HID ? gWf + (size_t)h * NcF : gbf;
  float acc = 0.f;
  for (int q = 0; q < P; ++q) acc = fmaf(src[q * D + d], __ldg(ob_to_b + p * P + q), acc);
  if (h < HID) {
    if (made_mask(3, D, h, c)) gW3[(size_t)h * DP + c] += acc;
  } else {
    gb3[c] += acc;
  }
}

// ---------------------------------------------------------------------------------------------- linear layers
// Both GEMM kernels split the CTA into independent thread groups ("slices", 8 x BN/8 threads, one 8 x 8 register block per
// thread) that stage their own row chunks in private shared memory and synchronise with named barriers only, so the
// slices of a CTA drift out of phase and one slice's global loads overlap another's FMAs.
__device__ __forceinline__ void slice_sync(int slice, int threads) {
  // immediate barrier ids (a register id would make ptxas reserve all 16 barriers)
  if (slice == 0) asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
  else if (slice == 1) asm volatile("bar.sync 2, %0;" ::"r"(threads) : "memory");
  else if (slice == 2) asm volatile("bar.sync 3, %0;" ::"r"(threads) : "memory");
  else asm volatile("bar.sync 4, %0;" ::"r"(threads) : "memory");
}

// C[R][Nc] (+)= A[R][Kc] * B (+ bias on the value rows r % G == 0).  TRANS_B == false: B [Kc][Nc]; true: B given as [Nc][Kc].
// The whole K extent is resident in shared memory: B once per CTA (k-major), A per 64-row chunk (row-major, k contiguous).
// RM rows per thread: 8 (64-row chunks) for large batches, 2 (16-row chunks, 4x the CTAs) when the batch is small
template <int BN, bool TRANS_B, bool ACCUM, int RM>
__global__ void __launch_bounds__(LIN_THREADS) linear_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             const float* __restrict__ bias, float* __restrict__ C,
                                                             int64_t R, int Kc, int Nc, int G) {
  extern __shared__ __align__(16) float sm[];
  constexpr int TX = BN / 8, TS = 8 * TX, SL = LIN_THREADS / TS, LROWS = 8 * RM;
  const int KcP = (Kc + 3) & ~3, LDA = KcP + 4;
  float* Bs = sm;                                         // [KcP][BN]
  const int tid = threadIdx.x, slice = tid / TS, ts = tid % TS, tx = ts % TX, ty = ts / TX;
  float* As = sm + KcP * BN + slice * LROWS * LDA;        // [LROWS][LDA], private to the slice
  for (int i = tid; i < KcP * BN; i += LIN_THREADS) {
    const int k = i / BN, n = i % BN;
    Bs[i] = (k < Kc && n < Nc) ? (TRANS_B ? B[(int64_t)n * Kc + k] : B[(int64_t)k * Nc + n]) : 0.f;
  }
  __syncthreads();
  const int64_t chunks = (R + LROWS - 1) / LROWS;
  for (int64_t ch = (int64_t)blockIdx.x * SL + slice; ch < chunks; ch += (int64_t)gridDim.x * SL) {
    const int64_t r0 = ch * LROWS;
    const int rows = (int)min((int64_t)LROWS, R - r0);
    slice_sync(slice, TS);
    if ((Kc & 3) == 0) {
      const int k4n = Kc >> 2;
      for (int i = ts; i < LROWS * k4n; i += TS) {
        const int row = i / k4n, k4 = i % k4n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) v = *reinterpret_cast<const float4*>(A + (r0 + row) * Kc + 4 * k4);
        *reinterpret_cast<float4*>(As + row * LDA + 4 * k4) = v;
      }
    } else {
      for (int i = ts; i < LROWS * KcP; i += TS) {
        const int row = i / KcP, k = i % KcP;
        As[row * LDA + k] = (row < rows && k < Kc) ? A[(r0 + row) * Kc + k] : 0.f;
      }
    }
    slice_sync(slice, TS);
    float acc[RM][8];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < KcP; k += 4) {
      float4 a[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty * RM + i) * LDA + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + (k + kk) * BN + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + (k + kk) * BN + BN / 2 + tx * 4);
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
          acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
          acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
          acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int rl = ty * RM + i;
      if (rl >= rows) continue;
      const int64_t r = r0 + rl;
      const bool value_row = bias && (r % G) == 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n0 = h * (BN / 2) + tx * 4;
        if ((Nc & 3) == 0) {                 // 128-bit stores (rows are 16-byte aligned when Nc % 4 == 0)
          if (n0 < Nc) {
            float4 v = make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            if (value_row) {
              v.x += bias[n0]; v.y += bias[n0 + 1]; v.z += bias[n0 + 2]; v.w += bias[n0 + 3];   // bias blocks are only 4-byte aligned
            }
            float4* dst = reinterpret_cast<float4*>(C + r * Nc + n0);
            if (ACCUM) {
              const float4 o = *dst;
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *dst = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + j;
            if (n < Nc) {
              float v = acc[i][h * 4 + j];
              if (value_row) v += bias[n];
              if (ACCUM) v += C[r * Nc + n];
              C[r * Nc + n] = v;
            }
          }
        }
      }
    }
  }
}

// partial[cta][Kc + 1][Nc] = sum over the CTA's rows of X[r][k] * dY[r][n]; row Kc = sum over the value rows (bias gradient).
// Each slice accumulates an 8 (k) x 8 (n) block per thread over its own 16-row chunks; the slices are summed through
// shared memory in a fixed order at the end.
// WROWS rows per slice chunk: 32 for large batches (half the barrier / load phases per row), 16 when the batch is small
template <int BN, int WROWS>
__global__ void __launch_bounds__(LIN_THREADS) wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dY,
                                                            float* __restrict__ partial, int64_t R, int Kc, int Nc, int G) {
  extern __shared__ __align__(16) float sm[];
  constexpr int TXN = BN / 8, TS = 8 * TXN, SL = LIN_THREADS / TS;
  constexpr int LDX = HID + 4;
  constexpr int SLICE_FLOATS = WROWS * LDX + WROWS * BN + WROWS;
  const int tid = threadIdx.x, slice = tid / TS, ts = tid % TS, tx = ts % TXN, ty = ts / TXN;
  float* Xs = sm + slice * SLICE_FLOATS;         // [WROWS][LDX]   (k zero-padded to 64)
  float* Ys = Xs + WROWS * LDX;                  // [WROWS][BN]
  float* Ind = Ys + WROWS * BN;                  // [WROWS]
  float* Acc = sm + SL * SLICE_FLOATS;           // [HID + 1][BN]
  const bool kact = ty * 4 < Kc;
  float acc[8][8], accb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    accb[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][j] = 0.f;
  }
  const bool xvec = Kc == HID, yvec = (Nc & 3) == 0;
  if (!xvec)
    for (int i = ts; i < WROWS * LDX; i += TS) Xs[i] = 0.f;      // zero padding of the narrow-input case (slice-private)
  const int64_t chunks = (R + WROWS - 1) / WROWS;
  for (int64_t ch = (int64_t)blockIdx.x * SL + slice; ch < chunks; ch += (int64_t)gridDim.x * SL) {
    const int64_t r0 = ch * WROWS;
    const int rows = (int)min((int64_t)WROWS, R - r0);
    slice_sync(slice, TS);
    if (xvec) {
      for (int i = ts; i < WROWS * (HID / 4); i += TS) {
        const int rl = i / (HID / 4), k4 = i % (HID / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < rows) v = *reinterpret_cast<const float4*>(X + (r0 + rl) * HID + 4 * k4);
        *reinterpret_cast<float4*>(Xs + rl * LDX + 4 * k4) = v;
      }
    } else {
      // narrow input (the first layer: Kc = D): only the Kc live columns are refreshed, the padding was zeroed once above
      for (int i = ts; i < WROWS * Kc; i += TS) {
        const int rl = i / Kc, k = i % Kc;
        Xs[rl * LDX + k] = rl < rows ? X[(r0 + rl) * Kc + k] : 0.f;
      }
    }
    if (yvec) {
      const int n4n = Nc >> 2;
      for (int i = ts; i < WROWS * (BN / 4); i += TS) {
        const int rl = i / (BN / 4), n4 = i % (BN / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < rows && n4 < n4n) v = *reinterpret_cast<const float4*>(dY + (r0 + rl) * Nc + 4 * n4);
        *reinterpret_cast<float4*>(Ys + rl * BN + 4 * n4) = v;
      }
    } else {
      for (int i = ts; i < WROWS * BN; i += TS) {
        const int rl = i / BN, n = i % BN;
        Ys[i] = (rl < rows && n < Nc) ? dY[(r0 + rl) * Nc + n] : 0.f;
      }
    }
    if (ts < WROWS) Ind[ts] = (ts < rows && ((r0 + ts) % G) == 0) ? 1.f : 0.f;
    slice_sync(slice, TS);
#pragma unroll 2
    for (int r = 0; r < WROWS; ++r) {
      const float4 y0 = *reinterpret_cast<const float4*>(Ys + r * BN + tx * 4);
      const float4 y1 = *reinterpret_cast<const float4*>(Ys + r * BN + BN / 2 + tx * 4);
      const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      if (ty == 0) {
        const float ind = Ind[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) accb[j] = fmaf(ind, y[j], accb[j]);
      }
      if (kact) {
        const float4 x0 = *reinterpret_cast<const float4*>(Xs + r * LDX + ty * 4);
        const float4 x1 = *reinterpret_cast<const float4*>(Xs + r * LDX + HID / 2 + ty * 4);
        const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
      }
    }
  }
  // thread block (i, j) <-> k = (i < 4 ? ty*4 + i : 32 + ty*4 + i - 4), n = (j < 4 ? tx*4 + j : BN/2 + tx*4 + j - 4)
  for (int sl = 0; sl < SL; ++sl) {
    __syncthreads();
    if (slice == sl) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = i < 4 ? ty * 4 + i : HID / 2 + ty * 4 + (i - 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4);
          float* q = Acc + k * BN + n;
          *q = sl == 0 ? acc[i][j] : *q + acc[i][j];
        }
      }
      if (ty == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4);
          float* q = Acc + HID * BN + n;
          *q = sl == 0 ? accb[j] : *q + accb[j];
        }
      }
    }
  }
  __syncthreads();
  float* out = partial + (int64_t)blockIdx.x * (Kc + 1) * Nc;
  for (int i = tid; i < (Kc + 1) * Nc; i += LIN_THREADS) {
    const int k = i / Nc, n = i % Nc;
    out[i] = Acc[(k < Kc ? k : HID) * BN + n];
  }
}

// gW[k][n] += mask * sum_cta partial[cta][k][n];  gb[n] += sum_cta partial[cta][Kc][n]   (fixed summation order)
// block = 32 outputs x 8 slices of the CTA axis: coalesced partial reads, then a shared-memory tree over the slices
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ gW,
                                                           float* __restrict__ gb, int layer, int D, int Kc, int Nc) {
  __shared__ float red[8][33];
  const int ii = threadIdx.x & 31, ci = threadIdx.x >> 5;
  const int tot = (Kc + 1) * Nc;
  const int i = blockIdx.x * 32 + ii;
  float s = 0.f;
  if (i < tot) {
    // four independent chains: the loop is a chain of dependent L2 loads otherwise (9 us per layer for 148 partial blocks)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = ci;
    for (; c + 24 < n_cta; c += 32) {
      s0 += partial[(int64_t)c * tot + i];
      s1 += partial[(int64_t)(c + 8) * tot + i];
      s2 += partial[(int64_t)(c + 16) * tot + i];
      s3 += partial[(int64_t)(c + 24) * tot + i];
    }
    for (; c < n_cta; c += 8) s0 += partial[(int64_t)c * tot + i];
    s = (s0 + s1) + (s2 + s3);
  }
  red[ci][ii] = s;
  __syncthreads();
  if (ci == 0 && i < tot) {
#pragma unroll
    for (int c = 1; c < 8; ++c) s += red[c][ii];
    const int k = i / Nc, n = i % Nc;
    if (k < Kc) {
      if (made_mask(layer, D, k, n)) gW[i] += s;
    } else {
      gb[n] += s;
    }
  }
}

// ---------------------------------------------------------------------------------------------- tanh on jets
template <int D>
__global__ void tanh_fwd_kernel(const float* __restrict__ Z, float* __restrict__ Hh, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * HID) return;
  const int64_t n = i / HID;
  const int c = (int)(i % HID);
  const Jet<D> z = jload<D>(Z, n, HID, c);
  const float t = tanhf(z.v), d1 = 1.f - t * t;
  jstore<D>(Hh, n, HID, c, junary(z, t, d1, -2.f * t * d1));
}
// Hbar -> Zbar in place
template <int D>
__global__ void tanh_bwd_kernel(const float* __restrict__ Z, float* __restrict__ Hbar, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * HID) return;
  const int64_t n = i / HID;
  const int c = (int)(i % HID);
  const Jet<D> z = jload<D>(Z, n, HID, c);
  const Jet<D> ob = jload<D>(Hbar, n, HID, c);
  const float t = tanhf(z.v), d1 = 1.f - t * t, d2 = -2.f * t * d1, d3 = -2.f * d1 * d1 - 2.f * t * d2;
  Jet<D> zb = jzero<D>();
  junary_bwd(z, d1, d2, d3, ob, zb);
  jstore<D>(Hbar, n, HID, c, zb);
}

// ---- first conditioner layer, fused: Z1 = U W1 + b1 (K = D: no GEMM to speak of) and H1 = tanh(Z1) in one pass, and in
// reverse Zbar1 = tanh'(Z1; Hbar1) (in place) together with the input adjoint Ubar += Zbar1 W1^T.  16 lanes per walker, 4 hidden
// features each: every thread holds all G jet components of its features, so the tanh jet needs no exchange; rows are written
// with 128-bit coalesced stores.  Replaces linear_kernel + tanh_fwd_kernel and tanh_bwd_kernel + a separate 64 -> D product.
template <int D>
__global__ void __launch_bounds__(256) layer1_fwd_kernel(const float* __restrict__ U, const float* __restrict__ W1, const float* __restrict__ b1,
                                                         float* __restrict__ Z, float* __restrict__ Hh, int64_t N) {
  constexpr int G = D + 2;
  __shared__ float4 Ws[D][16];
  __shared__ float4 bs[16];
  if (threadIdx.x < D * 16) Ws[threadIdx.x / 16][threadIdx.x % 16] = make_float4(W1[4 * threadIdx.x], W1[4 * threadIdx.x + 1], W1[4 * threadIdx.x + 2], W1[4 * threadIdx.x + 3]);
  if (threadIdx.x < 16) bs[threadIdx.x] = make_float4(b1[4 * threadIdx.x], b1[4 * threadIdx.x + 1], b1[4 * threadIdx.x + 2], b1[4 * threadIdx.x + 3]);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = i >> 4;
  const int q = (int)(i & 15);
  if (n >= N) return;
  float z[G][4];
#pragma unroll
  for (int c = 0; c < G; ++c) {
    float4 acc = c == 0 ? bs[q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float u = __ldg(U + (n * G + c) * D + k);
      const float4 w = Ws[k][q];
      acc.x = fmaf(u, w.x, acc.x); acc.y = fmaf(u, w.y, acc.y); acc.z = fmaf(u, w.z, acc.z); acc.w = fmaf(u, w.w, acc.w);
    }
    z[c][0] = acc.x; z[c][1] = acc.y; z[c][2] = acc.z; z[c][3] = acc.w;
    *reinterpret_cast<float4*>(Z + (n * G + c) * HID + 4 * q) = acc;
  }
  float h[G][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    Jet<D> zj;
    zj.v = z[0][j];
#pragma unroll
    for (int g = 0; g < D; ++g) zj.g[g] = z[1 + g][j];
    zj.l = z[D + 1][j];
    const float t = tanhf(zj.v), d1 = 1.f - t * t;
    const Jet<D> hj = junary(zj, t, d1, -2.f * t * d1);
    h[0][j] = hj.v;
#pragma unroll
    for (int g = 0; g < D; ++g) h[1 + g][j] = hj.g[g];
    h[D + 1][j] = hj.l;
  }
#pragma unroll
  for (int c = 0; c < G; ++c) *reinterpret_cast<float4*>(Hh + (n * G + c) * HID + 4 * q) = make_float4(h[c][0], h[c][1], h[c][2], h[c][3]);
}

// Hbar -> Zbar in place; Ubar[R][D] += Zbar W1^T when Ubar != nullptr (W1 [D][64])
template <int D>
__global__ void __launch_bounds__(256) layer1_bwd_kernel(const float* __restrict__ Z, float* __restrict__ Hbar, const float* __restrict__ W1,
                                                         float* __restrict__ Ubar, int64_t N) {
  constexpr int G = D + 2;
  __shared__ float4 Ws[D][16];
  if (threadIdx.x < D * 16) Ws[threadIdx.x / 16][threadIdx.x % 16] = make_float4(W1[4 * threadIdx.x], W1[4 * threadIdx.x + 1], W1[4 * threadIdx.x + 2], W1[4 * threadIdx.x + 3]);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = i >> 4;                           // N is padded to whole warps by the launch: every lane takes part in the shuffles
  const int q = (int)(i & 15);
  const bool ok = n < N;
  float z[G][4], ob[G][4], zb[G][4];
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const float4 a = ok ? __ldg(reinterpret_cast<const float4*>(Z + (n * G + c) * HID) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b = ok ? *(reinterpret_cast<const float4*>(Hbar + (n * G + c) * HID) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    z[c][0] = a.x; z[c][1] = a.y; z[c][2] = a.z; z[c][3] = a.w;
    ob[c][0] = b.x; ob[c][1] = b.y; ob[c][2] = b.z; ob[c][3] = b.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    Jet<D> zj, oj;
    zj.v = z[0][j]; oj.v = ob[0][j];
#pragma unroll
    for (int g = 0; g < D; ++g) { zj.g[g] = z[1 + g][j]; oj.g[g] = ob[1 + g][j]; }
    zj.l = z[D + 1][j]; oj.l = ob[D + 1][j];
    const float t = tanhf(zj.v), d1 = 1.f - t * t, d2 = -2.f * t * d1, d3 = -2.f * d1 * d1 - 2.f * t * d2;
    Jet<D> r = jzero<D>();
    junary_bwd(zj, d1, d2, d3, oj, r);
    zb[0][j] = r.v;
#pragma unroll
    for (int g = 0; g < D; ++g) zb[1 + g][j] = r.g[g];
    zb[D + 1][j] = r.l;
  }
  if (ok) {
#pragma unroll
    for (int c = 0; c < G; ++c) *(reinterpret_cast<float4*>(Hbar + (n * G + c) * HID) + q) = make_float4(zb[c][0], zb[c][1], zb[c][2], zb[c][3]);
  }
  if (Ubar) {
#pragma unroll
    for (int c = 0; c < G; ++c) {
      float p[D];
#pragma unroll
      for (int o = 0; o < D; ++o) {
        const float4 w = Ws[o][q];
        p[o] = fmaf(zb[c][0], w.x, fmaf(zb[c][1], w.y, fmaf(zb[c][2], w.z, zb[c][3] * w.w)));
      }
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1)
#pragma unroll
        for (int o = 0; o < D; ++o) p[o] += __shfl_xor_sync(0xffffffffu, p[o], m);
      if (ok && q == 0) {
#pragma unroll
        for (int o = 0; o < D; ++o) Ubar[(n * G + c) * D + o] += p[o];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- box transform (made.py:108-183)
template <int D>
__global__ void box_kernel(const float* __restrict__ x, int64_t N, float L, int coord_mean, float* __restrict__ U,
                           float* __restrict__ LD) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Jet<D> X[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    X[d] = jzero<D>();
    X[d].v = x[n * D + d];
    X[d].g[d] = 1.f;
  }
  const float tol = 1e-7f;
  Jet<D> ld = jzero<D>();
  if (coord_mean) {
    Jet<D> mean = jzero<D>();
#pragma unroll
    for (int d = 0; d < D; ++d) jacc(mean, X[d]);
    mean = jscale(mean, 1.f / (float)D);
    const Jet<D> l = jsub(mean, X[0]);
    const Jet<D> w = jsub(X[D - 1], X[0]);
    Jet<D> space = jzero<D>();
    space.v = 2.f * L;
#pragma unroll
    for (int i = 0; i < D - 1; ++i) {
      const Jet<D> diff = jsub(X[i + 1], X[i]);
      Jet<D> den = space;
      den.v += tol;
      jstore<D>(U, n, D, i, jmul(diff, jrecip(den)));
      ld = jsub(ld, jlog(den));
      space = jsub(space, diff);
    }
    Jet<D> den = jscale(w, -1.f);
    den.v = (2.f * L - w.v) + tol;
    Jet<D> num = jsub(mean, l);
    num.v = (mean.v + L) - l.v;
    jstore<D>(U, n, D, D - 1, jmul(num, jrecip(den)));
    ld = jsub(ld, jlog(den));
  } else {
    Jet<D> u0 = jscale(X[0], 1.f / (2.f * L));
    u0.v = (X[0].v + L) / (2.f * L);
    jstore<D>(U, n, D, 0, u0);
    ld.v -= logf(2.f * L);
#pragma unroll
    for (int i = 1; i < D; ++i) {
      Jet<D> den = jscale(X[i - 1], -1.f);
      den.v = (L - X[i - 1].v) + tol;
      jstore<D>(U, n, D, i, jmul(jsub(X[i], X[i - 1]), jrecip(den)));
      ld = jsub(ld, jlog(den));
    }
  }
  jstore<D>(LD, n, 1, 0, ld);
}

// ---------------------------------------------------------------------------------------------- spline heads
// One WARP per (walker, dimension): lane q owns coefficient q (P <= 32), so the per-coefficient jets live in registers,
// table rows are read 128 bytes at a time, and the sums over coefficients are warp all-reduces of whole jets.
template <int D>
__device__ __forceinline__ Jet<D> warp_sum(Jet<D> j) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    j.v += __shfl_xor_sync(0xffffffffu, j.v, o);
    j.l += __shfl_xor_sync(0xffffffffu, j.l, o);
#pragma unroll
    for (int i = 0; i < D; ++i) j.g[i] += __shfl_xor_sync(0xffffffffu, j.g[i], o);
  }
  return j;
}
template <int D>
__device__ __forceinline__ Jet<D> warp_bcast(const Jet<D>& j, int src) {
  Jet<D> o;
  o.v = __shfl_sync(0xffffffffu, j.v, src);
  o.l = __shfl_sync(0xffffffffu, j.l, src);
#pragma unroll
  for (int i = 0; i < D; ++i) o.g[i] = __shfl_sync(0xffffffffu, j.g[i], src);
  return o;
}
template <int D>
__device__ __forceinline__ Jet<D> sig_jet(const Jet<D>& o) {
  const Sig g = sigmoid_derivs(o.v);
  return junary(o, g.s, g.d1, g.d2);
}

// Rows of a walker's conditioner output as they lie in HBM: lane q reads / writes the D consecutive floats of coefficient q
// (all dimensions at once: one 128-bit access per jet row at D = 4, fully coalesced across the warp).  The head kernels used to
// run one warp per (walker, dimension) with 4-byte accesses at a stride of D floats -- a quarter of every sector used per
// request, and they were bound by those requests, not by their warp all-reduces.
template <int D>
__device__ __forceinline__ void rows_load(const float* __restrict__ base, int64_t n, int width, int lane, bool act, float (&o)[D + 2][D]) {
#pragma unroll
  for (int c = 0; c < D + 2; ++c) {
    const float* p = base + (n * (D + 2) + c) * (int64_t)width + lane * D;
    if constexpr (D == 4) {
      const float4 v = act ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
      o[c][0] = v.x; o[c][1] = v.y; o[c][2] = v.z; o[c][3] = v.w;
    } else {
#pragma unroll
      for (int dd = 0; dd < D; ++dd) o[c][dd] = act ? __ldg(p + dd) : 0.f;
    }
  }
}
template <int D>
__device__ __forceinline__ void rows_store(float* __restrict__ base, int64_t n, int width, int lane, const float (&o)[D + 2][D]) {
#pragma unroll
  for (int c = 0; c < D + 2; ++c) {
    float* p = base + (n * (D + 2) + c) * (int64_t)width + lane * D;
    if constexpr (D == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
    } else {
#pragma unroll
      for (int dd = 0; dd < D; ++dd) p[dd] = o[c][dd];
    }
  }
}
template <int D>
__device__ __forceinline__ Jet<D> rows_jet(const float (&o)[D + 2][D], int d) {
  Jet<D> j;
  j.v = o[0][d];
#pragma unroll
  for (int i = 0; i < D; ++i) j.g[i] = o[1 + i][d];
  j.l = o[D + 1][d];
  return j;
}
template <int D>
__device__ __forceinline__ void rows_put(float (&o)[D + 2][D], int d, const Jet<D>& j) {
  o[0][d] = j.v;
#pragma unroll
  for (int i = 0; i < D; ++i) o[1 + i][d] = j.g[i];
  o[D + 1][d] = j.l;
}

// ---- IMADE head (made.py:66-81)
struct HeadArgs {
  const float* O;      // [R][D*P] conditioner output jets
  const float* U;      // [R][D]   layer input jets
  const float* tab;    // dense [T][4][32]
  int64_t N;
  int P, T;
  float reg;
  float wq[WF_MAX_P];  // remove_bias scale with the boundary constraints folded in (0 at constrained ends)
  float* sav;          // [N][D][sav_floats<D>]: S1, S2, dy of the forward pass -- the reverse kernel reloads these three sums over
                       // coefficients (18 floats per walker and dimension at D = 4) instead of repeating four of its seven warp all-reduces
};
template <int D> constexpr int sav_floats() { return (3 * (D + 2) + 3) & ~3; }
template <int D>
__device__ __forceinline__ void sav_put(float (&v)[sav_floats<D>()], int k, const Jet<D>& j) {
  v[k * (D + 2)] = j.v;
#pragma unroll
  for (int i = 0; i < D; ++i) v[k * (D + 2) + 1 + i] = j.g[i];
  v[k * (D + 2) + D + 1] = j.l;
}
template <int D>
__device__ __forceinline__ Jet<D> sav_get(const float (&v)[sav_floats<D>()], int k) {
  Jet<D> j;
  j.v = v[k * (D + 2)];
#pragma unroll
  for (int i = 0; i < D; ++i) j.g[i] = v[k * (D + 2) + 1 + i];
  j.l = v[k * (D + 2) + D + 1];
  return j;
}
template <int D>
__device__ __forceinline__ void sav_store(float* __restrict__ p, const Jet<D>& a, const Jet<D>& b, const Jet<D>& c) {
  float v[sav_floats<D>()] = {};
  sav_put<D>(v, 0, a); sav_put<D>(v, 1, b); sav_put<D>(v, 2, c);
#pragma unroll
  for (int k = 0; k < sav_floats<D>() / 4; ++k) reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
template <int D>
__device__ __forceinline__ void sav_load(const float* __restrict__ p, Jet<D>& a, Jet<D>& b, Jet<D>& c) {
  float v[sav_floats<D>()];
#pragma unroll
  for (int k = 0; k < sav_floats<D>() / 4; ++k) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p) + k);
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
  }
  a = sav_get<D>(v, 0); b = sav_get<D>(v, 1); c = sav_get<D>(v, 2);
}

template <int D>
struct ImadeLane {
  Jet<D> o, s, S1, r1, b, S2, r2, c, u, B0, B1, y, dy;
  Basis4 f;
  Sig sg;               // sigmoid and its derivatives at o.v (the reverse pass reuses them)
  float wq;
  bool act;
};

// everything the forward pass of one (walker, dimension) produces, per lane
template <int D, bool RELOAD>
__device__ __forceinline__ void imade_lane_fwd(const HeadArgs& a, int64_t n, int d, int lane, const Jet<D>& o, ImadeLane<D>& L) {
  L.act = lane < a.P;
  L.wq = L.act ? a.wq[lane] : 0.f;
  L.o = o;                                             // zero on the padding lanes (rows_load)
  L.sg = sigmoid_derivs(L.o.v);
  L.s = L.act ? junary(L.o, L.sg.s, L.sg.d1, L.sg.d2) : jzero<D>();
  if (RELOAD) sav_load<D>(a.sav + (n * D + d) * sav_floats<D>(), L.S1, L.S2, L.dy);
  else L.S1 = warp_sum(L.s);
  L.r1 = jrecip(L.S1);
  L.b = jzero<D>();
  if (L.wq != 0.f) {
    L.b = jmul(L.s, L.r1);
    L.b.v += a.reg;
    L.b = jscale(L.b, L.wq);
  }
  if (!RELOAD) L.S2 = warp_sum(L.b);
  L.r2 = jrecip(L.S2);
  L.c = jmul(L.b, L.r2);
  L.u = jload<D>(a.U, n, D, d);
  const NodeIdx ni = node_index(L.u.v, a.T);
  L.f = basis4(a.tab, ni, (float)(a.T - 1), lane);     // padded table columns (lane >= P) are zero
  L.B0 = junary(L.u, L.f.f[0], L.f.f[1], L.f.f[2]);
  L.B1 = junary(L.u, L.f.f[1], L.f.f[2], L.f.f[3]);
  if (!RELOAD) {                                                   // y itself is not needed by the reverse pass
    L.y = warp_sum(jmul(L.c, L.B0));
    L.dy = warp_sum(jmul(L.c, L.B1));
    L.dy.v += LOG_TOL;
  }
}

// One WARP per walker, lane q owns coefficient q (P <= 32); the D dimensions are processed one after the other on the rows
// loaded once.  Y [R][D] receives y at the REVERSED column (the Reverse layer, bijections.py:317-347); LDC [R][D] the log-det terms
template <int D>
__global__ void __launch_bounds__(HEAD_THREADS) imade_fwd_kernel(const __grid_constant__ HeadArgs a, float* __restrict__ Y,
                                                                 float* __restrict__ LDC) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= a.N) return;
  float o[D + 2][D];
  rows_load<D>(a.O, n, D * a.P, lane, lane < a.P, o);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    ImadeLane<D> L;
    imade_lane_fwd<D, false>(a, n, d, lane, rows_jet<D>(o, d), L);
    if (lane == 0) {
      jstore<D>(Y, n, D, D - 1 - d, L.y);
      jstore<D>(LDC, n, D, d, jlog(L.dy));
      if (a.sav) sav_store<D>(a.sav + (n * D + d) * sav_floats<D>(), L.S1, L.S2, L.dy);
    }
  }
}

// Ybar [R][D]: adjoint of the NEXT layer's input (so y_d's adjoint sits at column D-1-d); LDbar [R]: adjoint of log|det|
template <int D>
__global__ void __launch_bounds__(HEAD_THREADS, 4) imade_bwd_kernel(const __grid_constant__ HeadArgs a, const float* __restrict__ Ybar,
                                                                 const float* __restrict__ LDbar, float* __restrict__ Obar,
                                                                 float* __restrict__ Ubar) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= a.N) return;
  float o[D + 2][D], ob[D + 2][D];
  rows_load<D>(a.O, n, D * a.P, lane, lane < a.P, o);
  const Jet<D> lbar = jload<D>(LDbar, n, 1, 0);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    ImadeLane<D> L;
    imade_lane_fwd<D, true>(a, n, d, lane, rows_jet<D>(o, d), L);
    const Jet<D> ybar = jload<D>(Ybar, n, D, D - 1 - d);
    Jet<D> dybar = jzero<D>();
    jlog_bwd(L.dy, lbar, dybar);
    Jet<D> cbar = jzero<D>(), B0bar = jzero<D>(), B1bar = jzero<D>(), ub = jzero<D>();
    jmul_bwd(L.B0, ybar, cbar);
    jmul_bwd(L.B1, dybar, cbar);
    jmul_bwd(L.c, ybar, B0bar);
    jmul_bwd(L.c, dybar, B1bar);
    junary_bwd(L.u, L.f.f[1], L.f.f[2], L.f.f[3], B0bar, ub);
    junary_bwd(L.u, L.f.f[2], L.f.f[3], L.f.f[3], B1bar, ub);   // table order 4 clamps to 3 (quirk Q5)
    const Jet<D> ubar = warp_sum(ub);
    Jet<D> bbar = jzero<D>(), r2b = jzero<D>();
    jmul_bwd(L.r2, cbar, bbar);
    jmul_bwd(L.b, cbar, r2b);
    const Jet<D> r2bar = warp_sum(r2b);
    Jet<D> S2bar = jzero<D>();
    jrecip_bwd(L.S2, r2bar, S2bar);
    Jet<D> sbar = jzero<D>(), r1b = jzero<D>();
    if (L.wq != 0.f) {
      jacc(bbar, S2bar);
      const Jet<D> abar = jscale(bbar, L.wq);
      jmul_bwd(L.r1, abar, sbar);
      jmul_bwd(L.s, abar, r1b);
    }
    const Jet<D> r1bar = warp_sum(r1b);
    Jet<D> S1bar = jzero<D>();
    jrecip_bwd(L.S1, r1bar, S1bar);
    Jet<D> obar = jzero<D>();
    if (L.act) {
      jacc(sbar, S1bar);
      junary_bwd(L.o, L.sg.d1, L.sg.d2, L.sg.d3, sbar, obar);
    }
    rows_put<D>(ob, d, obar);
    if (lane == 0) jstore<D>(Ubar, n, D, d, ubar);
  }
  if (lane < a.P) rows_store<D>(Obar, n, D * a.P, lane, ob);
}

// ---- prior head (wavefunctions.py:54-71)
struct PriorArgs {
  const float* O;        // [R][D*P + D]: folded third layer (see fold_prior_kernel)
  const float* U;        // [R][D]
  const float* tab;      // orthonormalised tables, dense [T][4][32]
  int64_t N;
  int P, T;
  int cons_lo, cons_hi;  // dimensions [cons_lo, cons_hi) carry the 1/sqrt(2) (model_factory.py:124-129)
};

template <int D>
struct PriorLane {
  Jet<D> cpre, S, nrm, c, uc, B0, phi;
  Basis4 f;
  float sign, scale, uraw;
  bool inside;
};

template <int D>
__device__ __forceinline__ void prior_lane_fwd(const PriorArgs& a, int64_t n, int d, int lane, const Jet<D>& cpre, PriorLane<D>& L) {
  const int P = a.P, NcF = D * P + D;
  // O is the FOLDED third layer: column q*D + d = c'_q (un-normalised B coefficients), column D*P + d = sum_p o_p, whose sign
  // the conditioner's own normalisation leaves behind (model_factory.py:69-70; the L2 normalisations cancel its magnitude)
  const float sv = a.O[n * (D + 2) * (int64_t)NcF + D * P + d];
  L.sign = sv < 0.f ? -1.f : 1.f;
  L.cpre = cpre;                                       // zero on the padding lanes (rows_load)
  L.S = warp_sum(jmul(L.cpre, L.cpre));
  L.nrm = jrsqrt(L.S);
  L.c = jmul(L.cpre, L.nrm);
  const Jet<D> u = jload<D>(a.U, n, D, d);
  L.uraw = u.v;
  L.inside = (u.v > 0.f) && (u.v < 1.f);
  L.uc = L.inside ? u : jzero<D>();
  L.uc.v = fminf(fmaxf(u.v, 0.f), 1.f);
  const NodeIdx ni = node_index(L.uc.v, a.T);
  L.f = basis4(a.tab, ni, (float)(a.T - 1), lane);
  L.B0 = junary(L.uc, L.f.f[0], L.f.f[1], L.f.f[2]);
  L.scale = L.sign * ((d >= a.cons_lo && d < a.cons_hi) ? 0.70710678118654752f : 1.f);
  L.phi = jscale(warp_sum(jmul(L.c, L.B0)), L.scale);
}

// one warp per walker, as the IMADE heads
template <int D>
__global__ void __launch_bounds__(HEAD_THREADS) prior_fwd_kernel(const __grid_constant__ PriorArgs a, float* __restrict__ PHI) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= a.N) return;
  float o[D + 2][D];
  rows_load<D>(a.O, n, D * a.P + D, lane, lane < a.P, o);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    PriorLane<D> L;
    prior_lane_fwd<D>(a, n, d, lane, rows_jet<D>(o, d), L);
    if (lane == 0) jstore<D>(PHI, n, D, d, L.phi);
  }
}

template <int D>
__global__ void __launch_bounds__(HEAD_THREADS) prior_bwd_kernel(const __grid_constant__ PriorArgs a, const float* __restrict__ PHIbar,
                                                                 float* __restrict__ Obar, float* __restrict__ Ubar) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= a.N) return;
  const int P = a.P, NcF = D * P + D;
  float o[D + 2][D], ob[D + 2][D];
  rows_load<D>(a.O, n, NcF, lane, lane < P, o);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    PriorLane<D> L;
    prior_lane_fwd<D>(a, n, d, lane, rows_jet<D>(o, d), L);
    const Jet<D> phibar = jscale(jload<D>(PHIbar, n, D, d), L.scale);
    Jet<D> cbar = jzero<D>(), B0bar = jzero<D>(), ub = jzero<D>();
    jmul_bwd(L.B0, phibar, cbar);
    jmul_bwd(L.c, phibar, B0bar);
    junary_bwd(L.uc, L.f.f[1], L.f.f[2], L.f.f[3], B0bar, ub);
    Jet<D> ucbar = warp_sum(ub);
    Jet<D> cb = jzero<D>(), nb = jzero<D>();
    jmul_bwd(L.nrm, cbar, cb);
    jmul_bwd(L.cpre, cbar, nb);
    const Jet<D> nbar = warp_sum(nb);
    Jet<D> Sbar = jzero<D>();
    jrsqrt_bwd(L.S, nbar, Sbar);
    Jet<D> t = jzero<D>();
    jmul_bwd(L.cpre, Sbar, t);            // d(c*c) = 2 * (one-sided adjoint)
    jaxpy(cb, 2.f, t);
    rows_put<D>(ob, d, cb);               // adjoint of c'_q
    if (lane == 0) {
      jstore<D>(Obar, n, NcF, D * P + d, jzero<D>());                       // the sign column has no gradient
      if (!L.inside) {
        const float keep = (L.uraw == L.uc.v) ? ucbar.v : 0.f;   // clip passes the gradient only where it did not clamp
        ucbar = jzero<D>();
        ucbar.v = keep;
      }
      jstore<D>(Ubar, n, D, d, ucbar);
    }
  }
  if (lane < P) rows_store<D>(Obar, n, NcF, lane, ob);
}

// ---------------------------------------------------------------------------------------------- psi, H psi, E_loc and the adjoint seeds
struct FinalArgs {
  const float* x;         // [N][D]
  const float* PHI;       // [R][D]
  const float* LDbox;     // [R]
  const float* LDC;       // [L][R][D]
  int64_t N, ldc_stride;
  int n_layers, n_protons;
  float protons[WF_MAX_D];
  float running_average, inv_n;
  const float* ra_dev;    // nullable: device-resident running average (CUDA-graph replays), overrides running_average
  float* PHIbar;          // [R][D]
  float* LDbar;           // [R]
  float* psi; float* hpsi; float* eloc;   // nullable [N]
  double* sums;           // nullable [4]: += {sum E, sum E^2, count, sum psi^2}
};

template <int D>
__global__ void final_kernel(const __grid_constant__ FinalArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f, e2 = 0.f, cnt = 0.f, p2 = 0.f;
  if (n < a.N) {
    Jet<D> ld = jload<D>(a.LDbox, n, 1, 0);
    for (int l = 0; l < a.n_layers; ++l)
#pragma unroll
      for (int d = 0; d < D; ++d) jacc(ld, jload<D>(a.LDC + (int64_t)l * a.ldc_stride, n, D, d));
    Jet<D> phi[D];
#pragma unroll
    for (int d = 0; d < D; ++d) phi[d] = jload<D>(a.PHI, n, D, d);
    Jet<D> prod = phi[0];
#pragma unroll
    for (int d = 1; d < D; ++d) prod = jmul(prod, phi[d]);
    const Jet<D> half = jscale(ld, 0.5f);
    const float ex = expf(half.v);
    const Jet<D> E = junary(half, ex, ex, ex);
    const Jet<D> psi = jmul(prod, E);
    // soft-Coulomb potential (physics.py:60-76)
    float xs[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xs[d] = a.x[n * D + d];
    float V = 0.f;
    for (int p = 0; p < a.n_protons; ++p)
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float r = a.protons[p] - xs[d];
        V -= rsqrtf(fmaf(r, r, 1.f));
      }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) {
        const float r = xs[i] - xs[j];
        V += rsqrtf(fmaf(r, r, 1.f));
      }
    const float hpsi = fmaf(-0.5f, psi.l, V * psi.v);
    const float eloc = hpsi / (psi.v + 1e-8f);
    if (a.psi) a.psi[n] = psi.v;
    if (a.hpsi) a.hpsi[n] = hpsi;
    if (a.eloc) a.eloc[n] = eloc;
    // every walker enters the sums, as jnp.mean does (vqmc.py:200): a non-finite E_loc makes the loss non-finite, visibly,
    // exactly as it makes the adjoint seeds below -- and hence the gradient -- non-finite (same policy as wf_local_energy)
    e = eloc; e2 = eloc * eloc; cnt = 1.f; p2 = psi.v * psi.v;
    // custom_jvp of _loss_fn_efficient (vqmc.py:202-212)
    const float ravg = a.ra_dev ? __ldg(a.ra_dev) : a.running_average;
    const float ca = 2.f * (eloc - ravg) / psi.v - hpsi / (psi.v * psi.v);
    const float cb = 1.f / psi.v;
    Jet<D> psibar = jzero<D>();
    psibar.v = (ca + V * cb) * a.inv_n;
    psibar.l = -0.5f * cb * a.inv_n;
    Jet<D> prodbar = jzero<D>(), Ebar = jzero<D>(), halfbar = jzero<D>();
    jmul_bwd(E, psibar, prodbar);
    jmul_bwd(prod, psibar, Ebar);
    junary_bwd(half, ex, ex, ex, Ebar, halfbar);
    jstore<D>(a.LDbar, n, 1, 0, jscale(halfbar, 0.5f));
#pragma unroll
    for (int d = 0; d < D; ++d) {
      Jet<D> others = jzero<D>();
      others.v = 1.f;
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (k != d) others = jmul(others, phi[k]);
      Jet<D> pb = jzero<D>();
      jmul_bwd(others, prodbar, pb);
      jstore<D>(a.PHIbar, n, D, d, pb);
    }
  }
  if (a.sums) {
    double v[4] = {(double)e, (double)e2, (double)cnt, (double)p2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if ((threadIdx.x & 31) == 0 && v[2] > 0.0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(a.sums + k, v[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- Adam (jax.example_libraries.optimizers.adam)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                            int64_t n, float lr, float b1, float b2, float eps, float c1, float c2, const int64_t* __restrict__ step_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (step_dev) {   // device-resident step counter (CUDA-graph replays): bias corrections computed here
    const float e = (float)(*step_dev + 1);
    c1 = 1.f - powf(b1, e);
    c2 = 1.f - powf(b2, e);
  }
  const float gi = g[i];
  const float mi = (1.f - b1) * gi + b1 * m[i];
  const float vi = (1.f - b2) * gi * gi + b2 * v[i];
  m[i] = mi; v[i] = vi;
  p[i] = p[i] - lr * (mi / c1) / (sqrtf(vi / c2) + eps);
}

// ---------------------------------------------------------------------------------------------- host orchestration
struct NetOff { int64_t W1, b1, W2, b2, W3, b3, zero, end; };

NetOff net_offsets(int D, int P, int64_t base) {
  NetOff o;
  o.W1 = base; o.b1 = o.W1 + (int64_t)D * HID; o.W2 = o.b1 + HID; o.b2 = o.W2 + HID * HID; o.W3 = o.b2 + HID;
  o.b3 = o.W3 + (int64_t)HID * D * P; o.zero = o.b3 + (int64_t)D * P; o.end = o.zero + (int64_t)D * P;
  return o;
}
int n_nets_of(const wf_live_model* m) { return m->n_layers + 1; }
int net_P(const wf_live_model* m, int i) { return i < m->n_layers ? m->P_I : m->P_P; }
int prior_cols(const wf_live_model* m) { return m->D * m->P_P + m->D; }     // folded prior layer: D*P coefficients + D sign sums
int max_DP(const wf_live_model* m) { const int a = m->D * m->P_I, b = prior_cols(m); return a > b ? a : b; }

int check_model(const wf_live_model* m) {
  if (!m) return WF_ERR_INVALID_ARG;
  if (m->D < 2 || m->D > 4) return WF_ERR_UNSUPPORTED;
  if (m->prior_kind != WF_KIND_B || !m->has_box || m->bc_I != 3 || m->bc_P != 3) return WF_ERR_UNSUPPORTED;
  if (m->n_layers < 0 || m->n_layers > WF_MAX_LAYERS || m->P_I < 2 || m->P_I > WF_MAX_P || m->P_P < 2 || m->P_P > WF_MAX_P)
    return WF_ERR_INVALID_ARG;
  if (max_DP(m) > MAX_W) return WF_ERR_UNSUPPORTED;
  return WF_OK;
}

constexpr int WGRAD_CTAS = 296;

struct Fixed { int64_t wm, fold, partial, img, total; };
constexpr int64_t IMG_NET = 2 * ttc::img_floats(1, 64) + ttc::img_floats(1, 128) + ttc::img_floats(2, 64);   // W2, W2^T, W3, W3^T
Fixed fixed_floats(const wf_live_model* m) {
  Fixed f;
  f.wm = 0;
  for (int i = 0; i < n_nets_of(m); ++i) f.wm += (int64_t)m->D * HID + HID * HID + (int64_t)HID * m->D * net_P(m, i);
  f.fold = 2 * (int64_t)(HID + 1) * ((prior_cols(m) + 3) & ~3);       // folded prior layer (Wf | bf) and its gradient (gWf | gbf)
  f.partial = (int64_t)WGRAD_CTAS * (HID + 1) * MAX_W;
  f.img = IMG_NET * n_nets_of(m);                                     // tensor-core operand images of the layer-2 / 3 weights
  f.total = f.wm + f.fold + f.partial + f.img;
  return f;
}
int64_t per_row_floats(const wf_live_model* m) {
  const int D = m->D, nn = n_nets_of(m), DPm = max_DP(m);
  // U[nn+1], per net Z1 H1 Z2 H2 O, LDbox, LDC[L], PHI, PHIbar, LDbar, Obar, HbarA, HbarB, Ubar x2
  return (int64_t)(nn + 1) * D + (int64_t)nn * (4 * HID + DPm) + 1 + (int64_t)m->n_layers * D + 2 * D + 1 + DPm + 2 * HID + 2 * D +
         4 * (int64_t)m->n_layers * D;                              // + the saved head sums S1, S2, dy per flow layer (>= D sav_floats / G)
}

int smem_linear(int Kc, int BN, int lrows) {
  const int KcP = (Kc + 3) & ~3, SL = LIN_THREADS / BN;
  return (KcP * BN + SL * lrows * (KcP + 4)) * (int)sizeof(float);
}
int smem_wgrad(int BN, int wrows) {
  const int SL = LIN_THREADS / BN;
  return (SL * (wrows * (HID + 4) + wrows * BN + wrows) + (HID + 1) * BN) * (int)sizeof(float);
}

template <int BN, bool T, bool A, int RM>
int launch_linear_rm(const float* Ain, const float* B, const float* bias, float* C, int64_t R, int Kc, int Nc, int G, cudaStream_t s) {
  static bool attr[WF_MAX_DEVICES] = {};                  // the opt-in is per device, not per process
  const int dev = current_device();
  if (!attr[dev]) {
    WF_CUDA(cudaFuncSetAttribute(linear_kernel<BN, T, A, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr[dev] = true;
  }
  constexpr int SL = LIN_THREADS / BN, LROWS = 8 * RM;
  const int64_t tiles = ((R + LROWS - 1) / LROWS + SL - 1) / SL;
  const int grid = (int)(tiles < 2 * num_sms() ? tiles : 2 * num_sms());
  linear_kernel<BN, T, A, RM><<<grid, LIN_THREADS, smem_linear(Kc, BN, LROWS), s>>>(Ain, B, bias, C, R, Kc, Nc, G);
  WF_LAUNCH_CHECK();
  return WF_OK;
}
template <int BN, bool T, bool A>
int launch_linear_bn(const float* Ain, const float* B, const float* bias, float* C, int64_t R, int Kc, int Nc, int G, cudaStream_t s) {
  // small batches: 16-row chunks so that the rows spread over 4x as many CTAs
  const bool small = (R + 63) / 64 < (int64_t)num_sms() * (LIN_THREADS / BN);
  return small ? launch_linear_rm<BN, T, A, 2>(Ain, B, bias, C, R, Kc, Nc, G, s)
               : launch_linear_rm<BN, T, A, 4>(Ain, B, bias, C, R, Kc, Nc, G, s);
}
template <bool T, bool A>
int launch_linear(const float* Ain, const float* B, const float* bias, float* C, int64_t R, int Kc, int Nc, int G, cudaStream_t s) {
  return Nc <= 64 ? launch_linear_bn<64, T, A>(Ain, B, bias, C, R, Kc, Nc, G, s)
                  : launch_linear_bn<128, T, A>(Ain, B, bias, C, R, Kc, Nc, G, s);
}

template <int BN, int WROWS>
int launch_wgrad_bn(const float* X, const float* dY, float* partial, int grid, int64_t R, int Kc, int Nc, int G, cudaStream_t s) {
  static bool attr[WF_MAX_DEVICES] = {};
  const int dev = current_device();
  if (!attr[dev]) {
    WF_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN, WROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr[dev] = true;
  }
  wgrad_kernel<BN, WROWS><<<grid, LIN_THREADS, smem_wgrad(BN, WROWS), s>>>(X, dY, partial, R, Kc, Nc, G);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

int launch_wgrad(const float* X, const float* dY, float* partial, float* gW, float* gb, int layer, int D, int64_t R, int Kc, int Nc,
                 int G, cudaStream_t s) {
  if (ttc::wgrad_tc_ok(R, Kc, Nc)) {             // 64-wide inputs of large batches: tensor cores (train_tc.cuh)
    int n_cta = 0;
    const int st = ttc::launch_wgrad_tc(X, dY, partial, R, Kc, Nc, G, WGRAD_CTAS, &n_cta, s);
    if (st != WF_OK) return st;
    const int tot = (Kc + 1) * Nc;
    wgrad_reduce_kernel<<<(tot + 31) / 32, 256, 0, s>>>(partial, n_cta, gW, gb, layer, D, Kc, Nc);
    WF_LAUNCH_CHECK();
    return WF_OK;
  }
  const bool small = (R + 127) / 128 < (int64_t)num_sms();
  const int wrows = small ? 16 : 32;
  const int64_t tiles = (R + 4 * wrows - 1) / (4 * wrows);
  const int grid = (int)(tiles < WGRAD_CTAS ? tiles : WGRAD_CTAS);
  int st;
  if (Nc <= 64) st = small ? launch_wgrad_bn<64, 16>(X, dY, partial, grid, R, Kc, Nc, G, s) : launch_wgrad_bn<64, 32>(X, dY, partial, grid, R, Kc, Nc, G, s);
  else st = small ? launch_wgrad_bn<128, 16>(X, dY, partial, grid, R, Kc, Nc, G, s) : launch_wgrad_bn<128, 32>(X, dY, partial, grid, R, Kc, Nc, G, s);
  if (st != WF_OK) return st;
  const int tot = (Kc + 1) * Nc;
  wgrad_reduce_kernel<<<(tot + 31) / 32, 256, 0, s>>>(partial, grid, gW, gb, layer, D, Kc, Nc);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

void fill_wq_I(const wf_live_model* m, float* w) {
  // remove_bias on ones (isplines_jax.py:196-199), then {0:0} | {0:1} zero the end coefficients (isplines_jax.py:160-176)
  const int P = m->P_I, k = m->k_I;
  for (int q = 0; q < WF_MAX_P; ++q) w[q] = q < P ? 1.f : 0.f;
  for (int i = 0; i < k; ++i) {
    const int a = i + 1, b = P - (i + 2);
    if (a >= 0 && a < P) w[a] = w[a] * (float)(i + 1) / (float)k;
    if (b >= 0 && b < P) w[b] = w[b] * (float)(i + 1) / (float)k;
  }
  w[0] = 0.f; w[P - 1] = 0.f;
}

// The weight gradients are off the critical path of the reverse pass (the adjoint chain head -> layer 3 -> tanh -> layer 2 ->
// layer 1 -> next head never reads them), so they run on a second stream, forked and joined with events -- which a CUDA-graph
// capture of the calling stream records as plain dependencies.  For large batches every kernel fills the GPU and nothing is
// gained; for the shards of a multi-GPU step (6 k walkers per GPU at 8 GPUs) the step is a chain of 70 short kernels, and the
// 24 weight-gradient launches overlap the chain.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6 * (WF_MAX_LAYERS + 1)] = {};
  bool ok = false;
};
SideStream* side_stream() {
  static SideStream ss[WF_MAX_DEVICES];
  SideStream* a = &ss[current_device()];
  if (!a->ok) {
    if (cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : a->ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    a->ok = true;
  }
  return a;
}

template <int D>
int run_chunk(const wf_live_model* m, const wf_live_tables* t, const float* params, const float* protons, int n_protons,
              const float* x, int64_t N, float running_average, const float* ra_dev, float inv_n, float* grad, float* psi, float* hpsi,
              float* eloc, double* sums, float* ws, cudaStream_t s) {
  constexpr int G = D + 2;
  const int nn = n_nets_of(m), L = m->n_layers, DPm = max_DP(m);
  const int64_t R = N * G;
  const Fixed fx = fixed_floats(m);
  float* Wm = ws;
  const int NcF = prior_cols(m), NcFp = (NcF + 3) & ~3;
  float* Wf = Wm + fx.wm;                        // [64][NcF] folded prior third layer
  float* bf = Wf + (int64_t)HID * NcFp;          // [NcF]
  float* gWf = bf + NcFp;                        // gradient of the folded layer, same shapes
  float* gbf = gWf + (int64_t)HID * NcFp;
  float* partial = Wm + fx.wm + fx.fold;
  float* img = partial + fx.partial;
  float* p = img + fx.img;
  auto take = [&](int64_t n) { float* q = p; p += (n + 3) & ~(int64_t)3; return q; };   // keeps every array 16-byte aligned
  float* U[WF_MAX_LAYERS + 2];
  for (int i = 0; i <= nn; ++i) U[i] = take(R * D);
  float *Z1[WF_MAX_LAYERS + 1], *H1[WF_MAX_LAYERS + 1], *Z2[WF_MAX_LAYERS + 1], *H2[WF_MAX_LAYERS + 1], *O[WF_MAX_LAYERS + 1];
  for (int i = 0; i < nn; ++i) { Z1[i] = take(R * HID); H1[i] = take(R * HID); Z2[i] = take(R * HID); H2[i] = take(R * HID); O[i] = take(R * DPm); }
  float* LDbox = take(R);
  float* LDC = take((int64_t)L * R * D);
  float* PHI = take(R * D);
  float* PHIbar = take(R * D);
  float* LDbar = take(R);
  float* Obar = take(R * DPm);
  float* HbA = take(R * HID);
  float* HbB = take(R * HID);
  float* Ub[2] = {take(R * D), take(R * D)};
  float* SAV = take((int64_t)L * N * D * sav_floats<D>());

  // masked weights (model_factory.py:31-33)
  NetOff off[WF_MAX_LAYERS + 1];
  float *W1m[WF_MAX_LAYERS + 1], *W2m[WF_MAX_LAYERS + 1], *W3m[WF_MAX_LAYERS + 1];
  {
    int64_t base = 0;
    float* w = Wm;
    MaskJobs jobs;
    for (int i = 0; i < nn; ++i) {
      const int P = net_P(m, i), DP = D * P;
      off[i] = net_offsets(D, P, base);
      base = off[i].end;
      W1m[i] = w; w += D * HID;
      W2m[i] = w; w += HID * HID;
      W3m[i] = w; w += HID * DP;
      jobs.src[3 * i] = off[i].W1; jobs.dst[3 * i] = W1m[i] - Wm; jobs.K[3 * i] = D; jobs.N[3 * i] = HID;
      jobs.src[3 * i + 1] = off[i].W2; jobs.dst[3 * i + 1] = W2m[i] - Wm; jobs.K[3 * i + 1] = HID; jobs.N[3 * i + 1] = HID;
      jobs.src[3 * i + 2] = off[i].W3; jobs.dst[3 * i + 2] = W3m[i] - Wm; jobs.K[3 * i + 2] = HID; jobs.N[3 * i + 2] = DP;
    }
    mask_weights_kernel<<<dim3(8, 3 * nn), 256, 0, s>>>(params, Wm, jobs, D);
    fold_prior_kernel<<<((HID + 1) * NcF + 127) / 128, 128, 0, s>>>(W3m[nn - 1], params + off[nn - 1].b3, t->ob_to_b, D, m->P_P, Wf, bf);
    WF_LAUNCH_CHECK();
  }
  // tensor-core operand images (train_tc.cuh) of the layer-2 / 3 weights and their transposes, once per call
  const bool tc = ttc::lin_tc_ok(R, HID, HID);
  float *iW2f[WF_MAX_LAYERS + 1], *iW2b[WF_MAX_LAYERS + 1], *iW3f[WF_MAX_LAYERS + 1], *iW3b[WF_MAX_LAYERS + 1];
  if (tc) {
    ttc::PackJobs pj;
    int nj = 0;
    auto job = [&](const float* src, float* dst, int Kc, int Nc, int trans) {
      pj.src[nj] = src; pj.dst[nj] = dst; pj.Kc[nj] = Kc; pj.Nc[nj] = Nc; pj.KP[nj] = ttc::lin_kp(Kc); pj.NT[nj] = ttc::lin_nt(Nc);
      pj.trans[nj] = trans; ++nj;
    };
    for (int i = 0; i < nn; ++i) {
      float* b = img + IMG_NET * i;
      iW2f[i] = b; iW2b[i] = b + ttc::img_floats(1, 64); iW3f[i] = iW2b[i] + ttc::img_floats(1, 64); iW3b[i] = iW3f[i] + ttc::img_floats(1, 128);
      const float* W3 = i < L ? W3m[i] : Wf;
      const int N3 = i < L ? D * net_P(m, i) : NcF;
      job(W2m[i], iW2f[i], HID, HID, 0);
      job(W2m[i], iW2b[i], HID, HID, 1);
      if (ttc::lin_tc_ok(R, HID, N3)) job(W3, iW3f[i], HID, N3, 0);
      if (ttc::lin_tc_ok(R, N3, HID)) job(W3, iW3b[i], N3, HID, 1);
    }
    ttc::pack_b_kernel<<<dim3(8, nj), 256, 0, s>>>(pj);
    WF_LAUNCH_CHECK();
  }
  // C = A B (+ bias): tensor cores for the 64-wide layers of large batches, linear_kernel otherwise
  auto lin_fwd = [&](const float* A, const float* B, const float* im, const float* bias, float* C, int Kc, int Nc) {
    if (tc && ttc::lin_tc_ok(R, Kc, Nc)) return ttc::launch_lin_tc(A, im, bias, C, R, Kc, Nc, G, s);
    return launch_linear<false, false>(A, B, bias, C, R, Kc, Nc, G, s);
  };
  auto lin_bwd = [&](const float* A, const float* B, const float* im, float* C, int Kc, int Nc) {
    if (tc && ttc::lin_tc_ok(R, Kc, Nc)) return ttc::launch_lin_tc(A, im, nullptr, C, R, Kc, Nc, G, s);
    return launch_linear<true, false>(A, B, nullptr, C, R, Kc, Nc, G, s);
  };

  HeadArgs ha;
  ha.tab = t->dense_I; ha.N = N; ha.P = m->P_I; ha.T = m->T; ha.reg = m->reg;
  fill_wq_I(m, ha.wq);
  PriorArgs pa;
  pa.tab = t->dense_P; pa.N = N; pa.P = m->P_P; pa.T = m->T;
  pa.cons_lo = m->coord_mean ? 0 : 1;
  pa.cons_hi = m->coord_mean ? D - 1 : D;

  const int eb = 256;
  const int64_t nh = N * HID;
  const int hbw = (int)((N * 32 + HEAD_THREADS - 1) / HEAD_THREADS);   // spline heads: one warp per walker

  // ---------------- forward
  box_kernel<D><<<(int)((N + 127) / 128), 128, 0, s>>>(x, N, m->box, m->coord_mean, U[0], LDbox);
  WF_LAUNCH_CHECK();
  for (int i = 0; i < nn; ++i) {
    const int P = net_P(m, i), DP = D * P;
    int st;
    layer1_fwd_kernel<D><<<(int)((N * 16 + 255) / 256), 256, 0, s>>>(U[i], W1m[i], params + off[i].b1, Z1[i], H1[i], N);
    if ((st = lin_fwd(H1[i], W2m[i], iW2f[i], params + off[i].b2, Z2[i], HID, HID)) != WF_OK) return st;
    tanh_fwd_kernel<D><<<(int)((nh + eb - 1) / eb), eb, 0, s>>>(Z2[i], H2[i], N);
    if (i < L) st = lin_fwd(H2[i], W3m[i], iW3f[i], params + off[i].b3, O[i], HID, DP);
    else st = lin_fwd(H2[i], Wf, iW3f[i], bf, O[i], HID, NcF);                              // folded prior layer
    if (st != WF_OK) return st;
    if (i < L) {
      ha.O = O[i]; ha.U = U[i]; ha.sav = grad ? SAV + (int64_t)i * N * D * sav_floats<D>() : nullptr;
      imade_fwd_kernel<D><<<hbw, HEAD_THREADS, 0, s>>>(ha, U[i + 1], LDC + (int64_t)i * R * D);
    } else {
      pa.O = O[i]; pa.U = U[i];
      prior_fwd_kernel<D><<<hbw, HEAD_THREADS, 0, s>>>(pa, PHI);
    }
    WF_LAUNCH_CHECK();
  }
  FinalArgs fa;
  fa.x = x; fa.PHI = PHI; fa.LDbox = LDbox; fa.LDC = LDC; fa.N = N; fa.ldc_stride = R * D; fa.n_layers = L;
  fa.n_protons = n_protons;
  for (int i = 0; i < WF_MAX_D; ++i) fa.protons[i] = i < n_protons ? protons[i] : 0.f;
  fa.running_average = running_average; fa.inv_n = inv_n; fa.ra_dev = ra_dev;
  fa.PHIbar = PHIbar; fa.LDbar = LDbar; fa.psi = psi; fa.hpsi = hpsi; fa.eloc = eloc; fa.sums = sums;
  final_kernel<D><<<(int)((N + 127) / 128), 128, 0, s>>>(fa);
  WF_LAUNCH_CHECK();
  if (!grad) return WF_OK;

  // ---------------- backward
  SideStream* sd = side_stream();
  if (!sd) return (int)cudaErrorUnknown;
  cudaStream_t s2 = sd->stream;
  int ne = 0;
  // fork: s2 continues after everything issued on s so far; done: an event on s2 that s has to wait for later
  auto fork = [&]() { cudaEvent_t e = sd->ev[ne++]; WF_CUDA(cudaEventRecord(e, s)); WF_CUDA(cudaStreamWaitEvent(s2, e, 0)); return WF_OK; };
  auto done = [&](cudaEvent_t* out) { cudaEvent_t e = sd->ev[ne++]; WF_CUDA(cudaEventRecord(e, s2)); *out = e; return WF_OK; };
  cudaEvent_t free_Obar = nullptr, free_HbA = nullptr, free_HbB = nullptr;      // s2 has finished reading the buffer (previous net)
  int cur = 0;
  for (int i = nn - 1; i >= 0; --i) {
    const int P = net_P(m, i), DP = D * P;
    float* Ucur = Ub[cur];
    int st;
    if (free_Obar) WF_CUDA(cudaStreamWaitEvent(s, free_Obar, 0));
    if (i == L) {
      pa.O = O[i]; pa.U = U[i];
      prior_bwd_kernel<D><<<hbw, HEAD_THREADS, 0, s>>>(pa, PHIbar, Obar, Ucur);
    } else {
      ha.O = O[i]; ha.U = U[i]; ha.sav = SAV + (int64_t)i * N * D * sav_floats<D>();
      imade_bwd_kernel<D><<<hbw, HEAD_THREADS, 0, s>>>(ha, Ub[cur ^ 1], LDbar, Obar, Ucur);
    }
    WF_LAUNCH_CHECK();
    // ---- layer 3: weight gradient on s2, input adjoint on s
    if ((st = fork()) != WF_OK) return st;
    if (i == L) {
      // folded prior layer: gradient w.r.t. (Wf | bf) without a mask, then mapped back to (W3 | b3) through F^T and the MADE mask
      WF_CUDA(cudaMemsetAsync(gWf, 0, sizeof(float) * (size_t)(HID + 1) * NcFp, s2));   // gWf | gbf
      if ((st = launch_wgrad(H2[i], Obar, partial, gWf, gbf, 0, D, R, HID, NcF, G, s2)) != WF_OK) return st;
      unfold_prior_grad_kernel<<<((HID + 1) * DP + 127) / 128, 128, 0, s2>>>(gWf, gbf, t->ob_to_b, D, P, grad + off[i].W3, grad + off[i].b3);
      WF_LAUNCH_CHECK();
    } else {
      if ((st = launch_wgrad(H2[i], Obar, partial, grad + off[i].W3, grad + off[i].b3, 3, D, R, HID, DP, G, s2)) != WF_OK) return st;
    }
    if ((st = done(&free_Obar)) != WF_OK) return st;
    if (free_HbA) WF_CUDA(cudaStreamWaitEvent(s, free_HbA, 0));
    if (i == L) st = lin_bwd(Obar, Wf, iW3b[i], HbA, NcF, HID);
    else st = lin_bwd(Obar, W3m[i], iW3b[i], HbA, DP, HID);
    if (st != WF_OK) return st;
    tanh_bwd_kernel<D><<<(int)((nh + eb - 1) / eb), eb, 0, s>>>(Z2[i], HbA, N);
    WF_LAUNCH_CHECK();
    // ---- layer 2
    if ((st = fork()) != WF_OK) return st;
    if ((st = launch_wgrad(H1[i], HbA, partial, grad + off[i].W2, grad + off[i].b2, 2, D, R, HID, HID, G, s2)) != WF_OK) return st;
    if ((st = done(&free_HbA)) != WF_OK) return st;
    if (free_HbB) WF_CUDA(cudaStreamWaitEvent(s, free_HbB, 0));
    if ((st = lin_bwd(HbA, W2m[i], iW2b[i], HbB, HID, HID)) != WF_OK) return st;
    layer1_bwd_kernel<D><<<(int)((N * 16 + 255) / 256), 256, 0, s>>>(Z1[i], HbB, W1m[i], i > 0 ? Ucur : nullptr, N);
    WF_LAUNCH_CHECK();
    // ---- layer 1
    if ((st = fork()) != WF_OK) return st;
    if ((st = launch_wgrad(U[i], HbB, partial, grad + off[i].W1, grad + off[i].b1, 1, D, R, D, HID, G, s2)) != WF_OK) return st;
    if ((st = done(&free_HbB)) != WF_OK) return st;
    cur ^= 1;
  }
  if (free_HbB) WF_CUDA(cudaStreamWaitEvent(s, free_HbB, 0));       // join: s2 is in order, its last event covers everything on it
  return WF_OK;
}

}  // namespace

extern "C" int64_t wf_vqmc_param_floats(const wf_live_model* m) {
  if (check_model(m) != WF_OK) return -1;
  int64_t base = 0;
  for (int i = 0; i < n_nets_of(m); ++i) base = net_offsets(m->D, net_P(m, i), base).end;
  return base;
}

extern "C" int64_t wf_vqmc_grad_workspace_floats(const wf_live_model* m, int64_t walkers) {
  if (check_model(m) != WF_OK || walkers < 0) return -1;
  return fixed_floats(m).total + walkers * (m->D + 2) * per_row_floats(m) + 512;
}

extern "C" int wf_vqmc_loss_grad(const wf_live_model* m, const wf_live_tables* t, const float* params, const float* protons_host,
                                 int n_protons, const float* x, int64_t N, float running_average, const float* running_average_dev,
                                 float inv_n_total, float* grad, float* psi, float* hpsi, float* eloc, double* sums, float* workspace,
                                 int64_t workspace_floats, void* stream) {
  const int st = check_model(m);
  if (st != WF_OK) return st;
  if (N == 0) return WF_OK;
  if (!t || !t->dense_I || !t->dense_P || !t->ob_to_b || !params || !x || !workspace || N < 0) return WF_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return WF_ERR_INVALID_ARG;
  if (n_protons < 0 || n_protons > WF_MAX_D || (n_protons > 0 && !protons_host)) return WF_ERR_INVALID_ARG;
  const int64_t avail = workspace_floats - fixed_floats(m).total - 512;
  const int64_t per_walker = (m->D + 2) * per_row_floats(m);
  int64_t chunk = avail / per_walker;
  if (chunk < 1) return WF_ERR_INVALID_ARG;
  if (chunk > N) chunk = N;
  cudaStream_t s = (cudaStream_t)stream;
  for (int64_t lo = 0; lo < N; lo += chunk) {
    const int64_t n = (N - lo) < chunk ? (N - lo) : chunk;
    const float* xc = x + lo * m->D;
    float* pc = psi ? psi + lo : nullptr;
    float* hc = hpsi ? hpsi + lo : nullptr;
    float* ec = eloc ? eloc + lo : nullptr;
    int r;
    switch (m->D) {
      case 2: r = run_chunk<2>(m, t, params, protons_host, n_protons, xc, n, running_average, running_average_dev, inv_n_total, grad, pc, hc, ec, sums, workspace, s); break;
      case 3: r = run_chunk<3>(m, t, params, protons_host, n_protons, xc, n, running_average, running_average_dev, inv_n_total, grad, pc, hc, ec, sums, workspace, s); break;
      default: r = run_chunk<4>(m, t, params, protons_host, n_protons, xc, n, running_average, running_average_dev, inv_n_total, grad, pc, hc, ec, sums, workspace, s); break;
    }
    if (r != WF_OK) return r;
  }
  return WF_OK;
}

extern "C" int wf_adam_step(float* params, float* m, float* v, const float* grad, int64_t n, int64_t step, const int64_t* step_dev,
                            float lr, float b1, float b2, float eps, void* stream) {
  if (n == 0) return WF_OK;
  if (!params || !m || !v || !grad || n < 0 || step < 0) return WF_ERR_INVALID_ARG;
  const float c1 = 1.f - powf(b1, (float)(step + 1)), c2 = 1.f - powf(b2, (float)(step + 1));
  adam_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, m, v, grad, n, lr, b1, b2, eps, c1, c2, step_dev);
  WF_LAUNCH_CHECK();
  return WF_OK;
}
