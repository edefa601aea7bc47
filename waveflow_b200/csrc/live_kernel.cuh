// Persistent fused kernel for the live path (see live_device.cuh).  One CTA per SM; the conditioner weights stream through
// a two-slot shared-memory ring, each net fetched by one cp.async.bulk (TMA) while the previous one is being used.
#pragma once
#include "live_device.cuh"

namespace wf {

template <int D>
__device__ __forceinline__ float soft_coulomb(const float (&xs)[D], const float* protons, int n_protons) {
  // utils/physics.py:66-71
  float pe = 0.f;
  for (int p = 0; p < n_protons; ++p)
#pragma unroll
    for (int e = 0; e < D; ++e) { const float d = protons[p] - xs[e]; pe += 1.f / sqrtf(1.f + d * d); }
  float ee = 0.f;
#pragma unroll
  for (int i = 1; i < D; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) { const float d = xs[i] - xs[j]; ee += 1.f / sqrtf(1.f + d * d); }
  return ee - pe;
}

// Shared memory:  2 x conditioner net (double buffer, filled by cp.async.bulk / TMA one net ahead of the compute)
//                 | ob_to_b [32][32] | scratch [64][LIVE_THREADS] | reduction buffer | 2 mbarriers
struct LiveSmem {
  static __host__ __device__ size_t net_bytes(int D) { return (size_t)net_floats(D) * sizeof(float); }
  static __host__ __device__ size_t total(int D) {
    return 2 * net_bytes(D) + (size_t)WF_MAX_P * WF_MAX_P * sizeof(float) + (size_t)LIVE_SCRATCH * LIVE_THREADS * sizeof(float) +
           4 * (LIVE_THREADS / 32) * sizeof(double) + 2 * sizeof(uint64_t);
  }
};

template <int D, bool LAP>
__global__ void __launch_bounds__(LIVE_THREADS, 1) live_kernel(const __grid_constant__ LiveParams P) {
  using C = Ctx<D, LAP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NETF = net_floats(D);
  const wf_live_model& M = P.m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* nets_s = reinterpret_cast<float*>(smem_raw);                         // [2][NETF]
  float* ob_s = nets_s + 2 * NETF;                                            // [32][32], B prior only
  float* scratch = ob_s + WF_MAX_P * WF_MAX_P;                                // [64][LIVE_THREADS]
  double* red_s = reinterpret_cast<double*>(scratch + LIVE_SCRATCH * LIVE_THREADS);   // [4][warps]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red_s + 4 * (LIVE_THREADS / 32));

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (M.prior_kind == WF_KIND_B) {
    for (int i = tid; i < WF_MAX_P * WF_MAX_P; i += LIVE_THREADS) {
      const int r = i / WF_MAX_P, c = i % WF_MAX_P;
      ob_s[i] = (r < M.P_P && c < M.P_P) ? P.ob_to_b[r * M.P_P + c] : 0.f;
    }
  }
  __syncthreads();

  C cx;
  cx.init(lane);
  const Scratch S{scratch + tid};
  // small problems are spread over more SMs with fewer working warps per CTA (latency, not throughput, matters there)
  const int WPB = P.warps_per_cta * C::WPW;             // walkers per CTA batch
  const bool warp_works = warp < P.warps_per_cta;
  const int64_t n_batches = (P.N + WPB - 1) / WPB;
  const int n_nets = P.n_nets;
  const bool has_prior_net = M.prior_kind == WF_KIND_B || M.prior_kind == WF_KIND_M;
  double accE = 0.0, accE2 = 0.0, accN = 0.0, accP2 = 0.0;

  // weight pipeline: net g (a flat counter over batches x nets) lives in buffer g & 1
  const uint32_t net_bytes = (uint32_t)(NETF * sizeof(float));
  auto issue_net = [&](int64_t g) {
    const int b = (int)(g & 1);
    mbar_expect_tx(&bars[b], net_bytes);
    bulk_g2s(nets_s + (size_t)b * NETF, P.weights + (size_t)(g % n_nets) * NETF, net_bytes, &bars[b]);
  };
  const int64_t my_batches = blockIdx.x < n_batches ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t g_total = my_batches * n_nets;
  int64_t g = 0;
  if (tid == 0 && g_total > 0) issue_net(0);

  for (int64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int64_t w_raw = batch * WPB + (int64_t)warp * C::WPW + cx.slot;
    const bool lane_live = (!LAP || lane < C::WPW * C::G) && w_raw < P.N;
    const int64_t w = w_raw < P.N ? w_raw : P.N - 1;

    float xs[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xs[d] = __ldg(P.x + w * D + d);

    // ---------------------------------------------------------------- box transform (made.py:118-137,156-183)
    float us[D];        // 1-register bundles of the current layer input
    J ld = cx.constant(0.f);
    box_transform<D, LAP>(cx, M, xs, us, ld);

    float uout[D];      // flow output (value), for the `u` result
#pragma unroll
    for (int d = 0; d < D; ++d) uout[d] = 0.f;
    if (M.n_layers == 0) {
#pragma unroll
      for (int d = 0; d < D; ++d) uout[d] = cx.bv(us[d]);
    }
    J psi = cx.constant(1.f);
    float lp = 0.f;

    // ------------------------------------------- conditioner nets: (IMADE, Reverse) x L (made.py:66-81), then the prior
#pragma unroll 1
    for (int net_idx = 0; net_idx < n_nets; ++net_idx, ++g) {
      const bool is_prior = has_prior_net && net_idx == n_nets - 1;
      // every warp is done with the other buffer (it held net g-1): refill it with net g+1, then wait for net g
      __syncthreads();
      if (tid == 0 && g + 1 < g_total) issue_net(g + 1);
      mbar_wait(&bars[g & 1], (uint32_t)((g >> 1) & 1));
      const float* net = nets_s + (size_t)(g & 1) * NETF;
      if (!warp_works) continue;               // idle warps only take part in the CTA-wide weight hand-over

      float h[WF_HIDDEN];
      mlp_hidden<D, LAP>(cx, net, us, S, h);
      float ys[D];
#pragma unroll
      for (int d = 0; d < D; ++d) ys[d] = 0.f;
#pragma unroll 1
      for (int d = 0; d < D; ++d) {
        mlp_out<D, LAP>(cx, net, d, h, S);
        float xd = us[0];
#pragma unroll
        for (int dd = 1; dd < D; ++dd) xd = (d == dd) ? us[dd] : xd;
        const float xv = cx.bv(xd);
        if (!is_prior) {
          J y, dy;
          sigmoid_spline<D, LAP, 2, true>(cx, S, M.P_I, P.wq_I, M.reg, P.rec_I, P.lo_I, P.tab_I, M.T, xd, xv, y, dy);
          const float yf = cx.fold(y);
#pragma unroll
          for (int dd = 0; dd < D; ++dd) ys[dd] = (d == dd) ? yf : ys[dd];
          ld = cx.add(ld, cx.log(cx.addc(dy, LOG_TOL)));
        } else {
          // constrained dimensions (model_factory.py:124-129): 'mean' -> 0..D-2, 'first' -> 1..D-1
          const bool cons = M.coord_mean ? (d < D - 1) : (d >= 1);
          if (M.prior_kind == WF_KIND_B) {
            J phi = bprior_factor<D, LAP>(cx, S, M.P_P, P.wq_P, ob_s, P.tab_P, M.T, xd, xv, (M.bc_P & 4) != 0);
            if (!LAP) {
              float pr = phi.v * phi.v;
              if (cons) pr = pr / 2.f;
              lp += logf(pr + LOG_TOL);
            }
            if (cons) phi = cx.scale(phi, 0.70710678118654752f);
            psi = cx.mul(psi, phi);
          } else {
            const float xc = fminf(fmaxf(xv, 0.f), 1.f);
            const float xdc = (xv > 0.f && xv < 1.f) ? xd : 0.f;
            J y, dy;
            sigmoid_spline<D, LAP, 1, false>(cx, S, M.P_P, P.wq_P, 0.f, P.rec_P, P.lo_P, P.tab_P, M.T, xdc, xc, y, dy);
            lp += logf(y.v + LOG_TOL);
          }
        }
      }
      if (!is_prior) {
#pragma unroll
        for (int d = 0; d < D; ++d) us[d] = ys[D - 1 - d];      // Reverse (bijections.py:336-345)
        if (net_idx == M.n_layers - 1) {
#pragma unroll
          for (int d = 0; d < D; ++d) uout[d] = cx.bv(us[d]);
        }
      }
    }
    // psi = prod phi * exp(0.5 log_det)   (wavefunctions.py:67-71)
    if (M.prior_kind == WF_KIND_B) psi = cx.mul(psi, cx.exp(cx.scale(ld, 0.5f)));
    const float psi1 = cx.fold(psi);

    // ---------------------------------------------------------------- outputs
    if (warp_works && lane_live && cx.is_v) {
      if (P.u) {
#pragma unroll
        for (int d = 0; d < D; ++d) P.u[w * D + d] = uout[d];
      }
      if (P.logdet) P.logdet[w] = ld.v;
      if (P.logpdf) P.logpdf[w] = lp + ld.v;
      if (P.psi) P.psi[w] = psi.v;
    }
    if constexpr (LAP) {
      const float lapv = __shfl_sync(FULL, psi1, cx.gbase + D + 1);
      if (warp_works && lane_live && cx.is_g && P.grad) P.grad[w * D + (cx.comp - 1)] = psi1;
      if (warp_works && lane_live && cx.is_v) {
        const float V = soft_coulomb<D>(xs, P.protons, P.n_protons);
        const float hp = fmaf(-0.5f, lapv, V * psi.v);           // physics.py:84
        const float el = hp / (psi.v + 1e-8f);                  // vqmc.py:200
        if (P.lap) P.lap[w] = lapv;
        if (P.hpsi) P.hpsi[w] = hp;
        if (P.eloc) P.eloc[w] = el;
        accE += (double)el; accE2 += (double)el * (double)el; accN += 1.0; accP2 += (double)psi.v * (double)psi.v;
      }
    }
  }

  if constexpr (LAP) {
    if (P.sums) {
      double v[4] = {accE, accE2, accN, accP2};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], off);
        if (lane == 0) red_s[k * (LIVE_THREADS / 32) + warp] = v[k];
      }
      __syncthreads();
      if (tid < 4) {
        double s = 0.0;
        for (int i = 0; i < LIVE_THREADS / 32; ++i) s += red_s[tid * (LIVE_THREADS / 32) + i];
        atomicAdd(P.sums + tid, s);
      }
    }
  }
}

template <int D, bool LAP>
int launch_live(LiveParams& P, cudaStream_t s) {
  const size_t smem = LiveSmem::total(D);
  if (smem > 227 * 1024) return WF_ERR_UNSUPPORTED;
  const int wpw = LAP ? 32 / (D + 2) : 32;
  const int64_t warp_tasks = (P.N + wpw - 1) / wpw;
  int wpc = (int)((warp_tasks + num_sms() - 1) / num_sms());
  if (wpc < 1) wpc = 1;
  if (wpc > LIVE_THREADS / 32) wpc = LIVE_THREADS / 32;
  P.warps_per_cta = wpc;
  const int64_t wpb = (int64_t)wpc * wpw;
  const int64_t n_batches = (P.N + wpb - 1) / wpb;
  const int blocks = (int)(n_batches < num_sms() ? n_batches : num_sms());
  WF_CUDA(cudaFuncSetAttribute(live_kernel<D, LAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  live_kernel<D, LAP><<<blocks, LIVE_THREADS, smem, s>>>(P);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

#define WF_DECL_LIVE(D) \
  int launch_live_d##D##_lap0(LiveParams& P, cudaStream_t s); \
  int launch_live_d##D##_lap1(LiveParams& P, cudaStream_t s);
WF_DECL_LIVE(2) WF_DECL_LIVE(3) WF_DECL_LIVE(4)
#undef WF_DECL_LIVE

}  // namespace wf
