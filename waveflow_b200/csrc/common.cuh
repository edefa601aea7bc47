// Shared device/host helpers for the waveflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/waveflow_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "waveflow_b200 kernels target sm_100a (B200) only"
#endif

#define WF_LAUNCH_CHECK()                               \
  do {                                                  \
    cudaError_t e__ = cudaGetLastError();               \
    if (e__ != cudaSuccess) return (int)e__;            \
  } while (0)

#define WF_CUDA(call)                                   \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return (int)e__;            \
  } while (0)

namespace wf {

constexpr float LOG_TOL = 1e-7f;   // made.py:79, wavefunctions.py:34, distributions.py:140

constexpr int WF_MAX_DEVICES = 64;

// ordinal of the current device, clamped into the per-device caches below
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < WF_MAX_DEVICES ? dev : WF_MAX_DEVICES - 1;
}

inline int num_sms() {
  static int n[WF_MAX_DEVICES] = {};     // per device: one process may drive several GPUs
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

// ---------------------------------------------------------------------------------------------- mbarrier + bulk copy (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WF_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WF_DONE_%=;\n\t"
      "bra WF_WAIT_%=;\n\t"
      "WF_DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared (SASS: UBLKCP), completion signalled on an mbarrier.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// streaming (evict-first) global accesses for data touched exactly once
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------- table index (A3)
// Reference: x_l = int32(floor(x*(T-1))), x_r = int32(ceil(x*(T-1))) (isplines_jax.py:47-49); JAX gather semantics:
// negative indices wrap once (Python style), then everything clamps into [0, T-1].
struct NodeIdx {
  int l, r;   // gather rows after wrap + clamp
  float dx;   // x - x_l/(T-1) with the RAW x_l (isplines_jax.py:54)
};
__device__ __forceinline__ NodeIdx node_index(float x, int T) {
  const float np_ = (float)(T - 1);
  const float xs = x * np_;
  const float fl = floorf(xs), ce = ceilf(xs);
  // saturating float->int conversion (cvt.rzi.s32.f32 saturates; NaN -> 0)
  int il = __float2int_rz(fl), ir = __float2int_rz(ce);
  NodeIdx n;
  n.dx = x - (float)il / np_;
  il = il < 0 ? il + T : il;
  ir = ir < 0 ? ir + T : ir;
  n.l = min(max(il, 0), T - 1);
  n.r = min(max(ir, 0), T - 1);
  return n;
}

__device__ __forceinline__ float lerp_tab(float yl, float yr, float np_, float dx) {
  // y_l + ((y_r - y_l) * n_points) * dx   (isplines_jax.py:55-56)
  return fmaf((yr - yl) * np_, dx, yl);
}

// tanh / sigmoid through ex2.approx + rcp.approx (about 1e-7 absolute error, measured in tests/test_gpu_live.py):
// the accurate libdevice versions cost ~3x the issue slots
__device__ __forceinline__ float fast_tanh(float x) {
  const float xc = fminf(fmaxf(x, -15.f), 15.f);
  const float e = __expf(2.f * xc);
  return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ float fast_sigmoid(float x) {
  return __fdividef(1.f, 1.f + __expf(-x));
}

// ---------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
// two uniforms in [0, 1) for (seed, sample, column, attempt)
__device__ __forceinline__ void philox_uniform2(uint64_t seed, uint64_t sample, uint32_t col, uint32_t attempt, float& a, float& b) {
  uint32_t c[4] = {(uint32_t)sample, (uint32_t)(sample >> 32), col, attempt};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  a = (float)(c[0] >> 8) * (1.0f / 16777216.0f);
  b = (float)(c[1] >> 8) * (1.0f / 16777216.0f);
}

}  // namespace wf
