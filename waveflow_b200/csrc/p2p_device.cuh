// Device side of the one-shot estimator all-reduce over NVLink peer memory (see p2p.cu for the protocol).  A header so
// that the same code runs (a) as the stand-alone 32-thread kernel, (b) in the tail of the local-energy kernels (the last
// CTA to finish does the exchange: no second launch), and (c) in the single-launch self-test that emulates all ranks as
// the warps of ONE CTA (mutually waiting kernels must never be separate launches on one GPU).
#pragma once
#include "common.cuh"

namespace wf {
namespace p2p {

constexpr int SLOT_DOUBLES = 8;            // { double v[4]; uint64 flag; pad[3] }
constexpr int HEADER_DOUBLES = 8;          // { uint64 sticky_error; pad[7] } after the 2 * world slots
constexpr int MAX_WORLD = 2 * WF_MAX_D;
constexpr long long TIMEOUT_CYCLES = 4000000000ll;   // ~2 s: far beyond legitimate skew, short of looking like a hung GPU

__host__ __device__ constexpr int64_t buffer_doubles(int world) { return (int64_t)2 * world * SLOT_DOUBLES + HEADER_DOUBLES; }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct Args {
  const unsigned long long* peer_bufs;     // device array of `world` buffer addresses (entry r = rank r's buffer)
  int rank, world;
  unsigned long long step;                 // 1, 2, 3, ... identical on all ranks
  double* out;                             // [4]
};

// Executed by ONE full warp.  local[4]: this rank's block sums (readable by every lane).  vals: 4 * MAX_WORLD doubles of
// shared memory private to this warp.  Returns nothing; out[0..3] = sum over ranks in rank order (bit-identical on every
// rank) or NaN after a timeout.  A rank that ever timed out sets the sticky error word of ITS buffer and from then on
// contributes NaN, so the failure reaches every peer with the next exchange instead of staying local.
__device__ __forceinline__ void allreduce_warp(const Args& a, const double* local, double* vals, long long timeout_cycles) {
  const int p = threadIdx.x & 31;
  const int parity = (int)(a.step & 1ull);
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  unsigned long long* my_buf = reinterpret_cast<unsigned long long*>(a.peer_bufs[a.rank]);
  unsigned long long* err_word = my_buf + (size_t)2 * a.world * SLOT_DOUBLES;
  const bool poisoned = *reinterpret_cast<volatile unsigned long long*>(err_word) != 0ull;
  bool failed = false;
  if (p < a.world) {
    // 1. my values -> slot [parity][rank] of peer p (plain stores, then the flag with release semantics at system scope)
    double* dst = reinterpret_cast<double*>(a.peer_bufs[p]) + ((size_t)parity * a.world + a.rank) * SLOT_DOUBLES;
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = poisoned ? qnan : local[k];
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(dst + 4), a.step);
    // 2. wait for rank p's values in MY buffer
    const double* src = reinterpret_cast<const double*>(my_buf) + ((size_t)parity * a.world + p) * SLOT_DOUBLES;
    const long long t0 = clock64();
    while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(src + 4)) != a.step) {
      if (clock64() - t0 > timeout_cycles) { failed = true; break; }   // never hang the GPU on a lost peer
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[p * 4 + k] = src[k];
  }
  const bool any_failed = __any_sync(0xffffffffu, failed);
  if (any_failed && p == 0) *reinterpret_cast<volatile unsigned long long*>(err_word) = a.step;
  __syncwarp();
  if (p < 4) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += vals[r * 4 + p];       // fixed order: identical bits on every rank
    a.out[p] = (any_failed || poisoned) ? qnan : s;
  }
}

}  // namespace p2p
}  // namespace wf
