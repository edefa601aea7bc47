// One-shot all-reduce of the estimator block sums over NVLink peer memory.
//
// The VQMC step ends with an exchange of four doubles {sum E, sum E^2, n, sum psi^2} between the ranks.  At
// 8 GPUs the local-energy kernel takes 0.24 ms and an NCCL all-reduce of 32 bytes ~0.025 ms: pure latency.  Here every rank
// stores its four values plus a sequence flag straight into a slot of every peer's buffer (P2P stores over NVLink /
// NVSwitch into symmetric memory), then waits for the flags of all slots of its OWN buffer and adds the slots in rank
// order -- one small kernel, no ring, and bit-identical results on all ranks.
//
// Buffer of one rank (symmetric: same layout everywhere): slot[parity][src_rank] = { double v[4]; uint64 flag; pad[3] }.
// The parity alternates with the step: a rank can only be a whole step ahead of a peer (it needs the peer's flag to finish),
// so two generations of slots are enough.
#include "common.cuh"

namespace {

constexpr int SLOT_DOUBLES = 8;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(32) p2p_allreduce_kernel(const unsigned long long* __restrict__ peer_bufs, int rank, int world,
                                                           unsigned long long step, const double* __restrict__ local,
                                                           double* __restrict__ out, long long timeout_cycles) {
  __shared__ double vals[WF_MAX_D * 2][4];
  __shared__ int failed;
  const int p = threadIdx.x;
  if (p == 0) failed = 0;
  __syncwarp();
  const int parity = (int)(step & 1ull);
  if (p < world) {
    // 1. my values -> slot [parity][rank] of peer p (plain stores, then the flag with release semantics at system scope)
    double* dst = reinterpret_cast<double*>(peer_bufs[p]) + ((size_t)parity * world + rank) * SLOT_DOUBLES;
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = local[k];
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(dst + 4), step);
    // 2. wait for rank p's values in MY buffer
    const double* src = reinterpret_cast<const double*>(peer_bufs[rank]) + ((size_t)parity * world + p) * SLOT_DOUBLES;
    const long long t0 = clock64();
    while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(src + 4)) != step) {
      if (clock64() - t0 > timeout_cycles) { atomicExch(&failed, 1); break; }   // never hang the GPU on a lost peer
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[p][k] = src[k];
  }
  __syncwarp();
  if (p < 4) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += vals[r][p];       // fixed order: identical bits on every rank
    out[p] = failed ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
}

}  // namespace

extern "C" int64_t wf_p2p_allreduce_buffer_bytes(int world) {
  return world >= 1 && world <= 2 * WF_MAX_D ? (int64_t)2 * world * SLOT_DOUBLES * (int64_t)sizeof(double) : -1;
}

extern "C" int wf_p2p_allreduce_sums(const uint64_t* peer_bufs_dev, int rank, int world, uint64_t step, const double* local,
                                     double* out, void* stream) {
  if (!peer_bufs_dev || !local || !out || world < 1 || world > 2 * WF_MAX_D || rank < 0 || rank >= world || step == 0)
    return WF_ERR_INVALID_ARG;
  // ~2 s at 2 GHz: far beyond any legitimate skew between ranks, short enough not to look like a hung device
  p2p_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(peer_bufs_dev), rank, world,
                                                          (unsigned long long)step, local, out, 4000000000ll);
  WF_LAUNCH_CHECK();
  return WF_OK;
}
