// One-shot all-reduce of the estimator block sums over NVLink peer memory.
//
// The VQMC step ends with an exchange of four doubles {sum E, sum E^2, n, sum psi^2} between the ranks.  At
// 8 GPUs the local-energy kernel takes 0.24 ms and an NCCL all-reduce of 32 bytes ~0.025 ms: pure latency.  Here every rank
// stores its four values plus a sequence flag straight into a slot of every peer's buffer (P2P stores over NVLink /
// NVSwitch into symmetric memory), then waits for the flags of all slots of its OWN buffer and adds the slots in rank
// order -- one small kernel (or the tail of the local-energy kernel itself, see wf_local_energy_p2p), no ring, and
// bit-identical results on all ranks.
//
// Buffer of one rank (symmetric: same layout everywhere): slot[parity][src_rank] = { double v[4]; uint64 flag; pad[3] },
// followed by a header { uint64 sticky_error; pad[7] }.  The parity alternates with the step: a rank can only be a whole
// step ahead of a peer (it needs the peer's flag to finish), so two generations of slots are enough.
#include "p2p_device.cuh"

using namespace wf;

namespace {

__global__ void __launch_bounds__(32) p2p_allreduce_kernel(p2p::Args a, const double* __restrict__ local, long long timeout_cycles) {
  __shared__ double vals[p2p::MAX_WORLD * 4];
  p2p::allreduce_warp(a, local, vals, timeout_cycles);
}

// All `world` ranks emulated by the warps of ONE CTA (co-resident by construction, so the mutual flag waits are safe on a
// single GPU): warp r runs rank r's side of the protocol on buffer r.  skip_rank >= 0: that rank never shows up.
__global__ void __launch_bounds__(32 * p2p::MAX_WORLD) p2p_emulated_kernel(const unsigned long long* __restrict__ peer_bufs, int world,
                                                                            unsigned long long step, const double* __restrict__ locals,
                                                                            double* __restrict__ outs, int skip_rank,
                                                                            long long timeout_cycles) {
  __shared__ double vals[p2p::MAX_WORLD][p2p::MAX_WORLD * 4];
  const int r = threadIdx.x >> 5;
  if (r >= world || r == skip_rank) return;
  p2p::Args a{peer_bufs, r, world, step, outs + 4 * r};
  p2p::allreduce_warp(a, locals + 4 * r, vals[r], timeout_cycles);
}

}  // namespace

extern "C" int64_t wf_p2p_allreduce_buffer_bytes(int world) {
  return world >= 1 && world <= p2p::MAX_WORLD ? p2p::buffer_doubles(world) * (int64_t)sizeof(double) : -1;
}

extern "C" int wf_p2p_allreduce_sums(const uint64_t* peer_bufs_dev, int rank, int world, uint64_t step, const double* local,
                                     double* out, void* stream) {
  if (!peer_bufs_dev || !local || !out || world < 1 || world > p2p::MAX_WORLD || rank < 0 || rank >= world || step == 0)
    return WF_ERR_INVALID_ARG;
  p2p::Args a{reinterpret_cast<const unsigned long long*>(peer_bufs_dev), rank, world, (unsigned long long)step, out};
  p2p_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, local, p2p::TIMEOUT_CYCLES);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

extern "C" int wf_p2p_allreduce_emulated(const uint64_t* peer_bufs_dev, int world, uint64_t step, const double* locals,
                                         double* outs, int skip_rank, int64_t timeout_cycles, void* stream) {
  if (!peer_bufs_dev || !locals || !outs || world < 1 || world > p2p::MAX_WORLD || step == 0 || skip_rank >= world)
    return WF_ERR_INVALID_ARG;
  if (timeout_cycles <= 0) timeout_cycles = p2p::TIMEOUT_CYCLES;
  p2p_emulated_kernel<<<1, 32 * world, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(peer_bufs_dev), world,
                                                                  (unsigned long long)step, locals, outs, skip_rank,
                                                                  (long long)timeout_cycles);
  WF_LAUNCH_CHECK();
  return WF_OK;
}
