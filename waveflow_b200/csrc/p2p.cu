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

// ---------------------------------------------------------------------------------------------- flat float32 vectors
// One-shot all-reduce (sum) of a float32 vector of a few hundred KB (the flat parameter gradient of the VQMC training step:
// 48 280 floats at D = 4) over NVLink peer memory, same idea as above with one CTA per chunk of VEC_CHUNK floats:
//   1. CTA b stores chunk b of the local vector into slot [parity][rank] of EVERY peer's buffer (128-bit stores straight into
//      peer memory), fences, and raises flag [parity][rank][b] = step on every peer;
//   2. it waits for the flags [parity][r][b] of all ranks r in its OWN buffer and adds the world copies of chunk b in rank
//      order -- identical bits on every rank, no dependence on arrival order;
//   3. the last CTA to finish advances the device-side step counter (CUDA-graph replays have no host-side arguments).
// A CTA publishes before it waits, and every CTA waits only for REMOTE CTAs of the same chunk, so there is no circular wait
// as long as each rank's grid eventually runs (<= 32 CTAs: co-resident on any GPU).
// Four doubles (the loss sums of the step) ride along with chunk 0.
// Buffer of one rank: data [2][world][n_pad] float | flags [2][world][n_chunks] uint64 | sums [2][world][4] double |
// header {sticky_error, pad...}.
constexpr int VEC_CHUNK = 2048;
constexpr int VEC_THREADS = 256;

__host__ __device__ inline int64_t vec_pad(int64_t n) { return (n + VEC_CHUNK - 1) / VEC_CHUNK * VEC_CHUNK; }
__host__ __device__ inline int64_t vec_chunks(int64_t n) { return vec_pad(n) / VEC_CHUNK; }
__host__ __device__ inline int64_t vec_flag_offset_bytes(int world, int64_t n) { return (int64_t)2 * world * vec_pad(n) * 4; }
__host__ __device__ inline int64_t vec_sums_offset_bytes(int world, int64_t n) {
  return vec_flag_offset_bytes(world, n) + (int64_t)2 * world * vec_chunks(n) * 8;
}
__host__ __device__ inline int64_t vec_buffer_bytes(int world, int64_t n) {
  return vec_sums_offset_bytes(world, n) + (int64_t)2 * world * 4 * 8 + 64;      // + {sum E, sum E^2, n, sum psi^2} slots + header
}

__global__ void __launch_bounds__(VEC_THREADS) p2p_allreduce_vec_kernel(const unsigned long long* __restrict__ peer_bufs, int rank, int world,
                                                                        unsigned long long step_arg, unsigned long long* step_dev,
                                                                        float* __restrict__ data, int64_t n, const double* __restrict__ sums,
                                                                        double* __restrict__ sums_out, unsigned int* done_counter,
                                                                        long long timeout_cycles) {
  __shared__ int failed_s;
  const unsigned long long step = step_dev ? *reinterpret_cast<volatile unsigned long long*>(step_dev) : step_arg;
  const int parity = (int)(step & 1ull);
  const int64_t n_pad = vec_pad(n), n_chunks = vec_chunks(n);
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) failed_s = 0;
  __syncthreads();
  char* my_buf = reinterpret_cast<char*>(peer_bufs[rank]);
  unsigned long long* err_word = reinterpret_cast<unsigned long long*>(my_buf + vec_buffer_bytes(world, n) - 64);
  const bool poisoned = *reinterpret_cast<volatile unsigned long long*>(err_word) != 0ull;
  const float qnan = __int_as_float(0x7fc00000);
  // 1. publish chunk b to every peer
  const int64_t e0 = b * VEC_CHUNK;
  for (int p = 0; p < world; ++p) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<char*>(peer_bufs[p]) + (((int64_t)parity * world + rank) * n_pad + e0) * 4);
    for (int i = tid; i < VEC_CHUNK / 4; i += VEC_THREADS) {
      const int64_t e = e0 + 4 * i;
      float4 v;
      v.x = e + 0 < n ? data[e + 0] : 0.f; v.y = e + 1 < n ? data[e + 1] : 0.f;
      v.z = e + 2 < n ? data[e + 2] : 0.f; v.w = e + 3 < n ? data[e + 3] : 0.f;
      if (poisoned) v = make_float4(qnan, qnan, qnan, qnan);
      dst[i] = v;
    }
    if (b == 0 && sums && tid < 4) {
      double* sd = reinterpret_cast<double*>(reinterpret_cast<char*>(peer_bufs[p]) + vec_sums_offset_bytes(world, n)) +
                   ((int64_t)parity * world + rank) * 4;
      sd[tid] = poisoned ? __longlong_as_double(0x7ff8000000000000ll) : sums[tid];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peer_bufs[tid]) + vec_flag_offset_bytes(world, n)) +
                               ((int64_t)parity * world + rank) * n_chunks + b;
    p2p::st_release_sys(flag, step);
    // 2. wait for rank `tid`'s chunk b in MY buffer
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(my_buf + vec_flag_offset_bytes(world, n)) +
                                     ((int64_t)parity * world + tid) * n_chunks + b;
    const long long t0 = clock64();
    while (p2p::ld_acquire_sys(mine) != step) {
      if (clock64() - t0 > timeout_cycles) { atomicExch(&failed_s, 1); break; }
    }
  }
  __syncthreads();
  const bool failed = failed_s != 0;
  if (failed && tid == 0) *reinterpret_cast<volatile unsigned long long*>(err_word) = step;
  for (int i = tid; i < VEC_CHUNK; i += VEC_THREADS) {
    const int64_t e = e0 + i;
    if (e < n) {
      float s = 0.f;
      for (int r = 0; r < world; ++r)
        s += *reinterpret_cast<const volatile float*>(my_buf + (((int64_t)parity * world + r) * n_pad + e) * 4);
      data[e] = (failed || poisoned) ? qnan : s;
    }
  }
  if (b == 0 && sums && sums_out && tid < 4) {
    double s = 0.0;
    for (int r = 0; r < world; ++r)
      s += *reinterpret_cast<const volatile double*>(my_buf + vec_sums_offset_bytes(world, n) + (((int64_t)parity * world + r) * 4 + tid) * 8);
    sums_out[tid] = (failed || poisoned) ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
  // 3. the last CTA advances the step counter of graph replays
  if (step_dev && done_counter) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(done_counter, 1u) == gridDim.x - 1) { *done_counter = 0u; *step_dev = step + 1ull; }
    }
  }
}

}  // namespace

extern "C" int64_t wf_p2p_allreduce_vec_buffer_bytes(int world, int64_t n) {
  return world >= 1 && world <= p2p::MAX_WORLD && n >= 0 ? vec_buffer_bytes(world, n) : -1;
}

extern "C" int wf_p2p_allreduce_vec(const uint64_t* peer_bufs_dev, int rank, int world, uint64_t step, uint64_t* step_dev, float* data,
                                    int64_t n, const double* sums, double* sums_out, uint32_t* done_counter, void* stream) {
  if (!peer_bufs_dev || !data || n < 0 || world < 1 || world > p2p::MAX_WORLD || rank < 0 || rank >= world) return WF_ERR_INVALID_ARG;
  if (!step_dev && step == 0) return WF_ERR_INVALID_ARG;
  if (step_dev && !done_counter) return WF_ERR_INVALID_ARG;
  if (n == 0) return WF_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(data) & 3) return WF_ERR_INVALID_ARG;
  if ((sums != nullptr) != (sums_out != nullptr)) return WF_ERR_INVALID_ARG;
  p2p_allreduce_vec_kernel<<<(int)vec_chunks(n), VEC_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(peer_bufs_dev), rank, world, (unsigned long long)step,
      reinterpret_cast<unsigned long long*>(step_dev), data, n, sums, sums_out, done_counter, p2p::TIMEOUT_CYCLES);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

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
