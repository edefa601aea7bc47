// Tensor-core variant of the fused live path (same arithmetic as live_kernel.cuh, conditioner layers 2 and 3 on tcgen05).
//
// The forward-Laplacian of the live path propagates, for every walker, G = D + 2 jet components (value, D gradient entries,
// Laplacian) through the conditioner MLPs.  A linear layer acts on every component alike, so with rows = (walker, component)
// the hidden and output layers are plain [rows, 64] x [64, 64 | 32 D] products -- 86 % of the dense FLOPs of the path.
// live_kernel.cuh evaluates them with FFMAs fed by broadcast shared-memory loads (ncu: FMA pipe 37 %, short-scoreboard
// stalls on the weight LDS.128, 12 warps per SM); here they run on the 5th-generation tensor cores:
//
//   * one TILE = 128 rows = 4 warps x (32 / G) walkers x G components; the lane <-> (walker, component) mapping inside a
//     warp is exactly Ctx<D, LAP>'s, so the jet algebra of live_device.cuh (warp shuffles between the components of a
//     walker) is reused unchanged, and thread r of a warpgroup owns TMEM lane r;
//   * the A operand (activations) lives in TENSOR MEMORY: every thread writes its row of tanh outputs as two TF32-exact
//     planes (hi, lo) with tcgen05.st, the MMAs read them with the .ts form ([d_tmem], [a_tmem], b_desc) -- no shared-memory
//     staging of the activations, no swizzling, no bank conflicts;
//   * the B operand (weights, hi / lo planes) is a pre-swizzled SWIZZLE_128B K-major image produced once per parameter set
//     (wf_live_pack_tc) and dropped into shared memory by one cp.async.bulk per layer and net;
//   * "3xTF32": D = A_hi B_hi + A_hi B_lo + A_lo B_hi accumulated in fp32 in TMEM -- float32-grade results (the scheme of
//     tc_gemm.cuh, measured there at 3e-6 relative for K = 512);
//   * accumulators are read back with tcgen05.ld by the thread that owns the row and the tanh / sigmoid-normalise-spline /
//     B-prior jet algebra runs as the epilogue, straight from tensor memory.
//
// CTA = 512 threads = 2 tiles x 2 warpgroups.  Both warpgroups of a tile hold the same 128 rows (warp w and warp w + 4 share
// a TMEM lane quarter) and split the per-row work: hidden features 0..31 / 32..63 in the tanh layers, output dimensions
// [0, D/2) / [D/2, D) in the spline glue; the per-dimension results are exchanged through the scratch columns.
// TMEM: 2 tiles x (A_hi 64 + A_lo 64 + accumulator 128) = 512 columns.  Shared memory: W2 planes 32 KB + W3 planes 16 D KB +
// 32 floats of scratch per thread (64 KB).
//
// Reference: the same lines as live_device.cuh / live_kernel.cuh.
#pragma once
#include "live_device.cuh"
#include "p2p_device.cuh"

namespace wf {
namespace ltc {

constexpr int THREADS = 512;
constexpr int TILE_THREADS = 256;          // two warpgroups
constexpr int TILES = 2;
constexpr int SCR = 32;                    // floats of thread-private scratch
constexpr int W2_PLANE_BYTES = WF_HIDDEN * WF_HIDDEN * 4;          // 16 KB
__host__ __device__ constexpr int w3_plane_bytes(int D) { return D * WF_MAX_P * WF_HIDDEN * 4; }   // 8 KB per dimension
__host__ __device__ constexpr int small_floats(int D) { return D * WF_HIDDEN + WF_HIDDEN + WF_HIDDEN + D * WF_MAX_P; }
// packed image of one conditioner (floats): W2 hi | W2 lo | W3 hi | W3 lo | W1 [D][64] | b1 [64] | b2 [64] | b3 [D][32]
__host__ __device__ constexpr int net_floats_tc(int D) {
  return 2 * (W2_PLANE_BYTES / 4) + 2 * (w3_plane_bytes(D) / 4) + small_floats(D);
}

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128 (A: lane = row, one column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// The same, executed by a WHOLE (converged) warp with warp-uniform operands: one lane is elected inside.  Issued from a
// single-lane branch instead, the compiler cannot keep the descriptors in uniform registers and wraps every MMA in an
// ELECT / R2UR.BROADCAST x4 / BRA.U.ANY loop (~100 cycles per MMA: the 24-MMA groups of this kernel were issue bound).
__device__ __forceinline__ void umma_tf32_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// issue only: the registers are valid after the next tmem_wait_ld() (lets the next chunk's load fly under the current chunk's math)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// wait::ld that names the destination registers as read-write operands: uses of them cannot be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld4(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
  tmem_wait_ld();
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart); see tc_gemm.cuh
__device__ __forceinline__ uint64_t smem_desc_sw128(const void* p) {
  const uint64_t addr = (uint64_t)(smem_u32(p) >> 4) & 0x3FFFull;
  return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __host__ __forceinline__ float tf32_rn(float a) {       // round-to-nearest-even to TF32 (weight images, host tests)
  uint32_t u;
  memcpy(&u, &a, 4);
  u += 0x0FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&a, &u, 4);
  return a;
}
// activations: one instruction (cvt.rna.tf32.f32, round to nearest, ties away)
__device__ __forceinline__ float tf32_rna(float a) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a));
  return __uint_as_float(u);
}

// mbarrier wait with a watchdog: a protocol error must end the kernel (trap -> launch failure), never hang the device
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}

// 32 output columns of one conditioner layer on the tensor cores:
//   acc[128 x 32] = A (hi + lo planes in TMEM) x W[rows r0 .. r0 + 31]^T (hi / lo images in shared memory, N_ROWS rows per k-block).
// The layers are issued in 32-column groups, each committed to its own mbarrier, so that the warpgroup that consumes a group
// starts its epilogue as soon as THAT group is done instead of waiting for the whole layer.
template <int N_ROWS, int NCOLS = 32>
__device__ __forceinline__ void issue_group32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, const unsigned char* b_hi,
                                              const unsigned char* b_lo, int r0) {
  constexpr uint32_t idesc = instr_desc_tf32(NCOLS);
  uint64_t dh = smem_desc_sw128(b_hi + (size_t)r0 * 128), dl = smem_desc_sw128(b_lo + (size_t)r0 * 128);
  // rolled on purpose (code size: the kernel must stay inside the instruction cache); 8 k-steps x 3 MMAs
#pragma unroll 1
  for (int ks = 0; ks < WF_HIDDEN / 8; ++ks) {
    const uint32_t acol = (uint32_t)(ks * 8);
    umma_tf32_ts_warp(d_tmem, a_hi + acol, dh, idesc, ks != 0 ? 1u : 0u);
    umma_tf32_ts_warp(d_tmem, a_hi + acol, dl, idesc, 1u);
    umma_tf32_ts_warp(d_tmem, a_lo + acol, dh, idesc, 1u);
    // next k-step: 32 bytes further inside the 128-byte swizzle span, or the start of the next 32-float k-block
    const uint64_t adv = ((ks & 3) == 3) ? (uint64_t)((N_ROWS * 128 - 96) >> 4) : (uint64_t)(32 >> 4);
    dh += adv; dl += adv;
  }
}

// sigmoid / tanh through ex2.approx.ftz + rcp.approx: __expf carries a denormal fix-up (a compare and two predicated multiplies
// per call) that the saturating forms below do not need (ex2 -> 0 or inf gives exactly 1 / 0 / +-1)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tc_sigmoid(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
// no clamp needed: ex2 saturates to 0 / inf, rcp(inf) = 0, so the form returns exactly -1 / +1 for large |x|
__device__ __forceinline__ float tc_tanh(float x) { return fmaf(-2.f, rcp_ftz(ex2_ftz(2.8853900817779268f * x) + 1.f), 1.f); }
// sum of t over the D gradient lanes of a walker, valid ON THE LAPLACIAN LANE ONLY (the only consumer in the tanh layers).
// D = 4: a two-step tree over shfl_up plus one hop (3 shuffles, 2 adds) instead of D shuffles and D adds.
template <int D, bool LAP>
__device__ __forceinline__ float lap_gsum(const Ctx<D, LAP>& cx, float t) {
  if constexpr (LAP && D == 4) {
    const float u = t + __shfl_up_sync(FULL, t, 2);      // lane g4: t4 + t2, lane g3: t3 + t1
    const float w = u + __shfl_up_sync(FULL, u, 1);      // lane g4: t4 + t2 + t3 + t1
    return __shfl_up_sync(FULL, w, 1);                   // Laplacian lane = g4 + 1
  } else return cx.gsum(t);
}
// tanh on a 1-register bundle (same algebra as tanh_bundle of live_device.cuh)
template <int D, bool LAP>
__device__ __forceinline__ float tc_tanh_bundle(const Ctx<D, LAP>& cx, float a) {
  if constexpr (LAP) {
    const float th = tc_tanh(cx.bv(a));
    const float f1 = 1.f - th * th;
    const float gg = lap_gsum<D, LAP>(cx, a * a);
    float r = f1 * a;
    if (cx.is_l) r = fmaf(-2.f * th * f1, gg, r);
    return cx.is_v ? th : r;
  } else return tc_tanh(a);
}

struct ScratchT {
  float* col;
  __device__ __forceinline__ float& operator[](int slot) const { return col[slot * THREADS]; }
};

// 8 activations -> (hi, lo) planes -> this thread's row of the A operand (columns col .. col + 7).
// hi = TF32(h) (round to nearest), lo = h - hi EXACTLY (|lo| <= 2^-11 |h|); the tensor core reads the top 19 bits of lo,
// i.e. truncates it: the pair carries h to 2^-21 relative, 2 instructions per value.
__device__ __forceinline__ void store_planes8(uint32_t a_hi, uint32_t a_lo, uint32_t col, const float (&h)[8]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float f = tf32_rna(h[t]);
    hi[t] = __float_as_uint(f);
    lo[t] = __float_as_uint(h[t] - f);
  }
  tmem_st8(a_hi + col, hi);
  tmem_st8(a_lo + col, lo);
}

// ---------------------------------------------------------------------------------------------------------------------
// sigmoid_spline of live_device.cuh with the conditioner outputs in REGISTERS (read from tensor memory): pass A is unrolled
// over the 32 coefficient slots; only s_q's own component is parked in the scratch column for the window pass, whose
// pending second-derivative term is rebuilt from it:  p = s'' o'^2 = (1 - 2 s) m^2 / (s (1 - s))  with  m = s' o'.
// The running sums carry only this lane's component (m) and the pending term (p); the VALUE of a sum, which every lane
// needs for the products that follow, is the value lane's m and is fetched by one shuffle per sum at the end (same bits as
// accumulating it redundantly on every lane, 3 FMAs per coefficient cheaper).
// rec_t: the compact node records transposed to [node][8 window slots][4 derivative orders], one 128-bit load per slot.
struct MP { float m, p; };

template <int D, bool LAP, int NOUT, bool PREFIX_ONE>
__device__ __forceinline__ void sigmoid_spline_regs(const Ctx<D, LAP>& cx, uint32_t tacc, const float* __restrict__ bias,
                                                    const ScratchT& S, int P, const float* __restrict__ wq,
                                                    const float* __restrict__ cwq, float wsum, float reg,
                                                    const float4* __restrict__ rec_t, const int32_t* __restrict__ lo,
                                                    const float* __restrict__ dense, int T, float xd, float xv, J& y, J& dy) {
  constexpr int NK = LAP ? NOUT + 2 : NOUT;
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(xv, T);
  const int lo_l = __ldg(lo + n.l), lo_r = __ldg(lo + n.r);
  const int sh = lo_r - lo_l;
  const bool local = (sh == 0) || (sh == 1);
  const int lo_w = local ? lo_l : 0;
  auto axpy = [&](float s, const J& a, MP& acc) {
    acc.m = fmaf(s, a.m, acc.m);
    if constexpr (LAP) acc.p = fmaf(s, a.p, acc.p);
  };
  auto full = [&](const MP& a) { return J{a.m, LAP ? a.p : 0.f, cx.bv(a.m)}; };

  // the prefix sum over the bases below the window (identically 1 there) is the running weighted sum SW at q == lo_w, and its
  // weight the host-computed prefix sum cwq[lo_w] of wq: a snapshot instead of a second predicated accumulation per coefficient
  MP Ssum = {0.f, 0.f}, SW = {0.f, 0.f}, PRE = {0.f, 0.f};
  const float Wpre = PREFIX_ONE ? cwq[lo_w] : 0.f;
  // conditioner outputs of this dimension: 8 accumulator columns at a time straight from tensor memory (+ bias on the value
  // lane); the chunk loop stays rolled to keep the kernel inside the instruction cache
  const float vmask = cx.is_v ? 1.f : 0.f;            // the bias enters the value component only
  uint32_t nxt[8];
  tmem_ld8_issue(tacc, nxt);
  tmem_wait_ld8(nxt);
#pragma unroll 1
  for (int c = 0; c < WF_MAX_P / 8; ++c) {
    float o8[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) o8[t] = __uint_as_float(nxt[t]);
    if (c + 1 < WF_MAX_P / 8) tmem_ld8_issue(tacc + (uint32_t)((c + 1) * 8), nxt);      // lands while this chunk is processed
    // straight-line code for the 8 coefficients of the chunk (no per-coefficient branch: the eight sigmoid chains overlap);
    // slots q >= P hold exact zeros (zero-padded weights) and are masked to sigmoid(-inf) = 0 with all derivatives 0
    const float4 bA = *reinterpret_cast<const float4*>(bias + c * 8), bB = *reinterpret_cast<const float4*>(bias + c * 8 + 4);
    const float b8[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int q = c * 8 + t;
      const float oq = fmaf(vmask, b8[t], o8[t]);
      const float ovr = cx.bv(oq);
      const float ov = q < P ? ovr : -INFINITY;
      const float s = tc_sigmoid(ov);
      const float d1 = s * (1.f - s);
      // cx.unary with a zero pending term, written out (the compiler cannot fold f1 * 0)
      J sq;
      if constexpr (LAP) {
        const float dm = d1 * oq;
        sq.m = cx.is_v ? s : dm;
        sq.p = cx.is_g ? (1.f - 2.f * s) * dm * oq : 0.f;
      } else { sq.m = s; sq.p = 0.f; }
      sq.v = s;
      const float w = wq[q];
      if (PREFIX_ONE) {
        const bool snap = q == lo_w;                   // SW holds the bases 0 .. q - 1 here
        PRE.m = snap ? SW.m : PRE.m;
        if constexpr (LAP) PRE.p = snap ? SW.p : PRE.p;
      }
      Ssum.m += sq.m;
      if constexpr (LAP) Ssum.p += sq.p;
      axpy(w, sq, SW);
      S[q] = sq.m;
    }
    tmem_wait_ld8(nxt);
  }
  if (PREFIX_ONE) {
    if (lo_w >= P) PRE = SW;                           // (window entirely beyond the last basis: every basis is in the prefix)
  }
  auto reload = [&](int qc) {
    J sq;
    sq.m = S[qc];
    sq.v = cx.bv(sq.m);
    if constexpr (LAP) {
      const float d1 = fmaxf(sq.v * (1.f - sq.v), 1e-30f);
      sq.p = cx.is_g ? (1.f - 2.f * sq.v) * sq.m * sq.m * rcp_ftz(d1) : 0.f;
    } else sq.p = 0.f;
    return sq;
  };
  MP Sk[NK];
  float Wk[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) { Sk[k] = MP{0.f, 0.f}; Wk[k] = 0.f; }
  if (PREFIX_ONE) { Sk[0] = PRE; Wk[0] = Wpre; }
  if (local) {
#pragma unroll 4
    for (int t = 0; t < WF_WIN; ++t) {
      const int q = lo_l + t;
      const int qc = q < P ? q : P - 1;
      const float w = q < P ? wq[qc] : 0.f;
      const J sq = reload(qc);
      const int tr = t - sh;                     // slot of this basis in the right node's record
      const float4 L4 = __ldg(rec_t + (size_t)n.l * WF_WIN + t);
      const float4 R4 = __ldg(rec_t + (size_t)n.r * WF_WIN + (tr < 0 ? 0 : tr));
      const float yl[4] = {L4.x, L4.y, L4.z, L4.w};
      const float yrr[4] = {R4.x, R4.y, R4.z, R4.w};
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int nd = k < 3 ? k : 3;
        const float yr = tr < 0 ? ((PREFIX_ONE && nd == 0) ? 1.f : 0.f) : yrr[nd];
        const float fw = lerp_tab(yl[nd], yr, np_, n.dx) * w;
        axpy(fw, sq, Sk[k]);
        Wk[k] += fw;
      }
    }
  } else {
    // reference-exact dense evaluation (arguments outside [0, 1]: JAX gather wrap / clamp semantics)
    if (PREFIX_ONE) { Sk[0] = MP{0.f, 0.f}; Wk[0] = 0.f; }
    for (int q = 0; q < P; ++q) {
      const J sq = reload(q);
      const float w = wq[q];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int nd = k < 3 ? k : 3;
        const float yl = __ldg(dense + ((size_t)n.l * 4 + nd) * WF_MAX_P + q);
        const float yr = __ldg(dense + ((size_t)n.r * 4 + nd) * WF_MAX_P + q);
        const float fw = lerp_tab(yl, yr, np_, n.dx) * w;
        axpy(fw, sq, Sk[k]);
        Wk[k] += fw;
      }
    }
  }
  // r = 1/S;  Z = SW * r + reg * Wsum;  iz = 1/Z;  N_k = Sk * r + reg * Wk;  A_k = N_k * iz
  const J r = cx.recip(full(Ssum));
  const J iz = cx.recip(cx.addc(cx.mul(full(SW), r), reg * wsum));
  J A[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) A[k] = cx.mul(cx.addc(cx.mul(full(Sk[k]), r), reg * Wk[k]), iz);
  if constexpr (LAP) {
    y = spline_assemble<D, LAP>(cx, A[0], A[1], A[2], xd);
    if (NOUT == 2) dy = spline_assemble<D, LAP>(cx, A[1], A[2], A[NK - 1], xd);
  } else {
    y = A[0];
    if (NOUT == 2) dy = A[NOUT - 1];
  }
}

// B prior factor with the third layer pre-multiplied by mask @ ob_to_b (bit 2 of bc_P; see bprior_factor): accumulator column q
// of this dimension holds c'_q, column 31 the sum of the raw conditioner outputs.  Four columns at a time from tensor memory.
template <int D, bool LAP>
__device__ __forceinline__ J bprior_factor_regs(const Ctx<D, LAP>& cx, uint32_t tacc, const float* __restrict__ bias, int P,
                                                const float* __restrict__ tab, int T, float xd_in, float xv) {
  constexpr int NK = LAP ? 3 : 1;
  const float xc = fminf(fmaxf(xv, 0.f), 1.f);
  const float xd = ((xv > 0.f) && (xv < 1.f)) ? xd_in : 0.f;
  const float np_ = (float)(T - 1);
  const NodeIdx n = node_index(xc, T);
  const float osum = tmem_ld1(tacc + (uint32_t)(WF_MAX_P - 1));
  const float sgn = cx.bv(cx.is_v ? osum + bias[WF_MAX_P - 1] : osum) < 0.f ? -1.f : 1.f;
  J A[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) A[k] = J{0.f, 0.f, 0.f};
  J Q = {0.f, 0.f, 0.f};
  uint32_t nxt[4];
  tmem_ld4_issue(tacc, nxt);
  tmem_wait_ld4(nxt);
#pragma unroll 1
  for (int j0 = 0; j0 < P; j0 += 4) {
    float o4[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) o4[t] = __uint_as_float(nxt[t]);
    if (j0 + 4 < P) tmem_ld4_issue(tacc + (uint32_t)(j0 + 4), nxt);
    float f[NK][4];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)n.l * 4 + k) * WF_MAX_P + j0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)n.r * 4 + k) * WF_MAX_P + j0));
      f[k][0] = lerp_tab(a.x, b.x, np_, n.dx); f[k][1] = lerp_tab(a.y, b.y, np_, n.dx);
      f[k][2] = lerp_tab(a.z, b.z, np_, n.dx); f[k][3] = lerp_tab(a.w, b.w, np_, n.dx);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float oq = cx.is_v ? o4[t] + bias[j0 + t] : o4[t];
      const float c = (j0 + t < P) ? oq : 0.f;
      const float cv = cx.bv(c);
#pragma unroll
      for (int k = 0; k < NK; ++k) { A[k].m = fmaf(f[k][t], c, A[k].m); A[k].v = fmaf(f[k][t], cv, A[k].v); }
      Q.v = fmaf(cv, cv, Q.v);
      if constexpr (LAP) {
        Q.m = cx.is_v ? Q.v : fmaf(2.f * cv, c, Q.m);
        Q.p = cx.is_g ? fmaf(2.f * c, c, Q.p) : 0.f;
      } else Q.m = Q.v;
    }
    tmem_wait_ld4(nxt);
  }
  J num;
  if constexpr (LAP) num = spline_assemble<D, LAP>(cx, A[0], A[1], A[2], xd); else num = A[0];
  const float isq = 1.f / sqrtf(Q.v);
  const J iq = cx.unary(Q, isq, -0.5f * isq / Q.v, 0.75f * isq / (Q.v * Q.v));
  return cx.scale(cx.mul(num, iq), sgn);
}

// 1 / sqrt(1 + d^2) with IEEE square root and division (utils/physics.py:66-71).  Deliberately NOT inlined: the correctly
// rounded sequences are ~50 instructions each and the potential needs D (n_protons + (D - 1) / 2) of them, once per walker.
static __device__ __noinline__ float inv_sqrt_1p_sq(float d) { return 1.f / sqrtf(1.f + d * d); }

template <int D>
__device__ __forceinline__ float soft_coulomb_tc(const float (&xs)[D], const float* protons, int n_protons) {
  float pe = 0.f;
  for (int p = 0; p < n_protons; ++p)
#pragma unroll 1
    for (int e = 0; e < D; ++e) pe += inv_sqrt_1p_sq(protons[p] - xs[e]);
  float ee = 0.f;
#pragma unroll
  for (int i = 1; i < D; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) ee += inv_sqrt_1p_sq(xs[i] - xs[j]);
  return ee - pe;
}

// Shared memory (1024-byte aligned): W2 hi | W2 lo | W3 hi | W3 lo | 2 x (W1, b1, b2, b3) | scratch [32][512] | reduction
// [4][16] doubles | barriers ...
struct Smem {
  static __host__ __device__ constexpr size_t w2_off() { return 0; }
  static __host__ __device__ constexpr size_t w3_off() { return 2 * (size_t)W2_PLANE_BYTES; }
  static __host__ __device__ constexpr size_t small_off(int D) { return w3_off() + 2 * (size_t)w3_plane_bytes(D); }
  static __host__ __device__ constexpr size_t scr_off(int D) { return small_off(D) + 2 * (size_t)small_floats(D) * 4; }
  static __host__ __device__ constexpr size_t red_off(int D) { return scr_off(D) + (size_t)SCR * THREADS * 4; }
  static __host__ __device__ constexpr size_t bar_off(int D) { return red_off(D) + 4 * (THREADS / 32) * sizeof(double); }
  static __host__ __device__ constexpr size_t total(int D) { return bar_off(D) + 128 + 1024; }   // + alignment slack
};

struct TcExtra {
  p2p::Args xchg;            // estimator exchange in the kernel tail (world > 1), else peer_bufs == nullptr
  unsigned int* done_counter;
};

template <int D, bool LAP>
__global__ void __launch_bounds__(THREADS, 1) live_tc_kernel(const __grid_constant__ LiveParams P, const __grid_constant__ TcExtra X) {
  using C = Ctx<D, LAP>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment (swizzled MMA operands) by an OFFSET on the shared array: an integer round trip of the pointer would
  // make every later access a generic LD.E / ST.E (long-scoreboard latency) instead of LDS / STS
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int N3 = D * WF_MAX_P;
  constexpr int NETF = net_floats_tc(D);
  constexpr uint32_t W2_BYTES = 2 * W2_PLANE_BYTES, W3_BYTES = 2 * w3_plane_bytes(D), SMALL_BYTES = small_floats(D) * 4;
  const wf_live_model& M = P.m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int quarter = warp & 3;                        // TMEM lane quarter of this warp

  unsigned char* w2_s = smem + Smem::w2_off();
  unsigned char* w3_s = smem + Smem::w3_off();
  float* small_s = reinterpret_cast<float*>(smem + Smem::small_off(D));      // [2][small_floats]: W1 | b1 | b2 | b3 of net g & 1
  float* scratch = reinterpret_cast<float*>(smem + Smem::scr_off(D));
  double* red_s = reinterpret_cast<double*>(smem + Smem::red_off(D));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bar_off(D));
  uint64_t* w2_full = bars;            // count 1 + tx: W2 planes + the small arrays of a net
  uint64_t* w3_full = bars + 1;
  uint64_t* l2_done = bars + 2;        // [TILES][2]: 32-column halves of layer 2, arrived by tcgen05.commit
  uint64_t* l3_done = bars + 6;        // [TILES][4]: per-dimension 32-column groups of layer 3
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
  int* w_cnt = reinterpret_cast<int*>(tmem_ptr + 1);   // [2]: teams done with W2 / W3 (cumulative over the nets)

  if (tid == 0) {
    mbar_init(w2_full, 1); mbar_init(w3_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&l2_done[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&l3_done[i], 1);
    w_cnt[0] = 0; w_cnt[1] = 0;
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  fence_before();
  __syncthreads();
  fence_after();
  // broadcast from lane 0: tells the compiler these are warp-uniform (the MMA issue below then runs on the uniform datapath)
  const uint32_t tmem_base = __shfl_sync(FULL, *tmem_ptr, 0);
  const int warp_u = __shfl_sync(FULL, warp, 0);
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

  C cx;
  cx.init(lane);
  const ScratchT S{scratch + tid};
  constexpr int WPT = 4 * C::WPW;                      // walkers per tile
  const int64_t n_tiles = (P.N + WPT - 1) / WPT;
  const int64_t slots = (int64_t)gridDim.x * TILES;
  // tile of (round, slot s) = round * slots + s * gridDim.x + blockIdx.x: consecutive tiles go to DIFFERENT CTAs, so the ragged
  // last round leaves ONE tile per CTA (which then gets all 16 warps, see below) rather than idle SMs
  const int64_t my_rounds = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + slots - 1) / slots : 0;
  const int n_nets = P.n_nets;
  const bool has_prior_net = M.prior_kind == WF_KIND_B || M.prior_kind == WF_KIND_M;
  const int64_t g_total = my_rounds * n_nets;
  double accE = 0.0, accE2 = 0.0, accN = 0.0, accP2 = 0.0;
  // phase parities of the per-tile MMA barriers (a slot that sits out a SOLO round does not advance)
  uint32_t par_s0 = 0, par_s1 = 0;

  auto issue_w2 = [&](int64_t g) {
    const float* src = P.weights + (size_t)(g % n_nets) * NETF;
    mbar_expect_tx(w2_full, W2_BYTES + SMALL_BYTES);
    bulk_g2s(w2_s, src, W2_BYTES, w2_full);
    bulk_g2s(small_s + (size_t)(g & 1) * small_floats(D), src + (W2_BYTES + W3_BYTES) / 4, SMALL_BYTES, w2_full);
  };
  auto issue_w3 = [&](int64_t g) {
    mbar_expect_tx(w3_full, W3_BYTES);
    bulk_g2s(w3_s, P.weights + (size_t)(g % n_nets) * NETF + W2_BYTES / 4, W3_BYTES, w3_full);
  };
  if (tid == 0 && g_total > 0) { issue_w2(0); issue_w3(0); }

  int64_t g = 0;
  int cum_teams = 0;                                    // arrivals expected on w_cnt[*] through the current net
  for (int64_t round = 0; round < my_rounds; ++round) {
    const int64_t t0 = round * slots + blockIdx.x, t1 = t0 + gridDim.x;
    // SOLO round: only slot 0 has a tile -> all four warpgroups work on it (the hidden features and the output dimensions
    // are split four ways instead of two): a lone tile is latency bound, this nearly halves its latency.  DUO: two teams of
    // two warpgroups, one tile each.
    const bool solo = t1 >= n_tiles;
    // the warps of the other team join slot 0's tile: they must not touch its tensor memory before slot 0's own team has
    // finished the previous (DUO) round there -- and slot 1's team must be done with its tile as well
    if (solo && round > 0) __syncthreads();
    const int n_wg = solo ? 4 : 2;
    const int slot = solo ? 0 : warp_u / (TILE_THREADS / 32);
    const int wg = solo ? warp_u >> 2 : (warp_u >> 2) & 1;   // warpgroup inside the team
    const int team_tid0 = solo ? 0 : slot * TILE_THREADS;
    const int team_threads = solo ? THREADS : TILE_THREADS;
    const bool leader = tid == team_tid0;
    const int teams = solo ? 1 : 2;
    const uint32_t par = slot == 0 ? par_s0 : par_s1;
    const uint32_t cbase = tmem_base + (uint32_t)(slot * 256);
    const uint32_t a_hi = cbase, a_lo = cbase + 64, dacc = cbase + 128;   // MMA operands (lane field 0)
    const uint32_t my_hi = a_hi + lane_addr, my_lo = a_lo + lane_addr, my_acc = dacc + lane_addr;
    const int fw = WF_HIDDEN / n_wg;                     // hidden features of the tanh layers per warpgroup
    const int fbase = wg * fw;
    // output dimensions of the spline glue handled by this warpgroup
    const int d_lo = solo ? (wg < D ? wg : D) : (wg == 0 ? 0 : D / 2);
    const int d_hi = solo ? (wg < D ? wg + 1 : D) : (wg == 0 ? D / 2 : D);
    auto owner_col = [&](int owner_wg) { return scratch + team_tid0 + owner_wg * 128 + (tid & 127); };
    auto owner_of = [&](int d) { return solo ? d : (d >= D / 2 ? 1 : 0); };

    const int64_t tile_idx = slot == 0 ? t0 : t1;
    const int64_t w_raw = tile_idx * WPT + (int64_t)quarter * C::WPW + cx.slot;
    const bool lane_live = (!LAP || lane < C::WPW * C::G) && w_raw < P.N;
    const int64_t w = w_raw < P.N ? w_raw : P.N - 1;

    float us[D];
    // per-DIMENSION partial results (log-det contributions summed over the nets, prior factor, log prior): the final
    // combination below always runs over the dimensions in index order, so every walker's result is bit-identical whichever
    // team layout (SOLO / DUO) -- i.e. whichever batch size or sharding -- evaluated it
    float ldbox, ldv[D], phiv[D], lpv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { ldv[d] = 0.f; phiv[d] = cx.is_v ? 1.f : 0.f; lpv[d] = 0.f; }
    {
      float xs[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xs[d] = __ldg(P.x + w * D + d);
      J ld = cx.constant(0.f);
      box_transform<D, LAP>(cx, M, xs, us, ld);
      ldbox = cx.fold(ld);                                // 1-register bundle (every warpgroup holds it; warpgroup 0 exports it)
    }
    if (M.n_layers == 0 && wg == 0 && lane_live && cx.is_v && P.u) {
#pragma unroll
      for (int d = 0; d < D; ++d) P.u[w * D + d] = us[d];
    }

#pragma unroll 1
    for (int net_idx = 0; net_idx < n_nets; ++net_idx, ++g) {
      const bool is_prior = has_prior_net && net_idx == n_nets - 1;
      cum_teams += teams;
      // W1 | b1 | b2 | b3 of this net were dropped into shared memory together with its W2 planes
      mbar_wait_guard(w2_full, (uint32_t)(g & 1));
      const float* W1 = small_s + (size_t)(g & 1) * small_floats(D);
      const float* b1 = W1 + D * WF_HIDDEN;
      const float* b2 = b1 + WF_HIDDEN;
      const float* b3 = b2 + WF_HIDDEN;
      const uint32_t mpar = (par + (uint32_t)net_idx) & 1u;

      // ------------------------------------------------ layer 1 (K = D, CUDA cores) -> A planes
#pragma unroll 1
      const float vmask = cx.is_v ? 1.f : 0.f;          // biases enter the value component only
      for (int c = 0; c < fw / 8; ++c) {
        const int j0 = fbase + c * 8;
        float h[8];
        {
          const float4 ba = *reinterpret_cast<const float4*>(b1 + j0), bb = *reinterpret_cast<const float4*>(b1 + j0 + 4);
          h[0] = vmask * ba.x; h[1] = vmask * ba.y; h[2] = vmask * ba.z; h[3] = vmask * ba.w;
          h[4] = vmask * bb.x; h[5] = vmask * bb.y; h[6] = vmask * bb.z; h[7] = vmask * bb.w;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {                   // the same 8 weights for every thread: two broadcast 128-bit loads per input
          const float4 wa = *reinterpret_cast<const float4*>(W1 + d * WF_HIDDEN + j0);
          const float4 wb = *reinterpret_cast<const float4*>(W1 + d * WF_HIDDEN + j0 + 4);
          h[0] = fmaf(us[d], wa.x, h[0]); h[1] = fmaf(us[d], wa.y, h[1]); h[2] = fmaf(us[d], wa.z, h[2]); h[3] = fmaf(us[d], wa.w, h[3]);
          h[4] = fmaf(us[d], wb.x, h[4]); h[5] = fmaf(us[d], wb.y, h[5]); h[6] = fmaf(us[d], wb.z, h[6]); h[7] = fmaf(us[d], wb.w, h[7]);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) h[t] = tc_tanh_bundle<D, LAP>(cx, h[t]);
        store_planes8(my_hi, my_lo, (uint32_t)(fbase + c * 8), h);
      }
      tmem_wait_st();
      fence_before();
      bar_sync(1 + slot, team_threads);
      // every 32-column group is issued by ONE lane of a different warp: the issue work (24 MMAs per group) is spread
      // instead of serialised in front of one warp's epilogue.  Layer 2: two groups (output units 0..31 / 32..63).
      // ONE 24-MMA group per layer (N = 64 here), issued by the first warp of the team (whole warp, warp-uniform branch, one lane
      // elected inside).  A tcgen05.mma of this size costs about the same whatever N <= 128 is, so splitting a layer into
      // 32- or 16-column groups multiplies the tensor time (measured: 4 x 32 columns 0.926 ms, 8 x 16 columns 0.956 ms per
      // 65 536 walkers) -- the whole layer in one group is what lets the epilogues start earliest.
      if (warp_u == (team_tid0 >> 5)) {
        fence_after();
        issue_group32<WF_HIDDEN, WF_HIDDEN>(dacc, a_hi, a_lo, w2_s, w2_s + W2_PLANE_BYTES, 0);
        umma_commit_warp(&l2_done[slot * 2]);
      }
      // the A planes are overwritten below: BOTH halves of layer 2 must have been read by the tensor core
      mbar_wait_guard(&l2_done[slot * 2], mpar);
      fence_after();
      if (leader) {                                     // the last team to get here refills W2 with the next net
        const int old = atomicAdd(&w_cnt[0], 1);
        if (old + 1 == cum_teams && g + 1 < g_total) issue_w2(g + 1);
      }

      // ------------------------------------------------ layer 2 epilogue: bias + tanh -> A planes
      {
        uint32_t nxt[8];
        tmem_ld8_issue(my_acc + (uint32_t)fbase, nxt);
        tmem_wait_ld8(nxt);
#pragma unroll 1
        for (int c = 0; c < fw / 8; ++c) {
          float v[8], h[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(nxt[t]);
          if (c + 1 < fw / 8) tmem_ld8_issue(my_acc + (uint32_t)(fbase + (c + 1) * 8), nxt);    // in flight under this chunk's tanh
          const float4 ba = *reinterpret_cast<const float4*>(b2 + fbase + c * 8), bb = *reinterpret_cast<const float4*>(b2 + fbase + c * 8 + 4);
          const float b8[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int t = 0; t < 8; ++t) h[t] = tc_tanh_bundle<D, LAP>(cx, fmaf(vmask, b8[t], v[t]));
          tmem_wait_ld8(nxt);
          store_planes8(my_hi, my_lo, (uint32_t)(fbase + c * 8), h);
        }
      }
      tmem_wait_st();
      fence_before();
      bar_sync(1 + slot, team_threads);
      {
        // layer 3: one 32-column group per output dimension, issued by warp (d - d_lo) of the warpgroup that owns dimension d
        if (warp_u == (team_tid0 >> 5)) {                 // layer 3: all D x 32 columns in one group
          fence_after();
          mbar_wait_guard(w3_full, (uint32_t)(g & 1));
          issue_group32<N3, N3>(dacc, a_hi, a_lo, w3_s, w3_s + w3_plane_bytes(D), 0);
          umma_commit_warp(&l3_done[slot * 4]);
        }
      }
      mbar_wait_guard(&l3_done[slot * 4], mpar);
      fence_after();
      if (leader) {                                     // W3 is free once the layer-3 groups of ALL teams are done: refill it
        const int old = atomicAdd(&w_cnt[1], 1);
        if (old + 1 == cum_teams && g + 1 < g_total) issue_w3(g + 1);
      }

      // ------------------------------------------------ layer 3 epilogue: spline glue of this warpgroup's dimensions
      float ys[D];
#pragma unroll
      for (int d = 0; d < D; ++d) ys[d] = 0.f;
#pragma unroll 1
      for (int d = d_lo; d < d_hi; ++d) {
        const uint32_t tacc = my_acc + (uint32_t)(d * WF_MAX_P);      // this row's 32 conditioner outputs of dimension d
        const float* b3d = b3 + d * WF_MAX_P;
        float xd = us[0];
#pragma unroll
        for (int dd = 1; dd < D; ++dd) xd = (d == dd) ? us[dd] : xd;
        const float xv = cx.bv(xd);
        if (!is_prior) {
          J y, dy;
          sigmoid_spline_regs<D, LAP, 2, true>(cx, tacc, b3d, S, M.P_I, P.wq_I, P.cwq_I, P.wsum_I, M.reg, reinterpret_cast<const float4*>(P.rec_I_t),
                                               P.lo_I, P.tab_I, M.T, xd, xv, y, dy);
          const float yf = cx.fold(y);
#pragma unroll
          for (int dd = 0; dd < D; ++dd) ys[dd] = (d == dd) ? yf : ys[dd];
          const J l = cx.log(cx.addc(dy, LOG_TOL));
          const float lf = LAP ? cx.fold(l) : l.v;
#pragma unroll
          for (int dd = 0; dd < D; ++dd) ldv[dd] = (d == dd) ? ldv[dd] + lf : ldv[dd];
        } else {
          const bool cons = M.coord_mean ? (d < D - 1) : (d >= 1);
          if (M.prior_kind == WF_KIND_B) {
            J phi = bprior_factor_regs<D, LAP>(cx, tacc, b3d, M.P_P, P.tab_P, M.T, xd, xv);
            float lpd = 0.f;
            if (!LAP) {
              float pr = phi.v * phi.v;
              if (cons) pr = pr / 2.f;
              lpd = logf(pr + LOG_TOL);
            }
            if (cons) phi = cx.scale(phi, 0.70710678118654752f);
            const float pf = cx.fold(phi);
#pragma unroll
            for (int dd = 0; dd < D; ++dd) { phiv[dd] = (d == dd) ? pf : phiv[dd]; lpv[dd] = (d == dd) ? lpd : lpv[dd]; }
          } else {
            const float xc = fminf(fmaxf(xv, 0.f), 1.f);
            const float xdc = (xv > 0.f && xv < 1.f) ? xd : 0.f;
            J y, dy;
            sigmoid_spline_regs<D, LAP, 1, false>(cx, tacc, b3d, S, M.P_P, P.wq_P, P.cwq_I, P.wsum_P, 0.f, reinterpret_cast<const float4*>(P.rec_P_t),
                                                  P.lo_P, P.tab_P, M.T, xdc, xc, y, dy);
            const float lpd = logf(y.v + LOG_TOL);
#pragma unroll
            for (int dd = 0; dd < D; ++dd) lpv[dd] = (d == dd) ? lpd : lpv[dd];
          }
        }
      }
      // ------------------------------------------------ exchange the per-dimension results between the warpgroups of the team
      if (!is_prior) {
#pragma unroll
        for (int d = 0; d < D; ++d)
          if (d >= d_lo && d < d_hi) S[d] = ys[d];
        fence_before();                                 // (the accumulator columns are free again after this barrier)
        bar_sync(1 + slot, team_threads);
#pragma unroll
        for (int d = 0; d < D; ++d) ys[d] = owner_col(owner_of(d))[d * THREADS];
#pragma unroll
        for (int d = 0; d < D; ++d) us[d] = ys[D - 1 - d];          // Reverse (bijections.py:336-345)
        if (net_idx == M.n_layers - 1 && wg == 0 && lane_live && cx.is_v && P.u) {      // (the value lane's bundle IS the value)
#pragma unroll
          for (int d = 0; d < D; ++d) P.u[w * D + d] = us[d];
        }
      }
    }
    // every thread tracks the phase of BOTH slots' barriers (it may serve either slot in a later round)
    par_s0 = (par_s0 + (uint32_t)n_nets) & 1u;
    if (!solo) par_s1 = (par_s1 + (uint32_t)n_nets) & 1u;

    // ------------------------------------------------ combine the warpgroups: psi = prod_d phi_d * exp(0.5 log_det)
    bar_sync(1 + slot, team_threads);                   // every warpgroup finished reading the previous exchange
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (d >= d_lo && d < d_hi) { S[d] = phiv[d]; S[D + d] = ldv[d]; S[2 * D + d] = lpv[d]; }
    fence_before();
    bar_sync(1 + slot, team_threads);
    float ld_tot = ldbox, lp_tot = 0.f;
    J psi_t = cx.constant(1.f);
#pragma unroll
    for (int d = 0; d < D; ++d) {                       // fixed order over the dimensions, identical in every warpgroup
      const float* col = owner_col(owner_of(d));
      const float pk = col[d * THREADS];
      ld_tot += col[(D + d) * THREADS];
      lp_tot += col[(2 * D + d) * THREADS];
      const J ph = J{pk, 0.f, cx.bv(pk)};
      if (d == 0) psi_t = ph;
      else {
        psi_t = cx.mul(psi_t, ph);
        if constexpr (LAP) { const float f = cx.fold(psi_t); psi_t = J{f, 0.f, psi_t.v}; }
      }
    }
    const J LD = J{ld_tot, 0.f, cx.bv(ld_tot)};
    if (M.prior_kind == WF_KIND_B) psi_t = cx.mul(psi_t, cx.exp(cx.scale(LD, 0.5f)));
    const float psi1 = cx.fold(psi_t);

    // ------------------------------------------------ outputs (warpgroup 0 of the team writes)
    const bool writer = wg == 0 && lane_live;
    if (writer && cx.is_v) {
      if (P.logdet) P.logdet[w] = LD.v;
      if (P.logpdf) P.logpdf[w] = lp_tot + LD.v;
      if (P.psi) P.psi[w] = psi_t.v;
    }
    if constexpr (LAP) {
      const float lapv = __shfl_sync(FULL, psi1, cx.gbase + D + 1);
      if (writer && cx.is_g && P.grad) P.grad[w * D + (cx.comp - 1)] = psi1;
      if (wg == 0) {
        float xs[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xs[d] = __ldg(P.x + w * D + d);
        if (writer && cx.is_v) {
          const float V = soft_coulomb_tc<D>(xs, P.protons, P.n_protons);
          const float hp = fmaf(-0.5f, lapv, V * psi_t.v);         // physics.py:84
          const float el = hp / (psi_t.v + 1e-8f);                // vqmc.py:200
          if (P.lap) P.lap[w] = lapv;
          if (P.hpsi) P.hpsi[w] = hp;
          if (P.eloc) P.eloc[w] = el;
          accE += (double)el; accE2 += (double)el * (double)el; accN += 1.0; accP2 += (double)psi_t.v * (double)psi_t.v;
        }
      }
    }
  }

  bool last_cta = false;
  if constexpr (LAP) {
    if (P.sums) {
      double v[4] = {accE, accE2, accN, accP2};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], off);
        if (lane == 0) red_s[k * (THREADS / 32) + warp] = v[k];
      }
      __syncthreads();
      if (tid < 4) {
        double s = 0.0;
        for (int i = 0; i < THREADS / 32; ++i) s += red_s[tid * (THREADS / 32) + i];
        atomicAdd(P.sums + tid, s);
      }
      // estimator exchange in the tail of the kernel: the last CTA to retire publishes the block sums to the peers
      if (X.xchg.peer_bufs != nullptr) {
        __shared__ int is_last;
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          const unsigned int done = atomicAdd(X.done_counter, 1u);
          is_last = (done == gridDim.x - 1) ? 1 : 0;
          if (is_last) *X.done_counter = 0u;                 // ready for the next launch
        }
        __syncthreads();
        last_cta = is_last != 0;
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  if constexpr (LAP) {
    if (last_cta && warp == 1) {
      __threadfence();
      __shared__ double xv_s[p2p::MAX_WORLD * 4];
      __shared__ double loc_s[4];
      if (lane < 4) loc_s[lane] = atomicAdd(P.sums + lane, 0.0);     // coherent read of the completed sums
      __syncwarp();
      p2p::allreduce_warp(X.xchg, loc_s, xv_s, p2p::TIMEOUT_CYCLES);
    }
  }
}

template <int D, bool LAP>
int launch_live_tc(LiveParams& P, const TcExtra& X, cudaStream_t s) {
  const size_t smem = Smem::total(D);
  if (smem > 227 * 1024) return WF_ERR_UNSUPPORTED;
  const int wpw = LAP ? 32 / (D + 2) : 32;
  const int64_t wpt = 4 * wpw;
  const int64_t n_tiles = (P.N + wpt - 1) / wpt;
  // up to one tile per SM: every CTA runs its tile with all four warpgroups (SOLO); beyond that two tiles share an SM (DUO)
  const int blocks = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  WF_CUDA(cudaFuncSetAttribute(live_tc_kernel<D, LAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  live_tc_kernel<D, LAP><<<blocks, THREADS, smem, s>>>(P, X);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

}  // namespace ltc

#define WF_DECL_LIVE_TC(D) \
  int launch_live_tc_d##D##_lap0(LiveParams& P, const ltc::TcExtra& X, cudaStream_t s); \
  int launch_live_tc_d##D##_lap1(LiveParams& P, const ltc::TcExtra& X, cudaStream_t s);
WF_DECL_LIVE_TC(2) WF_DECL_LIVE_TC(3) WF_DECL_LIVE_TC(4)
#undef WF_DECL_LIVE_TC

}  // namespace wf
