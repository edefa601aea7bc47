// Tensor-core GEMMs of the training step (train.cu): the conditioner layers of the layer-wise jet formulation are plain
// [rows, K] x [K, N] products over rows = (walker, jet component), HBM bound once the FLOPs are off the CUDA cores
// (512 B of traffic per row of a 64-wide layer against 8 kFLOP).  Reference: the stax.serial(Dense, Tanh, Dense, Tanh, Dense)
// conditioners of model_factory.py:21-35 under value_and_grad(loss_fn_efficient), vqmc.py:193-221.
//
//   lin_tc_kernel<KP, NT>:  C[R][Nc] = A[R][Kc] * B (+ bias on the value rows),  Kc <= 64 KP,  Nc <= NT
//
// One CTA per SM, split into independent TEAMS of 128 threads (one tile of 128 rows each; they take turns on the tensor pipe,
// see the ticket in the kernel).  Per tile and 64-wide k-part a team
//   * reads its rows with coalesced 128-bit loads (thread t owns 16-byte unit t + 128 j of the row-major tile),
//   * splits every value into two TF32-exact planes (hi = rna(a), lo = a - hi) and stores them as the K-major
//     SWIZZLE_128B operand image the tensor core reads (row r of k-block b at b 16 KB + r 128 B, 16-byte units XOR (r & 7);
//     a quarter-warp writes one full 128-byte line: no bank conflicts),
//   * one warp issues 8 k-steps x 3 tcgen05.mma kind::tf32 ("3xTF32": A_hi B_hi + A_hi B_lo + A_lo B_hi, fp32 accumulate
//     in tensor memory) against the weight image (hi / lo planes, packed once per step by pack_b_kernel, fetched by one
//     cp.async.bulk per CTA) and commits to the team's mbarrier,
//   * the next tile's loads are issued before the team waits, so they fly under the MMAs and the epilogue,
//   * epilogue: tcgen05.ld (thread = row) -> bias -> padded shared staging (aliasing the dead A planes) -> coalesced 128-bit
//     stores.
// The result of a row does not depend on its position in a tile or on the grid, so chunked / sharded calls agree bitwise.
#pragma once
#include "live_tc.cuh"

namespace wf {
namespace ttc {

using namespace wf::ltc;

constexpr int TEAM = 128;
constexpr int A_PLANE = 128 * 64 * 4;           // 32 KB: [2 k-blocks][128 rows][128 B]
constexpr int A_TEAM = 2 * A_PLANE;             // hi | lo
constexpr int STAGE_LD = 68;                    // floats per staged output row (64 + 4: conflict-free 128-bit row-per-thread stores)
__host__ __device__ constexpr int teams_of(int KP, int NT) { return (KP == 1 && NT == 64) ? 3 : 2; }
__host__ __device__ constexpr int b_bytes(int KP, int NT) { return KP * NT * 512; }      // KP x (hi | lo) x [2][NT][128 B]
__host__ __device__ constexpr int img_floats(int KP, int NT) { return b_bytes(KP, NT) / 4; }
__host__ __device__ constexpr int smem_bytes(int KP, int NT) { return 1024 + b_bytes(KP, NT) + teams_of(KP, NT) * A_TEAM + 512 + 128; }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, M = 128, both operands K-major SWIZZLE_128B; whole converged warp, one lane elected
__device__ __forceinline__ void umma_tf32_ss_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// round-to-nearest (ties away) to TF32 on the integer pipe: cvt.rna.tf32.f32 issues on the quarter-rate XU pipe, and the operand
// staging of these kernels converts every element once
__device__ __forceinline__ float tf32_rna_i(float a) { return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u); }
// mbarrier wait for the warp-specialised pipelines below: a failed try_wait backs off with nanosleep, so that the roles that
// are ahead (producer, MMA issuer, the faster team) do not flood the shared-memory pipe with SYNCS polls while the others work
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; ; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(40);
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- weight images
struct PackJobs {
  static constexpr int MAX = 4 * (WF_MAX_LAYERS + 1);
  const float* src[MAX];
  float* dst[MAX];
  int Kc[MAX], Nc[MAX], KP[MAX], NT[MAX], trans[MAX];
};
// image element (n, k): k-part k / 64, then plane, then k-block (k % 64) / 32, row n, swizzled 16-byte unit.
// trans == 0: B is [Kc][Nc] (b(n, k) = B[k Nc + n]); trans == 1: B is [Nc][Kc]
__global__ void pack_b_kernel(const __grid_constant__ PackJobs jobs) {
  const int j = blockIdx.y;
  const float* __restrict__ B = jobs.src[j];
  float* __restrict__ img = jobs.dst[j];
  const int Kc = jobs.Kc[j], Nc = jobs.Nc[j], KP = jobs.KP[j], NT = jobs.NT[j], trans = jobs.trans[j];
  const int total = KP * 64 * NT;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = trans ? i / (KP * 64) : i % NT, k = trans ? i % (KP * 64) : i / NT;      // coalesced on the source side
    const float v = (n < Nc && k < Kc) ? (trans ? B[(int64_t)n * Kc + k] : B[(int64_t)k * Nc + n]) : 0.f;
    const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
    const int kp = k >> 6, kl = k & 63, kb = kl >> 5, kk = kl & 31;
    const int plane = NT * 64;                                                             // floats per plane
    const int off = kp * 2 * plane + kb * NT * 32 + n * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
    img[off] = hi;
    img[off + plane] = lo;
  }
}

// ---------------------------------------------------------------------------------------------- C = A B (+ bias)
template <int KP, int NT>
__global__ void __launch_bounds__(teams_of(KP, NT) * TEAM, 1)
lin_tc_kernel(const float* __restrict__ A, const float* __restrict__ Bimg, const float* __restrict__ bias, float* __restrict__ C,
              int64_t R, int Kc, int Nc, int G) {
  constexpr int TEAMS = teams_of(KP, NT), BB = b_bytes(KP, NT), HALVES = NT / 64;
  constexpr uint32_t TMEM_COLS = TEAMS * NT <= 128 ? 128 : 256;
  constexpr uint32_t idesc = instr_desc_tf32(NT);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* Bs = sm;
  const int tid = threadIdx.x, team = tid >> 7, ttid = tid & 127, tw = ttid >> 5;
  unsigned char* As = Bs + BB + team * A_TEAM;
  float* stage = reinterpret_cast<float*>(As);
  float* bias_s = reinterpret_cast<float*>(Bs + BB + TEAMS * A_TEAM);           // [NT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + BB + TEAMS * A_TEAM + 512);  // [0] weights, [1 + team] MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile uint32_t* turn = tmem_slot + 1;                   // ticket of the team whose MMAs go next (see below)

  if (tid < 32) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 0) {
    *turn = 0;
    mbar_init(&bars[0], 1);
    for (int i = 0; i < TEAMS; ++i) mbar_init(&bars[1 + i], 1);
    mbar_fence_init();
    mbar_expect_tx(&bars[0], BB);
    bulk_g2s(Bs, Bimg, BB, &bars[0]);
  }
  for (int i = tid; i < NT; i += TEAMS * TEAM) bias_s[i] = (bias && i < Nc) ? bias[i] : 0.f;
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_tmem = tmem_base + (uint32_t)(team * NT);
  const uint32_t d_lane = d_tmem + ((uint32_t)(tw * 32) << 16);

  const int64_t n_tiles = (R + 127) / 128;
  const int urow = Kc >> 2;                                  // 16-byte units per input row
  const int64_t stride = (int64_t)gridDim.x * TEAMS;
  int64_t tile = (int64_t)blockIdx.x * TEAMS + team;
  float4 raw[16];
  auto load = [&](int64_t t, int kp) {
    const int64_t r0 = t * 128;
    const int rows = (int)min((int64_t)128, R - r0);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int u = ttid + TEAM * j, row = u >> 4, cu = kp * 16 + (u & 15);
      raw[j] = (row < rows && cu < urow) ? __ldg(reinterpret_cast<const float4*>(A + (r0 + row) * Kc) + cu) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (tile < n_tiles) load(tile, 0);
  bool first = true;
  uint32_t par = 0;
  int it = 0;
  for (; tile < n_tiles; tile += stride, ++it) {
    const int64_t r0 = tile * 128;
    const int rows = (int)min((int64_t)128, R - r0);
    // teams of this CTA that have a tile in this round (only the last round can be short)
    const int m_round = (int)min((int64_t)TEAMS, n_tiles - ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TEAMS);
#pragma unroll 1
    for (int kp = 0; kp < KP; ++kp) {
      // ---- split + stage the operand image of this k-part
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int u = ttid + TEAM * j, row = u >> 4, c = u & 15;
        const float4 v = raw[j];
        float4 hi, lo;
        hi.x = tf32_rna_i(v.x); hi.y = tf32_rna_i(v.y); hi.z = tf32_rna_i(v.z); hi.w = tf32_rna_i(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        const int off = (c >> 3) * (128 * 128) + row * 128 + (((c & 7) ^ (row & 7)) << 4);
        *reinterpret_cast<float4*>(As + off) = hi;
        *reinterpret_cast<float4*>(As + A_PLANE + off) = lo;
      }
      fence_proxy_async_smem();
      bar_sync(1 + team, TEAM);
      if (tw == 0) {
        if (first) mbar_wait_guard(&bars[0], 0);              // weight image landed
        // The teams take turns on the tensor pipe.  Left alone they run in LOCKSTEP (a clock64 timeline of one CTA,
        // profiles/r02_train_lin_tc_timeline.txt: all three stage, issue, read TMEM and store at the same time; every phase is
        // bound by a resource they then share -- shared-memory bandwidth, the MMA queue, TMEM reads, the store path -- and a
        // round of three tiles takes the SUM of the contended phases, 13.5 k cycles).  Issuing the 24 MMAs of one team as one
        // block, in ticket order, lets the first team leave the MMA phase after a third of the time: -10 % per layer.
        // (Making the staging and epilogue phases exclusive as well was slower: alone they are latency bound.)
        const uint32_t ticket = (uint32_t)((it * KP) * TEAMS + kp * m_round + team);
        if ((ttid & 31) == 0)
          while (*turn != ticket) __nanosleep(20);
        __syncwarp();
        fence_after();
        uint64_t ah = smem_desc_sw128(As), al = smem_desc_sw128(As + A_PLANE);
        uint64_t bh = smem_desc_sw128(Bs + kp * (NT * 512)), bl = smem_desc_sw128(Bs + kp * (NT * 512) + NT * 256);
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks) {
          umma_tf32_ss_warp(d_tmem, ah, bh, idesc, (kp | ks) != 0 ? 1u : 0u);
          umma_tf32_ss_warp(d_tmem, ah, bl, idesc, 1u);
          umma_tf32_ss_warp(d_tmem, al, bh, idesc, 1u);
          const bool nb = (ks & 3) == 3;                      // next 32-float k-block
          const uint64_t aadv = nb ? (uint64_t)((128 * 128 - 96) >> 4) : 2ull, badv = nb ? (uint64_t)((NT * 128 - 96) >> 4) : 2ull;
          ah += aadv; al += aadv; bh += badv; bl += badv;
        }
        umma_commit_warp(&bars[1 + team]);
        if ((ttid & 31) == 0) *turn = ticket + 1;
      }
      first = false;
      // ---- the next loads fly under the MMAs and the epilogue
      if (kp + 1 < KP) load(tile, kp + 1);
      else if (tile + stride < n_tiles) load(tile + stride, 0);
      mbar_wait_guard(&bars[1 + team], par);
      par ^= 1;
      fence_after();
    }
    // ---- epilogue: accumulator rows -> (+ bias on value rows) -> staging -> coalesced stores
    const float vmask = (((uint32_t)(r0 % G) + (uint32_t)ttid) % (uint32_t)G) == 0 ? 1.f : 0.f;
#pragma unroll 1
    for (int h = 0; h < HALVES; ++h) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[32];
        tmem_ld32(d_lane + (uint32_t)(h * 64 + ch * 32), v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + h * 64 + ch * 32 + 4 * q);
          float4 o;
          o.x = fmaf(vmask, b.x, v[4 * q]); o.y = fmaf(vmask, b.y, v[4 * q + 1]);
          o.z = fmaf(vmask, b.z, v[4 * q + 2]); o.w = fmaf(vmask, b.w, v[4 * q + 3]);
          *reinterpret_cast<float4*>(stage + ttid * STAGE_LD + ch * 32 + 4 * q) = o;
        }
      }
      fence_before();
      bar_sync(1 + team, TEAM);
      const int uh = min(16, (Nc - 64 * h) >> 2);            // 16-byte units of this 64-column half that exist
      const int total = rows * uh;
      for (int u = ttid; u < total; u += TEAM) {
        const int row = uh == 16 ? u >> 4 : u / uh, c = u - row * uh;
        const float4 o = *reinterpret_cast<const float4*>(stage + row * STAGE_LD + 4 * c);
        *reinterpret_cast<float4*>(C + (r0 + row) * Nc + 64 * h + 4 * c) = o;
      }
      bar_sync(1 + team, TEAM);
    }
  }
  fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int KP, int NT>
int launch_lin_tc_t(const float* A, const float* img, const float* bias, float* C, int64_t R, int Kc, int Nc, int G, cudaStream_t s) {
  static bool attr[WF_MAX_DEVICES] = {};
  const int dev = current_device();
  if (!attr[dev]) {
    WF_CUDA(cudaFuncSetAttribute(lin_tc_kernel<KP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(KP, NT)));
    attr[dev] = true;
  }
  constexpr int TEAMS = teams_of(KP, NT);
  const int64_t tiles = (R + 127) / 128, want = (tiles + TEAMS - 1) / TEAMS;
  const int grid = (int)(want < num_sms() ? want : num_sms());
  lin_tc_kernel<KP, NT><<<grid, TEAMS * TEAM, smem_bytes(KP, NT), s>>>(A, img, bias, C, R, Kc, Nc, G);
  WF_LAUNCH_CHECK();
  return WF_OK;
}

// shapes the tensor-core path takes (everything else stays on linear_kernel)
inline bool lin_tc_ok(int64_t R, int Kc, int Nc) {
  return R >= (int64_t)128 * num_sms() && (Kc & 3) == 0 && (Nc & 3) == 0 && Kc >= 32 && Kc <= 128 && Nc >= 32 && Nc <= 128 &&
         !(Kc > 64 && Nc > 64);
}
inline int lin_kp(int Kc) { return Kc > 64 ? 2 : 1; }
inline int lin_nt(int Nc) { return Nc > 64 ? 128 : 64; }

inline int launch_lin_tc(const float* A, const float* img, const float* bias, float* C, int64_t R, int Kc, int Nc, int G, cudaStream_t s) {
  if (Kc > 64) return launch_lin_tc_t<2, 64>(A, img, bias, C, R, Kc, Nc, G, s);
  if (Nc > 64) return launch_lin_tc_t<1, 128>(A, img, bias, C, R, Kc, Nc, G, s);
  return launch_lin_tc_t<1, 64>(A, img, bias, C, R, Kc, Nc, G, s);
}

// ---------------------------------------------------------------------------------------------- dW = X^T dY (+ bias gradient)
//   wgrad_tc_kernel:  partial[cta][k][n] = sum over the CTA's rows of X[r][k] dY[r][n]  (k < Kc, Kc = 64 or a multiple of 4 below),
//                     partial[cta][Kc][n] = sum over its value rows (r % G == 0) of dY[r][n]
// The reduction runs over ROWS, so both MMA operands have to be row-contiguous ("K-major" with K = rows) while the arrays
// in HBM are feature-contiguous.  The transposition costs nothing extra:
//   * a producer thread streams 32-row sub-tiles of X and dY, as they lie in HBM, into a shared-memory ring of up to 8
//     stages (two cp.async.bulk per sub-tile; ~160 KB in flight per SM -- with 3 stages of 64 rows the kernel was bound by
//     the round trip of the ring, not by HBM);
//   * "A team" (8 warps, thread n owns output column n = TMEM lane n, two warps per lane quarter) reads COLUMN n of the dY sub-tile (consecutive
//     lanes -> consecutive floats: no bank conflicts), splits it into TF32 hi / lo planes and writes them with
//     tcgen05.st into tensor memory: the A operand of the .ts MMA form (lane = M index, one column per row r);
//   * "B team" (4 warps, thread (k, half)) reads column k of the X sub-tile and writes row k of the K-major
//     SWIZZLE_128B image (the XOR swizzle makes the row-per-lane 128-bit stores conflict free); image row 64 is the
//     value-row indicator, so the bias gradient falls out of the same MMAs as output column 64;
//   * one warp issues 4 x 3 tcgen05.mma (M = 128, N = 80, K = 8, 3xTF32) per sub-tile into ONE accumulator
//     D[n][k] (80 TMEM columns) that lives across all sub-tiles of the CTA; operands are triple buffered and handed
//     over with mbarriers (tcgen05.commit frees a stage), so loads, transposition and MMAs of neighbouring sub-tiles overlap.
// The per-CTA partial sums are then added in a fixed order by wgrad_reduce_kernel: deterministic gradients.
constexpr int WG_ROWS = 32;                     // rows per sub-tile = one 32-float k-block
constexpr int WG_MAX_STAGES = 8;                // raw ring (as many as fit: 7 at Kc = 64, Nc = 116)
constexpr int WG_OPS = 3;                       // operand stages (A planes in TMEM, B image in shared memory)
constexpr int WG_NB = 80;                       // image rows: 64 input features + value-row indicator, padded to N % 16 == 0
constexpr int WG_BPLANE = WG_NB * 128, WG_BSTAGE = 2 * WG_BPLANE;
constexpr int WG_SMEM = 227 * 1024;
constexpr int WG_RING = WG_SMEM - 1024 - WG_OPS * WG_BSTAGE - 512;
constexpr int WG_THREADS = 448;              // 8 A-team warps, 4 B-team warps, producer, MMA issuer
__host__ __device__ inline int wg_stage_bytes(int Kc, int Nc) { return (WG_ROWS * (Kc + Nc) * 4 + 127) & ~127; }

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const float* __restrict__ X, const float* __restrict__ dY, float* __restrict__ partial, int64_t R, int Kc, int Nc, int G) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* Bimg = sm;                                        // [WG_OPS][hi | lo][80 rows][128 B]
  unsigned char* raw = sm + WG_OPS * WG_BSTAGE;                    // [S][X rows | dY rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + WG_RING);
  uint64_t *raw_full = bars, *raw_empty = bars + 8, *ops_full = bars + 16, *ops_empty = bars + 20, *done = bars + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stage_bytes = wg_stage_bytes(Kc, Nc), x_bytes = WG_ROWS * Kc * 4;
  const int S = min(WG_MAX_STAGES, WG_RING / stage_bytes);

  for (int i = tid; i < WG_OPS * WG_BSTAGE / 16; i += WG_THREADS) reinterpret_cast<float4*>(Bimg)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < WG_MAX_STAGES; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 12); }
    for (int i = 0; i < WG_OPS; ++i) { mbar_init(&ops_full[i], 12); mbar_init(&ops_empty[i], 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_tmem = tmem_base + 256;                         // accumulator D[n][k]: 80 columns
  const int64_t n_sub = (R + WG_ROWS - 1) / WG_ROWS;
  const int n_it = (int)((n_sub - blockIdx.x + gridDim.x - 1) / gridDim.x);        // sub-tiles of this CTA (>= 1)

  if (warp == 12) {
    // ---- producer: two bulk copies per sub-tile, S sub-tiles in flight
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < n_it; ++it) {
        const int64_t sub = blockIdx.x + (int64_t)it * gridDim.x;
        const int rows = (int)min((int64_t)WG_ROWS, R - sub * WG_ROWS);
        if (it >= S) mbar_wait_backoff(&raw_empty[s], ph ^ 1);
        const uint32_t bx = (uint32_t)rows * (uint32_t)Kc * 4u, by = (uint32_t)rows * (uint32_t)Nc * 4u;
        mbar_expect_tx(&raw_full[s], bx + by);
        bulk_g2s(raw + s * stage_bytes, X + sub * WG_ROWS * Kc, bx, &raw_full[s]);
        bulk_g2s(raw + s * stage_bytes + x_bytes, dY + sub * WG_ROWS * Nc, by, &raw_full[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 13) {
    // ---- MMA issue: 4 k-steps x 3 per sub-tile into the CTA-lifetime accumulator
    constexpr uint32_t idesc = instr_desc_tf32(WG_NB);
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < n_it; ++it) {
      mbar_wait_backoff(&ops_full[st], ph);
      fence_after();
      const uint32_t a_hi = tmem_base + (uint32_t)(st * 64), a_lo = a_hi + 32;
      uint64_t bh = smem_desc_sw128(Bimg + st * WG_BSTAGE), bl = smem_desc_sw128(Bimg + st * WG_BSTAGE + WG_BPLANE);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t acol = (uint32_t)(ks * 8);
        umma_tf32_ts_warp(d_tmem, a_hi + acol, bh, idesc, (it | ks) != 0 ? 1u : 0u);
        umma_tf32_ts_warp(d_tmem, a_hi + acol, bl, idesc, 1u);
        umma_tf32_ts_warp(d_tmem, a_lo + acol, bh, idesc, 1u);
        bh += 2; bl += 2;                                          // 32 bytes further inside the 128-byte swizzle span
      }
      umma_commit_warp(&ops_empty[st]);
      if (++st == WG_OPS) { st = 0; ph ^= 1; }
    }
    umma_commit_warp(done);
  } else if (warp < 8) {
    // ---- A team: column n of dY -> tensor memory (hi | lo), lane n.  Two warps per lane quarter (rows 0..15 / 16..31 of the
    // sub-tile): with one warp per scheduler the dependent cvt / sub chains were the bottleneck of the whole kernel.
    const int q = warp & 3, c = warp >> 2, n = 32 * q + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool live = n < Nc;
    int s = 0, st = 0; uint32_t ph = 0, pho = 0;
    for (int it = 0; it < n_it; ++it) {
      const int64_t sub = blockIdx.x + (int64_t)it * gridDim.x;
      const int rows = (int)min((int64_t)WG_ROWS, R - sub * WG_ROWS);
      mbar_wait_backoff(&raw_full[s], ph);
      if (it >= WG_OPS) { mbar_wait_backoff(&ops_empty[st], pho ^ 1); fence_after(); }
      const float* y = reinterpret_cast<const float*>(raw + s * stage_bytes + x_bytes) + (live ? n : 0) + 16 * c * Nc;
      uint32_t hi[16], lo[16];
      if (live && rows == WG_ROWS) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = y[j * Nc];
          const float h = tf32_rna_i(v);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(v - h);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = (live && 16 * c + j < rows) ? y[j * Nc] : 0.f;
          const float h = tf32_rna_i(v);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(v - h);
        }
      }
      tmem_st16(lane_addr + (uint32_t)(st * 64 + 16 * c), hi);
      tmem_st16(lane_addr + (uint32_t)(st * 64 + 32 + 16 * c), lo);
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ops_full[st]); mbar_arrive(&raw_empty[s]); }
      if (++s == S) { s = 0; ph ^= 1; }
      if (++st == WG_OPS) { st = 0; pho ^= 1; }
    }
    if (c == 0) {
    // ---- accumulator -> partial[cta][k][n]
    mbar_wait_backoff(done, 0);
    fence_after();
    float* out = partial + (int64_t)blockIdx.x * (Kc + 1) * Nc;
#pragma unroll 1
    for (int cc = 0; cc < 5; ++cc) {
      float v[16];
      tmem_ld16(lane_addr + (uint32_t)(256 + 16 * cc), v);
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = 16 * cc + j;
          if (k < Kc) out[k * Nc + n] = v[j];
          else if (k == 64) out[Kc * Nc + n] = v[j];
        }
      }
    }
    }
  } else {
    // ---- B team: column k of X -> row k of the K-major image (hi | lo), 16 rows per thread; image row 64 = value-row indicator
    const int t = tid - 256, k = t & 63, half = t >> 6;
    int s = 0, st = 0; uint32_t ph = 0, pho = 0;
    // (first row of the sub-tile) mod G, advanced per iteration in 32-bit arithmetic (a 64-bit modulo per sub-tile on the
    // critical path of this warp used to cost more than the transposition itself)
    uint32_t rem = (uint32_t)(((int64_t)blockIdx.x * WG_ROWS) % G);
    const uint32_t rem_step = (uint32_t)(((int64_t)gridDim.x * WG_ROWS) % G);
    for (int it = 0; it < n_it; ++it) {
      const int64_t sub = blockIdx.x + (int64_t)it * gridDim.x;
      const int rows = (int)min((int64_t)WG_ROWS, R - sub * WG_ROWS);
      mbar_wait_backoff(&raw_full[s], ph);
      if (it >= WG_OPS) mbar_wait_backoff(&ops_empty[st], pho ^ 1);
      const float* x = reinterpret_cast<const float*>(raw + s * stage_bytes) + k;
      unsigned char* img = Bimg + st * WG_BSTAGE;
      if (k < Kc) {                                     // image rows Kc .. 63 stay zero (narrow first layer)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = 4 * half + jj;
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = 4 * j + e;
            v[e] = r < rows ? x[r * Kc] : 0.f;
          }
          float4 hi, lo;
          hi.x = tf32_rna_i(v[0]); hi.y = tf32_rna_i(v[1]); hi.z = tf32_rna_i(v[2]); hi.w = tf32_rna_i(v[3]);
          lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
          const int off = k * 128 + ((j ^ (k & 7)) << 4);
          *reinterpret_cast<float4*>(img + off) = hi;
          *reinterpret_cast<float4*>(img + WG_BPLANE + off) = lo;
        }
      }
      if (t < WG_ROWS) {
        const int r = t;
        const float ind = (r < rows && ((rem + (uint32_t)r) % (uint32_t)G) == 0) ? 1.f : 0.f;
        *reinterpret_cast<float*>(img + 64 * 128 + ((r >> 2) << 4) + ((r & 3) << 2)) = ind;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ops_full[st]); mbar_arrive(&raw_empty[s]); }
      if (++s == S) { s = 0; ph ^= 1; }
      if (++st == WG_OPS) { st = 0; pho ^= 1; }
      rem = (rem + rem_step) % (uint32_t)G;
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

inline bool wgrad_tc_ok(int64_t R, int Kc, int Nc) {
  return R >= (int64_t)128 * num_sms() && Kc >= 4 && Kc <= 64 && (Kc & 3) == 0 && (Nc & 3) == 0 && Nc >= 4 && Nc <= 128;
}

// grid (= number of partial blocks written) is returned through *n_cta
inline int launch_wgrad_tc(const float* X, const float* dY, float* partial, int64_t R, int Kc, int Nc, int G, int max_cta, int* n_cta,
                           cudaStream_t s) {
  static bool attr[WF_MAX_DEVICES] = {};
  const int dev = current_device();
  if (!attr[dev]) {
    WF_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr[dev] = true;
  }
  const int64_t n_sub = (R + WG_ROWS - 1) / WG_ROWS;
  int grid = (int)(n_sub < num_sms() ? n_sub : num_sms());
  if (grid > max_cta) grid = max_cta;
  wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, s>>>(X, dY, partial, R, Kc, Nc, G);
  WF_LAUNCH_CHECK();
  *n_cta = grid;
  return WF_OK;
}

}  // namespace ttc
}  // namespace wf
