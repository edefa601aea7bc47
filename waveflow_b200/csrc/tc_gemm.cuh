// tcgen05 (5th-gen tensor core) GEMM main loop with float32-grade accuracy for the wide conditioner MLPs of the coupling
// flow (BASELINE config 5: width-512 FCNNs, K = 64 bins).
//
//   D[128 x NT] (TMEM, fp32)  =  A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T        ("3xTF32": a = a_hi + a_lo, both TF32-exact)
//
// A: activations [M][K] row-major, B: weights [N][K] row-major (i.e. already transposed), both K-major for the MMA.
// Tiles are staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B, 32 fp32 = 128 B along K per box) into a 2-stage ring;
// one elected thread issues tcgen05.mma.kind::tf32 (UMMA 128 x NT x 8), accumulators are double-buffered in TMEM so the
// epilogue of tile t overlaps the main loop of tile t+1.  Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM
// allocation), 2..5 = epilogue (thread <-> accumulator row).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"

namespace wf {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KB = 32;             // fp32 elements per k-block (one 128-byte swizzle span)
constexpr int UMMA_K = 8;          // K per tcgen05.mma.kind::tf32
constexpr int STAGES = 2;
constexpr int THREADS = 192;
constexpr int EPI_WARP0 = 2;

template <int NT>
struct Smem {
  static constexpr int A_BYTES = TILE_M * KB * 4;        // 16 KB
  static constexpr int B_BYTES = NT * KB * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;     // barriers + tmem pointer + alignment slack
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread <-> TMEM lane)
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(r);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw128(const void* p) {
  const uint64_t addr = (uint64_t)(smem_u32(p) >> 4) & 0x3FFFull;
  return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// TF32-exact split of a float: hi = round-to-nearest-even to 10 explicit mantissa bits, lo = TF32(a - hi)
__device__ __forceinline__ float tf32_rn(float a) {
  uint32_t u = __float_as_uint(a);
  u += 0x0FFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}

struct Maps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

// ---------------------------------------------------------------------------------------------- tensor maps (host)
inline PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// row-major [rows][cols] float32 matrix with leading dimension ld (elements), box = [box_rows][32 columns], 128-byte swizzle
inline int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows, int64_t ld = 0) {
  if (ld == 0) ld = cols;
  auto fn = encode_fn();
  if (!fn) return WF_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WF_OK : WF_ERR_INVALID_ARG;
}

inline int make_maps(Maps& m, const float* a_hi, const float* a_lo, int64_t M, int K, const float* b_hi, const float* b_lo, int64_t N, int nt) {
  int st;
  if ((st = make_map(&m.a_hi, a_hi, M, K, TILE_M)) != WF_OK) return st;
  if ((st = make_map(&m.a_lo, a_lo, M, K, TILE_M)) != WF_OK) return st;
  if ((st = make_map(&m.b_hi, b_hi, N, K, nt)) != WF_OK) return st;
  if ((st = make_map(&m.b_lo, b_lo, N, K, nt)) != WF_OK) return st;
  return WF_OK;
}


// Main loop + role dispatch.  Epi is a functor:  epi(m0, n_tile, row, tmem_row_addr)  called by every epilogue thread
// once per tile, where tmem_row_addr addresses column 0 of this thread's accumulator row (use tmem_ld32 / tmem_ld1).
template <int NT, class Epi>
__device__ __forceinline__ void gemm_mainloop(const Maps& maps, int64_t M, int K, int n_tiles, unsigned char* smem_raw, Epi& epi) {
  using S = Smem<NT>;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m_tiles = (M + TILE_M - 1) / TILE_M;
  const int64_t tiles = m_tiles * n_tiles;
  const int kblocks = K / KB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = (int)(tile / n_tiles) * TILE_M, n0 = (int)(tile % n_tiles) * NT;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) + 1) & 1));
          unsigned char* st = smem + (size_t)s * S::STAGE_BYTES;
          mbar_expect_tx(&full[s], (uint32_t)S::STAGE_BYTES);
          tma_load_2d(st, &maps.a_hi, kb * KB, m0, &full[s]);
          tma_load_2d(st + S::A_BYTES, &maps.a_lo, kb * KB, m0, &full[s]);
          tma_load_2d(st + 2 * S::A_BYTES, &maps.b_hi, kb * KB, n0, &full[s]);
          tma_load_2d(st + 2 * S::A_BYTES + S::B_BYTES, &maps.b_lo, kb * KB, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_tf32(NT);
      int64_t it = 0, t = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t) {
        const int a = (int)(t & 1);
        if (t >= 2) mbar_wait(&tempty[a], (uint32_t)(((t >> 1) + 1) & 1));
        fence_after();
        const uint32_t d = tmem_base + (uint32_t)(a * 256);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = (int)(it % STAGES);
          mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
          fence_after();
          unsigned char* st = smem + (size_t)s * S::STAGE_BYTES;
          const uint64_t da_hi = smem_desc_sw128(st), da_lo = smem_desc_sw128(st + S::A_BYTES);
          const uint64_t db_hi = smem_desc_sw128(st + 2 * S::A_BYTES), db_lo = smem_desc_sw128(st + 2 * S::A_BYTES + S::B_BYTES);
#pragma unroll
          for (int k = 0; k < KB / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);       // 32 bytes per k-step inside the swizzle span
            umma_tf32(d, da_hi + adv, db_hi + adv, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_tf32(d, da_hi + adv, db_lo + adv, idesc, 1u);
            umma_tf32(d, da_lo + adv, db_hi + adv, idesc, 1u);
          }
          umma_commit(&empty[s]);          // frees the stage once these MMAs have read it
        }
        umma_commit(&tfull[a]);            // accumulator complete
      }
    }
  } else {
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int64_t t = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t) {
      const int a = (int)(t & 1);
      mbar_wait(&tfull[a], (uint32_t)((t >> 1) & 1));
      fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256);
      epi((int)(tile / n_tiles) * TILE_M, (int)(tile % n_tiles), row, taddr);
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace wf
