// Weight gradient of the dense layer products on the tensor cores (sm_100a only):
//
//   C[m, n] = sum_k A[k, m] * B[k, n]            (dW = dHᵀ·x, layers/layers.py:61 autograd; K = #entities)
//
// Both operands are row-major [K, cols] — the reduction index is the SLOW one — so they are fed to
// tcgen05.mma (kind::tf32) as MN-major operands straight from the same hi/lo split arrays the forward and dx
// products use (eg_split_tf32): no transposed copy is ever made.  TMA box = 32 columns (one 128-byte swizzle
// span) x BK rows; a tile is a row of such boxes (MN atoms LBO apart, the 4-row K atoms of a box SBO apart).
// fp32 accuracy from 3xTF32 (hi·hi + hi·lo + lo·hi into one TMEM accumulator).  The 300 x 300 output has only
// 3 x 2 tiles, so K is split over the grid (split-K): every CTA reduces its K range into TMEM and writes one
// fp32 partial tile; gemm_tn_reduce_kernel adds the partials in split order (deterministic).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace eg {
namespace tn {

using namespace eg::tc;

constexpr int BM = 128;            // output rows per tile (columns of A)      (UMMA M)
constexpr int BN = 160;            // output columns per tile (columns of B)   (UMMA N, multiple of 16)
constexpr int BK = 16;             // reduction rows per stage = two 8-row K atoms
constexpr int UK = 8;              // K per tcgen05.mma for tf32
constexpr int BOXC = 32;           // columns per TMA box = 128 bytes = one swizzle span
constexpr int BOX_BYTES = BK * BOXC * 4;                 // 2 KB, 1024-byte aligned
constexpr int A_BOXES = BM / BOXC, B_BOXES = BN / BOXC;  // 4, 5
constexpr int A_TILE_BYTES = A_BOXES * BOX_BYTES;        // 8 KB
constexpr int B_TILE_BYTES = B_BOXES * BOX_BYTES;        // 10 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // 36 KB (hi + lo of both)
constexpr int STAGES = 5;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;     // two accumulator buffers at columns 0 and 256
constexpr int FLUSH_KB = 8;        // k-blocks (128 reduction rows) accumulated in TMEM before the epilogue folds them
                                   // into fp32 registers: the tensor core truncates on every accumulate (~3e-8
                                   // relative bias per MMA), so chains are kept to 3 * 16 MMAs
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;

// MN-major tf32 operand.  The only shared-memory layout the tensor core accepts for it is "128-byte swizzle with a
// 32-byte base" (UMMA LayoutType 1 = TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: 32-byte chunks of a 128-byte row are
// XOR-ed with the row index mod 4): ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in elements — MN atoms (= TMA boxes)
// BOX_BYTES apart, K atoms (4 rows of 128 B) 512 B apart.  With any other layout type the MMA returns zeros.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  desc |= (uint64_t)(BOX_BYTES >> 4) << 16;             // leading byte offset: next MN atom
  desc |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next K atom (4 rows)
  desc |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  desc |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return desc;
}

// kind::tf32, fp32 accumulate, A and B MN-major (bits 15, 16), M x N
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct Params {
  int64_t K;            // reduction length (rows of A and B)
  int m, n;             // output shape
  int k_blocks_per_split;
  float* partial;       // [splits][m_tiles*BM][n_tiles*BN]
  int ldp;              // n_tiles * BN
  int64_t split_stride; // m_tiles*BM * ldp
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;    // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int64_t kb_total = (p.K + BK - 1) / BK;
  const int64_t kb_begin = (int64_t)blockIdx.z * p.k_blocks_per_split;
  const int64_t kb_end = min(kb_total, kb_begin + p.k_blocks_per_split);
  const int n_kb = (int)max(kb_end - kb_begin, (int64_t)0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const int krow = (int)((kb_begin + kb) * BK);
#pragma unroll
        for (int bx = 0; bx < A_BOXES; ++bx) {
          tma_load_2d(st + bx * BOX_BYTES, &map_a_hi, &full[s], m0 + bx * BOXC, krow);
          tma_load_2d(st + A_TILE_BYTES + bx * BOX_BYTES, &map_a_lo, &full[s], m0 + bx * BOXC, krow);
        }
#pragma unroll
        for (int bx = 0; bx < B_BOXES; ++bx) {
          tma_load_2d(st + 2 * A_TILE_BYTES + bx * BOX_BYTES, &map_b_hi, &full[s], n0 + bx * BOXC, krow);
          tma_load_2d(st + 2 * A_TILE_BYTES + B_TILE_BYTES + bx * BOX_BYTES, &map_b_lo, &full[s], n0 + bx * BOXC, krow);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {                                                                   // whole warp, one elected lane issues (elect_one)
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc_mn(BM, BN);
      int s = 0; uint32_t ph = 0;
      const int n_chunks = (n_kb + FLUSH_KB - 1) / FLUSH_KB;
      for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        mbar_wait(&tempty[buf], (uint32_t)(((c >> 1) & 1) ^ 1));      // epilogue has drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tbase + (uint32_t)(buf * 256);
        const int kb_hi = min(n_kb, (c + 1) * FLUSH_KB);
        for (int kb = c * FLUSH_KB; kb < kb_hi; ++kb) {
          mbar_wait(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint32_t koff = (uint32_t)k * 1024u;             // next 8 rows (two K atoms) inside every box
              const uint64_t a_hi = make_desc_mn(st + koff);
              const uint64_t a_lo = make_desc_mn(st + A_TILE_BYTES + koff);
              const uint64_t b_hi = make_desc_mn(st + 2 * A_TILE_BYTES + koff);
              const uint64_t b_lo = make_desc_mn(st + 2 * A_TILE_BYTES + B_TILE_BYTES + koff);
              // small cross terms first, the dominant hi·hi last: what truncation there is hits the small terms
              umma_tf32(tmem_d, a_hi, b_lo, idesc, (kb != c * FLUSH_KB) || k != 0);
              umma_tf32(tmem_d, a_lo, b_hi, idesc, 1);
              umma_tf32(tmem_d, a_hi, b_hi, idesc, 1);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull[buf]);
        __syncwarp();
      }
    }
  } else {
    // epilogue: TMEM lane = output row; 5 chunks of 32 columns -> this split's partial tile
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float* dst = p.partial + (int64_t)blockIdx.z * p.split_stride + (int64_t)(m0 + row) * p.ldp + n0;
    float acc[BN];
#pragma unroll
    for (int e = 0; e < BN; ++e) acc[e] = 0.f;
    const int n_chunks = (n_kb + FLUSH_KB - 1) / FLUSH_KB;
    for (int c = 0; c < n_chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&tfull[buf], (uint32_t)((c >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256);
#pragma unroll
      for (int cc = 0; cc < BN / 32; ++cc) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)(cc * 32), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[cc * 32 + e] += v[e];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
#pragma unroll
    for (int e = 0; e < BN; e += 4)
      *reinterpret_cast<float4*>(dst + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

__global__ void gemm_tn_reduce_kernel(const float* __restrict__ partial, int splits, int64_t split_stride, int ldp,
                                      int m, int n, float* __restrict__ out, int64_t ldo) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)m * n) return;
  const int i = (int)(idx / n), j = (int)(idx - (int64_t)i * n);
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[(int64_t)s * split_stride + (int64_t)i * ldp + j];
  out[(int64_t)i * ldo + j] = acc;
}

// [K, cols_pad] fp32 row-major; box = 32 columns x BK rows, 128-byte swizzle on 32-byte chunks, OOB reads as zero.
static int make_map_mn(CUtensorMap* map, const float* base, int64_t K, int cols, int cols_pad) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return EG_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)cols_pad * 4};
  cuuint32_t box[2] = {(cuuint32_t)BOXC, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EG_OK : EG_ERR_INVALID;
}

struct Plan {
  int m_tiles, n_tiles, splits, k_blocks_per_split;
  size_t ws_bytes;
};
static Plan plan(int64_t K, int m, int n) {
  Plan pl;
  pl.m_tiles = (int)ceil_div((int64_t)m, (int64_t)BM);
  pl.n_tiles = (int)ceil_div((int64_t)n, (int64_t)BN);
  const int64_t kb_total = std::max<int64_t>(1, ceil_div(K, (int64_t)BK));
  int64_t splits = std::max<int64_t>(1, (int64_t)kNumSMs / ((int64_t)pl.m_tiles * pl.n_tiles));
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, kb_total / 8));     // at least 8 k-blocks per CTA
  pl.k_blocks_per_split = (int)ceil_div(kb_total, splits);
  pl.splits = (int)ceil_div(kb_total, (int64_t)pl.k_blocks_per_split);
  pl.ws_bytes = sizeof(float) * (size_t)pl.splits * (size_t)pl.m_tiles * BM * (size_t)pl.n_tiles * BN;
  return pl;
}

}  // namespace tn
}  // namespace eg

extern "C" {

size_t eg_gemm_tn_3xtf32_workspace_bytes(int64_t K, int m, int n) {
  if (K <= 0 || m <= 0 || n <= 0) return 0;
  return eg::tn::plan(K, m, n).ws_bytes;
}

int eg_gemm_tn_3xtf32(const float* A_hi, const float* A_lo, int m, int lda, const float* B_hi, const float* B_lo,
                      int n, int ldb, int64_t K, void* ws, size_t ws_bytes, float* out, int64_t ldo,
                      eg_stream_t stream_) {
  using namespace eg;
  using namespace eg::tn;
  if (K < 0 || m <= 0 || n <= 0 || lda < m || ldb < n || ldo < n || (lda % 4) || (ldb % 4)) return EG_ERR_INVALID;
  if (!A_hi || !A_lo || !B_hi || !B_lo || !out) return EG_ERR_INVALID;
  if (((uintptr_t)A_hi | (uintptr_t)A_lo | (uintptr_t)B_hi | (uintptr_t)B_lo) & 15) return EG_ERR_INVALID;
  if (K >= (1ll << 31)) return EG_ERR_UNSUPPORTED;
  cudaStream_t s = as_stream(stream_);
  if (K == 0) {
    EG_CUDA(cudaMemset2DAsync(out, sizeof(float) * (size_t)ldo, 0, sizeof(float) * (size_t)n, (size_t)m, s));
    return EG_OK;
  }
  const Plan pl = plan(K, m, n);
  if (!ws || ws_bytes < pl.ws_bytes) return EG_ERR_WORKSPACE;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map_mn(&ma_hi, A_hi, K, m, lda))) return rc;
  if ((rc = make_map_mn(&ma_lo, A_lo, K, m, lda))) return rc;
  if ((rc = make_map_mn(&mb_hi, B_hi, K, n, ldb))) return rc;
  if ((rc = make_map_mn(&mb_lo, B_lo, K, n, ldb))) return rc;
  Params p;
  p.K = K; p.m = m; p.n = n;
  p.k_blocks_per_split = pl.k_blocks_per_split;
  p.partial = reinterpret_cast<float*>(ws);
  p.ldp = pl.n_tiles * BN;
  p.split_stride = (int64_t)pl.m_tiles * BM * p.ldp;
  static PerDeviceOnce attr_once;
  EG_SET_SMEM_ONCE(attr_once,
                   EG_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  dim3 grid((unsigned)pl.m_tiles, (unsigned)pl.n_tiles, (unsigned)pl.splits);
  gemm_tn_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  EG_LAUNCHED();
  gemm_tn_reduce_kernel<<<(unsigned)ceil_div((int64_t)m * n, (int64_t)256), 256, 0, s>>>(
      p.partial, pl.splits, p.split_stride, p.ldp, m, n, out, ldo);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
