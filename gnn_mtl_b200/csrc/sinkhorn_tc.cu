// (b) Fused Sinkhorn half-sweep on the 5th-gen tensor cores (sm_100a only).
//
//   lse[i] = log sum_j exp(pot[j] - cost(A_i, B_j) * inv_reg)
//
// The 128 x 256 tile of dot products A_i·B_j is produced by tcgen05.mma (kind::tf32)
// with both operands fed by TMA (cp.async.bulk.tensor, 64-byte swizzle, 4 stages) and the
// accumulator living in TMEM; fp32 accuracy comes from the 3xTF32 split
//   a·b ≈ a_hi·b_hi + a_hi·b_lo + a_lo·b_hi      (hi = tf32(x), lo = tf32(x - hi), eg_split_tf32)
// issued as three MMAs per k-step into the same accumulator.  The epilogue reads
// the accumulator back with tcgen05.ld (one TMEM lane = one row of A = one
// thread), turns dots into costs, and folds them into a per-thread online
// log-sum-exp — the I x J cost is never written anywhere.  Two accumulator
// buffers (2 x 256 TMEM columns) let the epilogue of tile t overlap the MMAs of
// tile t+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2-5 = epilogue (TMEM lane quadrant = warp_id % 4).
#include <cuda.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace eg {

int g_tune_tc_pair = 1;          // knob 19: the tcgen05 kernels of this file on CTA pairs (cta_group::2; default) or single CTAs
                                 // (same bits; fused half-sweep 30000^2 x 300: 2.24-2.27 ms against 2.32-2.36 ms,
                                 //  [200k, 300] x [600 | 300, 300]^T 0.40 / 0.23 ms against 0.445 / 0.255 ms)

namespace tc {

constexpr int BM = 128;          // rows of A per tile  (UMMA M)
constexpr int BN = 256;          // rows of B per tile  (UMMA N)
constexpr int BK = 16;           // fp32 elements per k-block = one 64-byte swizzle atom
constexpr int UK = 8;            // K per tcgen05.mma for tf32
constexpr int STAGES = 4;        // 4 x 48 KB: ~2300 MMA-cycles of lead time for every TMA load
constexpr int ROW_BYTES = BK * 4;           // smem row pitch of an operand tile (= swizzle span)
constexpr int A_TILE_BYTES = BM * BK * 4;   // 8 KB
constexpr int B_TILE_BYTES = BN * BK * 4;   // 16 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // 48 KB
static_assert(ROW_BYTES == 64 || ROW_BYTES == 128, "operand rows must span one 64- or 128-byte swizzle atom");
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * BN * 8 + 256 + 1024;  // + (norm,pot) staging + barriers + align

// K-major operand tile whose rows span exactly one swizzle atom (64 or 128 B): rows are ROW_BYTES
// apart, 8-row groups 8*ROW_BYTES apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr >> 4) & 0x3FFF);   // start address  [0,14)
  desc |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major)
  desc |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;   // stride byte offset: one 8-row group
  desc |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  desc |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;   // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return desc;
}

// kind::tf32, fp32 accumulate, A and B K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// |a|^2 + |b|^2 - 2 a.b cancels catastrophically for close pairs (the pairs that dominate the log-sum-exp of an
// alignment problem), turning the ~2e-6 relative error of the 3xTF32 dot into 1e-3 of the cost.  Pairs whose
// squared distance comes out below a quarter of |a|^2 + |b|^2 (cosine > 0.75) are re-evaluated from the fp32
// rows: sum (a_k - b_k)^2, or the exact dot for the cosine cost.  Rare by construction, so the cost is noise.
constexpr float kNearFrac = 0.25f;
// All 32 lanes evaluate ONE pair together: coalesced 16-byte loads of both rows, shuffle reduction.
__device__ __forceinline__ float exact_pair_warp(const float* __restrict__ a, const float* __restrict__ b, int d,
                                                 int want_dot, int lane) {
  float acc = 0.f;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int k = lane; k < d / 4; k += 32) {
      const float4 x = __ldg(a4 + k), y = __ldg(b4 + k);
      if (want_dot) {
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
      } else {
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
      }
    }
  } else {
    for (int k = lane; k < d; k += 32) {
      const float x = __ldg(a + k), y = __ldg(b + k);
      acc = want_dot ? fmaf(x, y, acc) : fmaf(x - y, x - y, acc);
    }
  }
  return warp_sum(acc);
}

struct Params {
  int bn;                 // MODE 2: columns per tile (multiple of 16, <= BN), chosen so that padding is small
  int64_t nA, nB;
  int k_blocks;            // ceil(d_pad / BK)
  int d_pad;               // operand row length (multiple of 8)
  int cost;
  float inv_reg;
  const float* normA;
  const float* normB;
  const float* pot_in;
  int tiles_per_split;     // B tiles handled by one CTA (grid.y = splits)
  float* part_m;           // [splits, nA]                       (MODE 0)
  float* part_s;
  const float* pot_a;      // f_i                                 (MODE 1)
  double* loss;            // sum_ij P_ij * cost_ij               (MODE 1)
  float* row_sum;          // sum_j P_ij, atomically accumulated  (MODE 1, nullable)
  const float* A_raw;      // fp32 rows [nA, d] / [nB, d]: close pairs are re-evaluated exactly from these (MODE 0/1)
  const float* B_raw;
  int d;
  int kb_split;            // k-blocks taken from the first A operand; the rest come from A2 (MODE 2: A = [A1 | A2])
  float* out1;             // C[:, 0:n1)   row stride ld1          (MODE 2)
  float* out2;             // C[:, n1:nB)  row stride ld2          (MODE 2, nullable)
  int64_t ld1, ld2, n1;
};

// MODE 0: per-row online log-sum-exp of  pot_in[j] - cost*inv_reg          -> part_m / part_s
// MODE 1: plan statistics with P_ij = exp(pot_a[i] + pot_in[j] - cost*inv_reg) -> loss, row_sum
// MODE 2: plain 3xTF32 GEMM  C = [A1 | A2] · Bᵀ + bias  (pot_in carries the bias)   -> out1 / out2
// MODE 3: the same GEMM with SHORT ACCUMULATION CHAINS: the tensor core truncates on every accumulate (~3e-8 relative
//         per MMA, a one-sided bias that grows with the chain: 2e-6 over the 114 MMAs of K = 304), so here every
//         k-block (6 MMAs) starts a fresh TMEM accumulator and the epilogue warps add the finished block into fp32
//         registers (round-to-nearest) while the next block runs in the other buffer.  Error ~3e-7: the level of an
//         fp32 SIMT product, which is what a ReLU-feeding product needs (layers/layers.py:32,61 feed F.relu).
//         Column tile <= kFlushBN so that a thread's row of accumulators stays in registers.
//
// PAIR: the same kernel on pairs of CTAs (a 2-CTA cluster = the two SMs of a TPC, tcgen05.mma.cta_group::2).  The pair
// computes a 256-row tile: every CTA loads its own 128 rows of A and HALF of the B tile (the tensor cores exchange the
// halves), which takes the operand fetch + TMA fill of a 128 x 256 x 8 MMA from 160 cycles of the shared-memory port
// (more than its 128 cycles of math: the single-CTA kernel is bound by that port) to 107.  Rank 0 issues the MMAs for
// both SMs; TMA bytes of both CTAs count on its `full` barriers; tcgen05.commit multicasts `empty` / `tfull` to both;
// the peer's epilogue warps arrive on the leader's `tempty`.  Rows, epilogue and outputs stay per CTA.
constexpr int kFlushBN = 160;
template <int MODE, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lse_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
              const __grid_constant__ CUtensorMap map_a2_hi, const __grid_constant__ CUtensorMap map_a2_lo,
              const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // a CTA of a pair stages half a B tile: 32 KB stages, six of them in the same 192 KB
  constexpr int BTILE = PAIR ? B_TILE_BYTES / 2 : B_TILE_BYTES;
  constexpr int STG = 2 * A_TILE_BYTES + 2 * BTILE;
  constexpr int NST = PAIR ? (STAGES * STAGE_BYTES) / STG : STAGES;
  static_assert(NST * STG <= STAGES * STAGE_BYTES && (2 * NST + 6) * 8 + 4 <= 256, "stage ring / barrier block out of its space");
  uint8_t* stage_base = smem;                                     // NST stages of 1024-aligned tiles
  float2* colinfo = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES);   // [2][BN] (norm_b, pot_b)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 2 * BN * 8);
  uint64_t* full = bars;                 // [NST]
  uint64_t* empty = bars + NST;          // [NST]
  uint64_t* tfull = bars + 2 * NST;      // [3]  (MODE 3 rotates three accumulator buffers, the others two)
  uint64_t* tempty = bars + 2 * NST + 3;      // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)blockIdx.x * BM;
  constexpr bool GEMM = (MODE >= 2);
  const int bn = GEMM ? p.bn : BN;                        // UMMA N of this launch
  const int n_btiles = (int)((p.nB + bn - 1) / bn);
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(n_btiles, t_begin + p.tiles_per_split);
  // MODE 0/1: this CTA owns one row tile and walks its column tiles.  MODE 2 (GEMM) is persistent: the CTA walks
  // row tiles blockIdx.x, blockIdx.x + gridDim.x, ... and all column tiles of each, as one continuous pipeline
  // (a 3-tile CTA spent a third of its life in prologue and the un-overlapped last epilogue).
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs of the pair)
  const bool leader = crank == 0;
  const int n_rtiles = (int)((p.nA + BM - 1) / BM);
  // persistent GEMM: CTAs (pairs) walk row tiles (256-row pair tiles) blockIdx.x, + gridDim.x, ...
  const int n_walk = PAIR ? (n_rtiles + 1) / 2 : n_rtiles;
  const int walkers = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x, walker = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int my_rtiles = GEMM ? (n_walk - walker + walkers - 1) / walkers : 1;
  const int n_tiles = GEMM ? my_rtiles * n_btiles : max(t_end - t_begin, 0);
  auto tile_i0 = [&](int t) -> int64_t {
    if (!GEMM) return i0;
    const int64_t w = (int64_t)walker + (int64_t)(t / n_btiles) * walkers;
    return PAIR ? (2 * w + crank) * BM : w * BM;
  };
  auto tile_j0 = [&](int t) -> int { return GEMM ? (t % n_btiles) * bn : (t_begin + t) * BN; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 3; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], PAIR ? 8 : 4); }   // epilogue warps (of both CTAs)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation is a warp-wide instruction; this warp also frees it
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync();          // both CTAs' barriers and TMEM exist before any cross-CTA signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = tile_j0(t);
        const int ti0 = (int)tile_i0(t);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = stage_base + s * STG;
          if (PAIR) {
            // own 128 rows of A, own half (bn / 2 rows) of B; all bytes of both CTAs count on the leader's barrier
            if (leader) mbar_expect_tx(&full[s], 4 * A_TILE_BYTES + 2 * bn * BK * 4);
            const int jh = j0 + (int)crank * (bn / 2);
            if (!GEMM || kb < p.kb_split) {
              tma_load_2d_pair(st, &map_a_hi, &full[s], kb * BK, ti0);
              tma_load_2d_pair(st + A_TILE_BYTES, &map_a_lo, &full[s], kb * BK, ti0);
            } else {
              tma_load_2d_pair(st, &map_a2_hi, &full[s], (kb - p.kb_split) * BK, ti0);
              tma_load_2d_pair(st + A_TILE_BYTES, &map_a2_lo, &full[s], (kb - p.kb_split) * BK, ti0);
            }
            tma_load_2d_pair(st + 2 * A_TILE_BYTES, &map_b_hi, &full[s], kb * BK, jh);
            tma_load_2d_pair(st + 2 * A_TILE_BYTES + BTILE, &map_b_lo, &full[s], kb * BK, jh);
            if (++s == NST) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[s], 2 * A_TILE_BYTES + 2 * bn * BK * 4);   // the B boxes are bn rows tall
          if (!GEMM || kb < p.kb_split) {
            tma_load_2d(st, &map_a_hi, &full[s], kb * BK, ti0);
            tma_load_2d(st + A_TILE_BYTES, &map_a_lo, &full[s], kb * BK, ti0);
          } else {
            tma_load_2d(st, &map_a2_hi, &full[s], (kb - p.kb_split) * BK, ti0);
            tma_load_2d(st + A_TILE_BYTES, &map_a2_lo, &full[s], (kb - p.kb_split) * BK, ti0);
          }
          tma_load_2d(st + 2 * A_TILE_BYTES, &map_b_hi, &full[s], kb * BK, j0);
          tma_load_2d(st + 2 * A_TILE_BYTES + BTILE, &map_b_lo, &full[s], kb * BK, j0);
          if (++s == NST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp walks the loop, one elected lane issues (elect_one) ==========
    if (!PAIR || leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM, bn);
      int s = 0; uint32_t ph = 0;
      uint32_t chain = 0;                                  // MODE 3: accumulation chains issued so far
      for (int t = 0; t < n_tiles; ++t) {
        int buf = t & 1;
        uint32_t tmem_d = tbase + (uint32_t)(buf * BN);
        if (MODE != 3) {
          const uint32_t use = (uint32_t)(t >> 1);         // how many times this buffer was used before
          mbar_wait(&tempty[buf], (use & 1) ^ 1);          // epilogue has drained the buffer
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (MODE == 3) {                                 // a fresh accumulator per k-block, three buffers rotate
            buf = (int)(chain % 3u);                       // (3 x 160 TMEM columns: the MMAs run two chains ahead
            tmem_d = tbase + (uint32_t)(buf * kFlushBN);       //  of the fold, which hides the commit -> wake-up hand-off)
            mbar_wait(&tempty[buf], ((chain / 3u) & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          mbar_wait(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(stage_base + s * STG);
          const uint64_t a_hi = make_smem_desc(st);
          const uint64_t a_lo = make_smem_desc(st + A_TILE_BYTES);
          const uint64_t b_hi = make_smem_desc(st + 2 * A_TILE_BYTES);
          const uint64_t b_lo = make_smem_desc(st + 2 * A_TILE_BYTES + BTILE);
          // the last k-block may be partly past d_pad (TMA zero-fills it): skip those k-steps
          const int k_steps = min(BK / UK, (p.d_pad - kb * BK) / UK);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              if (k >= k_steps) break;
              const uint64_t koff = (uint64_t)((k * UK * 4) >> 4);   // 32 B per k-step, in 16-byte units
              const uint32_t acc0 = MODE == 3 ? (uint32_t)(k != 0) : (uint32_t)((kb | k) != 0);
              if (PAIR) {
                umma_tf32_pair(tmem_d, a_hi + koff, b_hi + koff, idesc, acc0);
                umma_tf32_pair(tmem_d, a_hi + koff, b_lo + koff, idesc, 1);
                umma_tf32_pair(tmem_d, a_lo + koff, b_hi + koff, idesc, 1);
              } else {
                umma_tf32(tmem_d, a_hi + koff, b_hi + koff, idesc, acc0);
                umma_tf32(tmem_d, a_hi + koff, b_lo + koff, idesc, 1);
                umma_tf32(tmem_d, a_lo + koff, b_hi + koff, idesc, 1);
              }
            }
            if (PAIR) {
              umma_commit_pair(&empty[s]);                    // frees the stage in both CTAs
              if (MODE == 3) umma_commit_pair(&tfull[buf]);
            } else {
              umma_commit(&empty[s]);                         // smem stage reusable once these MMAs retire
              if (MODE == 3) umma_commit(&tfull[buf]);        // chain complete: the epilogue folds it
            }
          }
          __syncwarp();
          if (MODE == 3) ++chain;
          if (++s == NST) { s = 0; ph ^= 1; }
        }
        if (MODE != 3) {
          if (elect_one()) { if (PAIR) umma_commit_pair(&tfull[buf]); else umma_commit(&tfull[buf]); }   // accumulator complete
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: 4 warps, one TMEM lane (= row of A) per thread =====================
    const int ep_tid = threadIdx.x - 64;                     // 0..127
    const int quad = warp & 3;                               // TMEM lane quadrant this warp may touch
    const int row_in_tile = quad * 32 + lane;
    const int64_t row = i0 + row_in_tile;
    const float na = (!GEMM && row < p.nA) ? p.normA[row] : 0.f;
    const float fa = (MODE == 1 && row < p.nA) ? p.pot_a[row] : -CUDART_INF_F;
    const bool near_ok = (!GEMM) && row < p.nA && p.A_raw != nullptr && p.B_raw != nullptr;
    float run_m = -CUDART_INF_F, run_s = 0.f;
    double loss_acc = 0.0;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int64_t j0 = tile_j0(t);
      const int64_t trow = tile_i0(t) + row_in_tile;        // == row except in the persistent GEMM mode
      // stage (norm_b, pot_b) for this tile; buffer `buf` was last read two tiles ago
      float2* ci = colinfo + buf * BN;
      for (int c = ep_tid; c < BN; c += 128) {
        int64_t j = j0 + c;
        if (GEMM)
          ci[c] = make_float2(0.f, (j < p.nB && p.pot_in) ? p.pot_in[j] : 0.f);
        else
          ci[c] = (j < p.nB) ? make_float2(p.normB[j], p.pot_in[j]) : make_float2(0.f, -CUDART_INF_F);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (MODE == 3) {
        float accf[kFlushBN];
#pragma unroll
        for (int c = 0; c < kFlushBN; ++c) accf[c] = 0.f;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          const uint32_t chain = (uint32_t)t * (uint32_t)p.k_blocks + (uint32_t)kb;
          const int cb = (int)(chain % 3u);
          mbar_wait(&tfull[cb], (chain / 3u) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb * kFlushBN);
#pragma unroll
          for (int c0 = 0; c0 < kFlushBN; c0 += 64) {        // two loads in flight per wait
            if (c0 < bn) {                                   // CTA-uniform
              uint32_t r0[32], r1[32];
              const bool two = (c0 + 32 < kFlushBN) && (c0 + 32 < bn);
              tmem_ld32_issue(ta + (uint32_t)c0, r0);
              if (two) tmem_ld32_issue(ta + (uint32_t)(c0 + 32), r1);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) accf[c0 + c] += __uint_as_float(r0[c]);
              if (two) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c0 + 32 + c < kFlushBN) accf[c0 + 32 + c] += __uint_as_float(r1[c]);
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) { if (PAIR && !leader) arrive_on_leader(&tempty[cb]); else mbar_arrive(&tempty[cb]); }
        }
        if (trow < p.nA) {
#pragma unroll
          for (int c = 0; c < kFlushBN; c += 4) {
            const int64_t j = j0 + c;
            if (c < bn && j < p.nB) {                        // nB, n1, bn are multiples of 4 (checked on the host)
              const float4 v = make_float4(accf[c] + ci[c].y, accf[c + 1] + ci[c + 1].y, accf[c + 2] + ci[c + 2].y,
                                           accf[c + 3] + ci[c + 3].y);
              float* dst = (j < p.n1) ? p.out1 + trow * p.ld1 + j : p.out2 + trow * p.ld2 + (j - p.n1);
              *reinterpret_cast<float4*>(dst) = v;
            }
          }
        }
        continue;
      }
      mbar_wait(&tfull[buf], (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < bn; c0 += 32) {
        float dot[32];
        tmem_ld32(taddr + (uint32_t)c0, dot);
        if (MODE == 2) {
#pragma unroll
          for (int c = 0; c < 32; ++c) dot[c] += ci[c0 + c].y;
          if (trow < p.nA) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              const int64_t j = j0 + c0 + c;
              if (j < p.nB && c0 + c < bn) {      // nB, n1, bn are multiples of 4 (checked on the host)
                const float4 v = make_float4(dot[c], dot[c + 1], dot[c + 2], dot[c + 3]);
                float* dst = (j < p.n1) ? p.out1 + trow * p.ld1 + j : p.out2 + trow * p.ld2 + (j - p.n1);
                *reinterpret_cast<float4*>(dst) = v;
              }
            }
          }
          continue;
        }
        float z[32];
        float zmax = -CUDART_INF_F;
        float near_min = 1.f;        // < 0 iff some pair of this chunk needs the exact re-evaluation
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float4 info = *reinterpret_cast<const float4*>(&ci[c0 + c]);   // (nb0, pot0, nb1, pot1)
          float cst0, cst1;
          if (p.cost == EG_COST_COSINE) {
            const float den0 = fmaxf(na, 1e-8f) * fmaxf(info.x, 1e-8f), den1 = fmaxf(na, 1e-8f) * fmaxf(info.z, 1e-8f);
            near_min = fminf(near_min, fminf(fmaf(1.f - 2.f * kNearFrac, den0, -dot[c]),
                                             fmaf(1.f - 2.f * kNearFrac, den1, -dot[c + 1])));
            cst0 = 1.0f - __fdividef(dot[c], den0);
            cst1 = 1.0f - __fdividef(dot[c + 1], den1);
          } else {
            const float s0 = na + info.x, s1 = na + info.z;
            const float sq0 = fmaf(-2.0f, dot[c], s0), sq1 = fmaf(-2.0f, dot[c + 1], s1);
            near_min = fminf(near_min, fminf(fmaf(-kNearFrac, s0, sq0), fmaf(-kNearFrac, s1, sq1)));
            cst0 = (p.cost == EG_COST_L2) ? fast_sqrt(fmaxf(sq0, 0.f)) : fmaxf(sq0, 0.f);
            cst1 = (p.cost == EG_COST_L2) ? fast_sqrt(fmaxf(sq1, 0.f)) : fmaxf(sq1, 0.f);
          }
          z[c] = fmaf(-cst0, p.inv_reg, info.y);       // pot = -inf on padded columns -> z = -inf
          z[c + 1] = fmaf(-cst1, p.inv_reg, info.w);
        }
        // rare and warp-uniform: lanes that found close pairs in this chunk take turns; for each such pair the
        // WHOLE warp re-evaluates it from the fp32 rows (exact_pair_warp) and the owner patches its z
        unsigned owners = (!GEMM) ? __ballot_sync(0xffffffffu, near_ok && near_min < 0.f) : 0u;
        while (owners) {
          const int src = __ffs(owners) - 1;
          owners &= owners - 1;
          unsigned mine = 0;
          if (lane == src) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float2 inf2 = ci[c0 + c];
              bool hit;
              if (p.cost == EG_COST_COSINE)
                hit = dot[c] > (1.f - 2.f * kNearFrac) * fmaxf(na, 1e-8f) * fmaxf(inf2.x, 1e-8f);
              else
                hit = fmaf(-2.0f, dot[c], na + inf2.x) < kNearFrac * (na + inf2.x);
              if (hit && inf2.y > -CUDART_INF_F) mine |= 1u << c;
            }
          }
          unsigned cols = __shfl_sync(0xffffffffu, mine, src);
          const float* arow = p.A_raw + (i0 + quad * 32 + src) * p.d;
          while (cols) {
            const int c = __ffs(cols) - 1;
            cols &= cols - 1;
            const float ex = exact_pair_warp(arow, p.B_raw + (j0 + c0 + c) * p.d, p.d, p.cost == EG_COST_COSINE, lane);
            if (lane == src) {
              const float2 inf2 = ci[c0 + c];
              float cst;
              if (p.cost == EG_COST_COSINE) cst = 1.0f - __fdividef(ex, fmaxf(na, 1e-8f) * fmaxf(inf2.x, 1e-8f));
              else cst = (p.cost == EG_COST_L2) ? fast_sqrt(ex) : ex;
              const float znew = fmaf(-cst, p.inv_reg, inf2.y);
#pragma unroll
              for (int cc = 0; cc < 32; ++cc)
                if (cc == c) z[cc] = znew;
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) zmax = fmaxf(zmax, z[c]);
        if (MODE == 0) {
          if (zmax > run_m) { run_s *= __expf(run_m - zmax); run_m = zmax; }
          if (run_m > -CUDART_INF_F) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) acc += __expf(z[c] - run_m);
            run_s += acc;
          }
        } else {
          // z = g_j - cost*inv_reg  =>  cost = (g_j - z) / inv_reg ;  P = exp(f_i + z)
          float psum = 0.f, pc = 0.f;
          const float reg = 1.0f / p.inv_reg;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float pe = __expf(fa + z[c]);                     // 0 on padded rows/columns
            const float gj = ci[c0 + c].y;
            psum += pe;
            pc += (pe > 0.f) ? pe * (gj - z[c]) * reg : 0.f;
          }
          run_s += psum;
          loss_acc += (double)pc;
        }
      }
      // release the accumulator buffer: all tcgen05.ld of this warp have completed (wait::ld above)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) { if (PAIR && !leader) arrive_on_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]); }
    }
    if (MODE == 0) {
      if (row < p.nA) {
        p.part_m[(int64_t)blockIdx.y * p.nA + row] = run_m;
        p.part_s[(int64_t)blockIdx.y * p.nA + row] = run_s;
      }
    } else if (MODE == 1) {
      if (row < p.nA && p.row_sum) atomicAdd(&p.row_sum[row], run_s);
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && p.loss) atomicAdd(p.loss, loss_acc);
    }
  }
  // ---- teardown: everyone done with TMEM before the owner frees it ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) {
    cluster_sync();                  // no remote arrive, multicast commit or pair MMA may still target a CTA that left
    if (warp == 1)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
  } else if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// One launch path for every instantiation: opt-in shared memory once per device, 2-CTA clusters for PAIR.
template <int MODE, bool PAIR>
static int launch_tc(dim3 grid, const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mb_hi,
                     const CUtensorMap& mb_lo, const CUtensorMap& ma2_hi, const CUtensorMap& ma2_lo, const Params& p,
                     cudaStream_t s) {
  static PerDeviceOnce once;
  EG_SET_SMEM_ONCE(once, EG_CUDA(cudaFuncSetAttribute(lse_tc_kernel<MODE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      SMEM_BYTES)));
  if (!PAIR) {
    lse_tc_kernel<MODE, false><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, ma2_hi, ma2_lo, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    EG_CUDA(cudaLaunchKernelEx(&cfg, lse_tc_kernel<MODE, true>, ma_hi, ma_lo, mb_hi, mb_lo, ma2_hi, ma2_lo, p));
  }
  EG_LAUNCHED();
  return EG_OK;
}

// ---- host side -------------------------------------------------------------------------------

// [n, d_pad] fp32 row-major; box = BK elements x box_rows rows, swizzle = row span, OOB reads as zero.
static int make_map(CUtensorMap* map, const float* base, int64_t n, int d_pad, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return EG_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)d_pad * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EG_OK : EG_ERR_INVALID;
}

// Row tiles x column splits: choose the split count whose CTA total best fills whole waves of 148
// SMs, keeping at least 4 column tiles per CTA so the prologue stays amortised.
static void pick_grid(int64_t nA, int64_t nB, int* splits, int* tiles_per_split) {
  const int64_t row_tiles = ceil_div(nA, BM);
  const int64_t b_tiles = ceil_div(nB, BN);
  int64_t best_s = 1;
  double best_eff = -1.0;
  const int64_t max_s = std::min<int64_t>(b_tiles, 64);
  for (int64_t sp = 1; sp <= max_s; ++sp) {
    const int64_t tps = ceil_div(b_tiles, sp);
    if (sp > 1 && tps < 4) break;
    const int64_t real_s = ceil_div(b_tiles, tps);
    const int64_t ctas = row_tiles * real_s;
    const int64_t waves = ceil_div(ctas, (int64_t)kNumSMs);
    // useful tile-work over allocated tile-work, with a small penalty per extra split (partials + prologue)
    const double eff = (double)(row_tiles * b_tiles) / (double)(waves * kNumSMs * tps) - 0.002 * (double)real_s;
    if (eff > best_eff + 1e-9) { best_eff = eff; best_s = real_s; }
  }
  const int64_t tps = ceil_div(b_tiles, best_s);
  *tiles_per_split = (int)tps;
  *splits = (int)ceil_div(b_tiles, tps);
}

}  // namespace tc

__global__ void lse_combine_tc_kernel(const float* __restrict__ part_m, const float* __restrict__ part_s,
                                      int n_split, int64_t n, const float* __restrict__ logw,
                                      float* __restrict__ pot_out, float* __restrict__ lse_out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m = -CUDART_INF_F, s = 0.f;
  for (int k = 0; k < n_split; ++k) {
    float om = part_m[(int64_t)k * n + i], os = part_s[(int64_t)k * n + i];
    float mx = fmaxf(m, om);
    if (mx > -CUDART_INF_F) { s = s * __expf(m - mx) + os * __expf(om - mx); m = mx; }
  }
  float l = s > 0.f ? m + logf(s) : -CUDART_INF_F;
  if (lse_out) lse_out[i] = l;
  if (pot_out) pot_out[i] = (logw ? logw[i] : 0.f) - l;
}

size_t lse_fused_tc_workspace(int64_t nA, int64_t nB, int d) {
  (void)d;
  int splits, tps;
  tc::pick_grid(nA, nB, &splits, &tps);
  return 2 * align_up(sizeof(float) * (size_t)splits * (size_t)nA);
}

namespace tc {
struct Launch {
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, ma2_hi, ma2_lo;
  Params p;
  int splits;
};
static bool use_pair() {
  static const int env = [] { const char* e = getenv("EG_TC_PAIR"); return e ? atoi(e) : -1; }();   // measurement override
  return (env >= 0 ? env : g_tune_tc_pair) != 0;
}
static int prepare(Launch* L, int cost, int64_t nA, int64_t nB, int d, const float* normA, const float* normB,
                   float inv_reg, const float* pot_in, const float* A_hi, const float* A_lo, const float* B_hi,
                   const float* B_lo, const float* A_raw, const float* B_raw) {
  // operands come from eg_split_tf32: [n, d_pad] with d_pad = d rounded up to 8; the k-loop runs over whole
  // BK-element blocks and relies on TMA zero fill past d_pad
  const int d_pad = (d + 7) / 8 * 8;
  if (nA >= (1ll << 31) || nB >= (1ll << 31)) return EG_ERR_UNSUPPORTED;
  if (((uintptr_t)A_hi | (uintptr_t)A_lo | (uintptr_t)B_hi | (uintptr_t)B_lo) & 15) return EG_ERR_INVALID;
  int rc;
  if ((rc = make_map(&L->ma_hi, A_hi, nA, d_pad, BM))) return rc;
  if ((rc = make_map(&L->ma_lo, A_lo, nA, d_pad, BM))) return rc;
  const int b_box = use_pair() ? BN / 2 : BN;               // a CTA of a pair stages half of the B tile
  if ((rc = make_map(&L->mb_hi, B_hi, nB, d_pad, b_box))) return rc;
  if ((rc = make_map(&L->mb_lo, B_lo, nB, d_pad, b_box))) return rc;
  L->ma2_hi = L->ma_hi;
  L->ma2_lo = L->ma_lo;
  Params& p = L->p;
  p = Params{};
  p.nA = nA; p.nB = nB; p.k_blocks = (d_pad + BK - 1) / BK; p.d_pad = d_pad; p.cost = cost; p.inv_reg = inv_reg;
  p.normA = normA; p.normB = normB; p.pot_in = pot_in;
  p.A_raw = A_raw; p.B_raw = B_raw; p.d = d;
  pick_grid(nA, nB, &L->splits, &p.tiles_per_split);
  return EG_OK;
}
}  // namespace tc

int lse_fused_tc(int cost, int64_t nA, int64_t nB, int d, const float* normA, const float* normB, float inv_reg,
                 const float* pot_in, const float* logw, float* pot_out, float* lse_out, const float* A_hi,
                 const float* A_lo, const float* B_hi, const float* B_lo, const float* A_raw, const float* B_raw,
                 void* ws, size_t ws_bytes, cudaStream_t s) {
  using namespace tc;
  Launch L;
  int rc = prepare(&L, cost, nA, nB, d, normA, normB, inv_reg, pot_in, A_hi, A_lo, B_hi, B_lo, A_raw, B_raw);
  if (rc) return rc;
  size_t half = align_up(sizeof(float) * (size_t)L.splits * (size_t)nA);
  if (ws_bytes < 2 * half) return EG_ERR_WORKSPACE;
  L.p.part_m = reinterpret_cast<float*>(ws);
  L.p.part_s = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + half);
  // pairs: an even number of row tiles (a tile past nA reads zeros and writes nothing)
  dim3 grid((unsigned)(use_pair() ? 2 * ceil_div(nA, 2 * BM) : ceil_div(nA, BM)), (unsigned)L.splits);
  rc = use_pair() ? launch_tc<0, true>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s)
                  : launch_tc<0, false>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s);
  if (rc) return rc;
  lse_combine_tc_kernel<<<(unsigned)ceil_div(nA, 256), 256, 0, s>>>(L.p.part_m, L.p.part_s, L.splits, nA, logw,
                                                                    pot_out, lse_out);
  EG_LAUNCHED();
  return EG_OK;
}

// Plan statistics on the same tiles: *loss = sum P∘cost, row_sum[i] = sum_j P_ij (both pre-zeroed by the caller).
int plan_fused_tc(int cost, int64_t nA, int64_t nB, int d, const float* normA, const float* normB, float inv_reg,
                  const float* f, const float* g, double* loss, float* row_sum, const float* A_hi,
                  const float* A_lo, const float* B_hi, const float* B_lo, const float* A_raw, const float* B_raw,
                  cudaStream_t s) {
  using namespace tc;
  Launch L;
  int rc = prepare(&L, cost, nA, nB, d, normA, normB, inv_reg, g, A_hi, A_lo, B_hi, B_lo, A_raw, B_raw);
  if (rc) return rc;
  L.p.pot_a = f; L.p.loss = loss; L.p.row_sum = row_sum;
  dim3 grid((unsigned)(use_pair() ? 2 * ceil_div(nA, 2 * BM) : ceil_div(nA, BM)), (unsigned)L.splits);
  return use_pair() ? launch_tc<1, true>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s)
                    : launch_tc<1, false>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s);
}


// C[m, n] = [A1 | A2][m, k1+k2] · B[n, k1+k2]ᵀ + bias, fp32 in/out, 3xTF32 on tcgen05.  Operands are the hi/lo
// splits (eg_split_tf32) with every K extent padded to a multiple of 16 (= one k-block).
int gemm_nt_tc(const float* A1_hi, const float* A1_lo, int k1p, const float* A2_hi, const float* A2_lo, int k2p,
               int64_t m, const float* B_hi, const float* B_lo, int64_t n, const float* bias, float* out1, int64_t ld1,
               int64_t n1, float* out2, int64_t ld2, int flush, cudaStream_t s) {
  using namespace tc;
  if (k1p <= 0 || k1p % BK || k2p < 0 || k2p % BK || n % 4 || n1 % 4 || n1 > n || n1 <= 0) return EG_ERR_INVALID;
  if (n1 < n && !out2) return EG_ERR_INVALID;
  if (m >= (1ll << 31) || n >= (1ll << 31)) return EG_ERR_UNSUPPORTED;
  if (((uintptr_t)out1 | (uintptr_t)out2) & 15 || (ld1 % 4) || (out2 && ld2 % 4)) return EG_ERR_INVALID;
  Launch L;
  int rc;
  const int kp = k1p + k2p;
  if ((rc = make_map(&L.ma_hi, A1_hi, m, k1p, BM))) return rc;
  if ((rc = make_map(&L.ma_lo, A1_lo, m, k1p, BM))) return rc;
  if (k2p) {
    if ((rc = make_map(&L.ma2_hi, A2_hi, m, k2p, BM))) return rc;
    if ((rc = make_map(&L.ma2_lo, A2_lo, m, k2p, BM))) return rc;
  } else {
    L.ma2_hi = L.ma_hi;
    L.ma2_lo = L.ma_lo;
  }
  // column tile: as few tiles as BN = 256 allows, each just wide enough (multiple of 16): n = 300 -> 2 x 160, not 2 x 256
  // (flush: short accumulation chains folded in registers, MODE 3 — the tile is capped so a row of sums fits)
  const bool pair = use_pair();
  const int64_t bn_cap = flush ? kFlushBN : BN;
  const int64_t n_ct = ceil_div(n, bn_cap);
  const int64_t gran = pair ? 32 : 16;                      // UMMA N granularity: 16 at M = 128, 32 for the 256-row pair MMA
  const int bn = (int)std::min<int64_t>(bn_cap, ceil_div(ceil_div(n, n_ct), gran) * gran);
  if ((rc = make_map(&L.mb_hi, B_hi, n, kp, pair ? bn / 2 : bn))) return rc;
  if ((rc = make_map(&L.mb_lo, B_lo, n, kp, pair ? bn / 2 : bn))) return rc;
  Params& p = L.p;
  p = Params{};
  p.nA = m; p.nB = n; p.k_blocks = kp / BK; p.d_pad = kp; p.kb_split = k1p / BK;
  p.pot_in = bias; p.out1 = out1; p.out2 = out2; p.ld1 = ld1; p.ld2 = ld2; p.n1 = n1;
  p.bn = bn;
  p.tiles_per_split = (int)ceil_div(n, (int64_t)bn);
  // persistent over row tiles (256-row tiles for pairs)
  dim3 grid(pair ? 2u * (unsigned)std::min<int64_t>(ceil_div(m, 2 * BM), kNumSMs / 2)
                 : (unsigned)std::min<int64_t>(ceil_div(m, BM), kNumSMs), 1);
  if (flush)
    return pair ? launch_tc<3, true>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s)
                : launch_tc<3, false>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s);
  return pair ? launch_tc<2, true>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s)
              : launch_tc<2, false>(grid, L.ma_hi, L.ma_lo, L.mb_hi, L.mb_lo, L.ma2_hi, L.ma2_lo, L.p, s);
}

}  // namespace eg
