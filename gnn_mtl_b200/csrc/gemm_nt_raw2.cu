// 2-SM version of the raw-operand NT 3xTF32 GEMM (gemm_nt_raw.cu): a pair of CTAs on the two SMs of a TPC computes one
// 256 x bn output tile with tcgen05.mma.cta_group::2.
//
// Why: the single-SM kernels are bound by the shared-memory port — a 128 x 160 x 8 tf32 MMA is 80 cycles of math but
// fetches 9.2 KB of 4-byte operands from shared memory (72 cycles at 128 B/clk) next to the TMA fill of the same
// memory (profiles/README.md, round 2).  In a CTA pair every SM stages only HALF of the B tile (the tensor cores
// exchange the halves), A comes from each SM's own tensor memory (the raw-operand path: fp32 tile -> converter warps
// -> hi | lo pair in TMEM), so per MMA an SM reads 2.5 KB of B from shared memory instead of 9.2 KB of A and B.
//
// Protocol (rank 0 = leader: the only issuer of MMAs):
//   * A: each CTA loads and converts its own 128 rows — rings and barriers are local;
//   * B: each CTA loads its half (bn/2 rows) of the host-made hi / lo pair with the cta_group::2 TMA form, whose bytes
//     count on the LEADER's ready barrier; the leader's B producer posts the expected bytes of both halves;
//   * ready[s] (leader) = 4 + 4 converter-warp arrivals (the peer's arrive remotely) + that expect_tx arrival;
//   * free[s], tfull[b]: multicast tcgen05.commit.cta_group::2 -> both CTAs' barriers;
//   * tempty[b] (leader) = 4 + 4 epilogue-warp arrivals; every CTA drains its own 128 accumulator rows.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace eg {
namespace ntr2 {
using namespace ::eg::tc;

constexpr int BM = 128;
// Columns per tile.  MMAs that accumulate into the same tile issue no faster than one per ~121 cycles whatever N is up
// to 192 (tools/mma_rate.cu: dependent tf32 MMAs with A from tensor memory; 128 x 160 x 8 is 80 cycles of math), so what
// counts is the NUMBER of column tiles per row, not their width: equal tiles (n = 600 -> 4 x 160, n = 300 -> 2 x 160)
// measured faster than 192-wide tiles plus a narrow last one (0.425 / 0.225 ms against 0.440 / 0.235 ms at 200k rows).
constexpr int BNMAX = 160;                   // 2 accumulators x 160 + 6 A stages x 32 columns = 512 TMEM columns;
                                             // UMMA N of a 256-row pair MMA must be a multiple of 32
constexpr int BK = 16;                       // fp32 per k-block = one 64-byte swizzle row
constexpr int UK = 8;
constexpr int ROW_BYTES = BK * 4;
constexpr int RA = 10;                       // raw A stages in flight (10 x 8 KB)
constexpr int SB = 6;                        // operand stages the MMAs read: this CTA's half of the B pair in shared memory
constexpr int SA = SB;                       // (6 x 10 KB) + its A pair in tensor memory (6 x 32 columns), one barrier pair
constexpr int A_BYTES = BM * ROW_BYTES;      // 8 KB
constexpr int B_BYTES = (BNMAX / 2) * ROW_BYTES;   // 5 KB: this CTA's half of the B tile
constexpr int B_STAGE = 2 * B_BYTES;
constexpr int TMEM_A0 = 2 * BNMAX;           // first TMEM column of the A stages
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = RA * A_BYTES + SB * B_STAGE + 2 * BNMAX * 4 + 512 + 1024;

struct Params {
  int64_t m, n;
  int k1, k2;              // K extents of the two A operands (k2 = 0: one operand)
  int kb1, k_blocks;       // k-blocks of A1, total
  int bn;                  // columns per full tile (multiple of 32, <= BNMAX); the last tile of a row may be narrower
  const float* bias;       // [n] or null
  float* out1; float* out2;
  int64_t ld1, ld2, n1;
  const float* addend;     // [m, ld_add] added to the result (n1 == n only) or null
  int64_t ld_add;
  unsigned long long* dbg; // nullable: CTA 0 adds clock cycles [A producer wait, conv wait raw, conv wait TMEM stage, conv work,
                           //   issuer wait A, issuer wait B, issuer issue, k-blocks, issuer wait accumulator, B producer wait]
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  desc |= (uint64_t)1 << 16;
  desc |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  desc |= (uint64_t)1 << 46;
  desc |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return desc;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void split4(const float4& x, float4& h, float4& l) {
  const float xs[4] = {x.x, x.y, x.z, x.w};
  float hs[4], ls[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    // cvt.rna.tf32.f32 = round the magnitude to 10 mantissa bits, ties away from zero: add half an ulp of the kept
    // part to the bit pattern and clear the 13 dropped bits (two full-rate integer ops instead of a conversion-pipe one)
    hs[e] = __uint_as_float((__float_as_uint(xs[e]) + 0x1000u) & 0xffffe000u);
    const float rem = xs[e] - hs[e];         // exact in fp32
    ls[e] = __uint_as_float((__float_as_uint(rem) + 0x1000u) & 0xffffe000u);
  }
  h = make_float4(hs[0], hs[1], hs[2], hs[3]);
  l = make_float4(ls[0], ls[1], ls[2], ls[3]);
}
// D[tmem of both CTAs] (+)= A[tmem of both CTAs] * B[halves in both CTAs' smem]; issued by the leader only
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
                 "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                 "r"(__float_as_uint(v[15]))
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_nt_raw2_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const Params p) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;                                     // RA x 8 KB raw A stages
  uint8_t* b_base = smem + RA * A_BYTES;                      // SB B-pair stages (1024-aligned: RA * 8 KB in front)
  float* bias_s = reinterpret_cast<float*>(b_base + SB * B_STAGE);   // [2][BNMAX]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + 2 * BNMAX * 4);
  uint64_t* full_a = bars;                      // [RA]  TMA bytes of a raw A stage landed
  uint64_t* empty_a = full_a + RA;              // [RA]  4 converter warps done reading it
  uint64_t* full_b = empty_a + RA;              // [SB]  B pair landed (TMA) and A pair written to TMEM (4 converter warps)
  uint64_t* empty_b = full_b + SB;              // [SB]  MMAs that read it retired
  uint64_t* tfull = empty_b + SB;               // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg_on = (p.dbg != nullptr) && blockIdx.x == 0;
  const int bn = p.bn;
  const int n_btiles = (int)((p.n + bn - 1) / bn);
  const uint32_t crank = cluster_ctarank();               // 0 = leader (issues the MMAs), 1 = peer
  const bool leader = crank == 0;
  const int n_clusters = (int)gridDim.x / 2, cluster_id = (int)blockIdx.x / 2;
  const int n_rtiles = (int)((p.m + 2 * BM - 1) / (2 * BM));            // 256-row tiles of the pair
  const int my_rtiles = (n_rtiles - cluster_id + n_clusters - 1) / n_clusters;
  const int n_tiles = my_rtiles * n_btiles;
  // first row of THIS CTA's half of pair tile t
  auto tile_i0 = [&](int t) -> int64_t {
    return ((int64_t)cluster_id + (int64_t)(t / n_btiles) * n_clusters) * (2 * BM) + (int64_t)crank * BM;
  };
  auto tile_j0 = [&](int t) -> int { return (t % n_btiles) * bn; };
  // columns of tile t: the row's last tile covers what is left, rounded up to the pair MMA's granularity of 32
  auto tile_bn = [&](int t) -> int { return min(bn, (int)((p.n - tile_j0(t) + 31) / 32) * 32); };
  // live k-steps of k-block kb: the last block of an operand may reach past its K extent (TMA zero-fills it)
  auto k_steps_of = [&](int kb) -> int {
    const int valid = (kb < p.kb1) ? p.k1 - kb * BK : p.k2 - (kb - p.kb1) * BK;
    return min(BK / UK, (valid + UK - 1) / UK);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < RA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 4); }
    // one "ready" barrier per operand stage: 4 converter warps (A pair in TMEM) + the B producer's expect_tx arrival;
    // one "free" barrier: the commit of the MMAs that read the stage (waited on by the converters and the B producer)
    for (int s = 0; s < SB; ++s) { mbar_init(&full_b[s], 9); mbar_init(&empty_b[s], 1); }   // 4 + 4 converter warps + expect_tx
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }      // both CTAs' epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                                          // both CTAs' barriers and TMEM exist before any cross-CTA signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer of A: raw fp32 tiles =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      long long w = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int ti0 = (int)tile_i0(t);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          const long long c0 = clock64();
          mbar_wait(&empty_a[s], ph ^ 1);
          if (dbg_on) w += clock64() - c0;
          mbar_expect_tx(&full_a[s], A_BYTES);
          if (kb < p.kb1) tma_load_2d(a_base + s * A_BYTES, &map_a1, &full_a[s], kb * BK, ti0);
          else tma_load_2d(a_base + s * A_BYTES, &map_a2, &full_a[s], (kb - p.kb1) * BK, ti0);
          if (++s == RA) { s = 0; ph ^= 1; }
        }
      }
      if (dbg_on) atomicAdd(p.dbg + 0, (unsigned long long)w);
    }
  } else if (warp == 2) {
    // ===================== TMA producer of B: the host-made hi / lo pair =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      long long w = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = tile_j0(t);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          const long long c0 = clock64();
          mbar_wait(&empty_b[s], ph ^ 1);
          if (dbg_on) w += clock64() - c0;
          uint8_t* st = b_base + s * B_STAGE;
          // the leader's barrier counts the bytes of both CTAs' halves (hi + lo, bn / 2 rows each, per CTA)
          if (leader) mbar_expect_tx(&full_b[s], 2 * bn * ROW_BYTES);
          const int jh = j0 + (int)crank * (tile_bn(t) / 2);              // the box stays bn / 2 rows; a narrow tile ignores the rest
          tma_load_2d_pair(st, &map_b_hi, &full_b[s], kb * BK, jh);         // B's parts are padded to 16 columns each
          tma_load_2d_pair(st + B_BYTES, &map_b_lo, &full_b[s], kb * BK, jh);
          if (++s == SB) { s = 0; ph ^= 1; }
        }
      }
      if (dbg_on) atomicAdd(p.dbg + 9, (unsigned long long)w);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop with warp-uniform values and one elected lane issues (tc_common.cuh, elect_one).
    if (leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      int sa = 0, sb = 0; uint32_t phb = 0;
      long long w_b = 0, w_tmem = 0, t_issue = 0, n_kb = 0;
      unsigned long long g0 = 0; long long k0 = 0;
      if (dbg_on) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0)); k0 = clock64(); }
      for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        long long c0 = 0;
        if (dbg_on) c0 = clock64();
        mbar_wait(&tempty[buf], (((uint32_t)(t >> 1)) & 1) ^ 1);
        if (dbg_on) w_tmem += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tbase + (uint32_t)(buf * BNMAX);
        const uint32_t idesc = make_idesc(2 * BM, tile_bn(t));
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          long long c2 = 0, c3 = 0;
          if (dbg_on) c2 = clock64();
          mbar_wait(&full_b[sb], phb);
          if (dbg_on) c3 = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(b_base + sb * B_STAGE);
          const uint32_t a_hi = tbase + (uint32_t)(TMEM_A0 + 32 * sa);        // lane 0 of the A stage: 16 hi | 16 lo columns
          const uint32_t a_lo = a_hi + 16u;
          const uint64_t b_hi = make_smem_desc(st);
          const uint64_t b_lo = make_smem_desc(st + B_BYTES);
          const int k_steps = k_steps_of(kb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              if (k >= k_steps) break;
              const uint64_t koff = (uint64_t)((k * UK * 4) >> 4);
              umma_tf32_ts(tmem_d, a_hi + (uint32_t)(k * UK), b_hi + koff, idesc, (uint32_t)((kb | k) != 0));
              umma_tf32_ts(tmem_d, a_hi + (uint32_t)(k * UK), b_lo + koff, idesc, 1);
              umma_tf32_ts(tmem_d, a_lo + (uint32_t)(k * UK), b_hi + koff, idesc, 1);
            }
            umma_commit_pair(&empty_b[sb]);                  // frees the stage in both CTAs
          }
          __syncwarp();
          if (dbg_on) { w_b += c3 - c2; t_issue += clock64() - c3; ++n_kb; }
          if (++sa == SA) sa = 0;
          if (++sb == SB) { sb = 0; phb ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tfull[buf]);
        __syncwarp();
      }
      if (dbg_on && lane == 0) {
        atomicAdd(p.dbg + 5, (unsigned long long)w_b);
        atomicAdd(p.dbg + 6, (unsigned long long)t_issue); atomicAdd(p.dbg + 7, (unsigned long long)n_kb);
        atomicAdd(p.dbg + 8, (unsigned long long)w_tmem);
        unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        atomicAdd(p.dbg + 10, (unsigned long long)(clock64() - k0)); atomicAdd(p.dbg + 11, g1 - g0);
      }
    }
  } else if (warp >= 8) {
    // ===================== A converters: row r of the raw tile -> lane r of a TMEM stage (hi | lo) ================
    const int r = (warp - 8) * 32 + lane;                      // row of the tile = TMEM lane (warp % 4 = lane quadrant)
    const int swz = (r >> 1) & 3;                              // 64-byte swizzle: 16-byte chunk c of row r sits at c ^ swz
    int ra = 0, sa = 0; uint32_t phr = 0, phs = 0;
    long long w_raw = 0, w_spl = 0, t_work = 0;
    for (int t = 0; t < n_tiles; ++t) {
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        const long long c0 = clock64();
        mbar_wait(&full_a[ra], phr);
        const long long c1 = clock64();
        const uint32_t row = smem_u32(a_base + ra * A_BYTES + r * ROW_BYTES);
        float hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                       : "r"(row + (uint32_t)((c ^ swz) << 4)));
          float4 h, l;
          split4(x, h, l);
          hi[4 * c] = h.x; hi[4 * c + 1] = h.y; hi[4 * c + 2] = h.z; hi[4 * c + 3] = h.w;
          lo[4 * c] = l.x; lo[4 * c + 1] = l.y; lo[4 * c + 2] = l.z; lo[4 * c + 3] = l.w;
        }
        const long long c2 = clock64();
        mbar_wait(&empty_b[sa], phs ^ 1);
        const long long c3 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(TMEM_A0 + 32 * sa);
        tmem_st16(ta, hi);
        tmem_st16(ta + 16u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        // the raw stage is released behind the tensor-memory stores that consumed the loaded registers (see gemm_nt_raw.cu)
        if (lane == 0) {
          mbar_arrive(&empty_a[ra]);
          if (leader) mbar_arrive(&full_b[sa]); else arrive_on_leader(&full_b[sa]);
        }
        if (dbg_on) { w_raw += c1 - c0; w_spl += c3 - c2; t_work += (c2 - c1) + (clock64() - c3); }
        if (++ra == RA) { ra = 0; phr ^= 1; }
        if (++sa == SA) { sa = 0; phs ^= 1; }
      }
    }
    if (dbg_on && warp == 8 && lane == 0) {
      atomicAdd(p.dbg + 1, (unsigned long long)w_raw); atomicAdd(p.dbg + 2, (unsigned long long)w_spl);
      atomicAdd(p.dbg + 3, (unsigned long long)t_work);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: one TMEM lane (= row of A) per thread =====================
    const int ep_tid = threadIdx.x - 128;                      // 0..127
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int64_t j0 = tile_j0(t);
      const int64_t trow = tile_i0(t) + row_in_tile;
      float* bs = bias_s + buf * BNMAX;
      for (int c = ep_tid; c < BNMAX; c += 128) {
        const int64_t j = j0 + c;
        bs[c] = (j < p.n && p.bias) ? p.bias[j] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tfull[buf], (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BNMAX);
      const int bn_t = tile_bn(t);
#pragma unroll 1
      for (int c0 = 0; c0 < bn_t; c0 += 32) {
        float dot[32];
        tmem_ld32(taddr + (uint32_t)c0, dot);
        if (trow < p.m) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int64_t j = j0 + c0 + c;
            if (j < p.n) {                      // n, n1, bn are multiples of 4 (checked on the host)
              float4 v = make_float4(dot[c] + bs[c0 + c], dot[c + 1] + bs[c0 + c + 1], dot[c + 2] + bs[c0 + c + 2],
                                     dot[c + 3] + bs[c0 + c + 3]);
              if (p.addend) {
                const float4 a = *reinterpret_cast<const float4*>(p.addend + trow * p.ld_add + j);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
              }
              float* dst = (j < p.n1) ? p.out1 + trow * p.ld1 + j : p.out2 + trow * p.ld2 + (j - p.n1);
              *reinterpret_cast<float4*>(dst) = v;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&tempty[buf]); else arrive_on_leader(&tempty[buf]); }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                    // no remote arrive, multicast commit or pair MMA may still target a CTA that left
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// [rows, k] fp32 row-major with row stride ld; box = BK elements x box_rows rows, 64-byte swizzle, OOB reads as zero
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int k, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return EG_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EG_OK : EG_ERR_INVALID;
}

}  // namespace ntr2

int gemm_nt_raw2(const float* A1, int64_t lda1, int k1, const float* A2, int64_t lda2, int k2, int64_t m,
                const float* B_hi, const float* B_lo, int64_t ldb, int64_t n, const float* bias, const float* addend,
                int64_t ld_add, float* out1, int64_t ld1, int64_t n1, float* out2, int64_t ld2, cudaStream_t s) {
  using namespace ntr2;
  if (k1 <= 0 || k2 < 0 || n % 4 || n1 % 4 || n1 > n || n1 <= 0 || (n1 < n && !out2)) return EG_ERR_INVALID;
  if (lda1 % 4 || (k2 && lda2 % 4) || ldb % 4 || ld1 % 4 || (out2 && ld2 % 4)) return EG_ERR_INVALID;
  if (((uintptr_t)A1 | (uintptr_t)A2 | (uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)out1 | (uintptr_t)out2 | (uintptr_t)addend) & 15)
    return EG_ERR_INVALID;
  if (addend && (n1 != n || ld_add % 4)) return EG_ERR_INVALID;
  if (m >= (1ll << 31) || n >= (1ll << 31)) return EG_ERR_UNSUPPORTED;
  const int kb1 = (k1 + BK - 1) / BK, kb2 = (k2 + BK - 1) / BK;
  if ((int64_t)(kb1 + kb2) * BK > ldb) return EG_ERR_INVALID;         // B holds the parts back to back, each padded to 16
  const int64_t n_ct = ceil_div(n, (int64_t)BNMAX);
  const int bn = (int)std::min<int64_t>(BNMAX, ceil_div(ceil_div(n, n_ct), (int64_t)32) * 32);   // pair MMA: N % 32 == 0
  CUtensorMap ma1, ma2, mbh, mbl;
  int rc;
  if ((rc = make_map(&ma1, A1, m, k1, lda1, BM))) return rc;
  if (k2) { if ((rc = make_map(&ma2, A2, m, k2, lda2, BM))) return rc; }
  else ma2 = ma1;
  if ((rc = make_map(&mbh, B_hi, n, (kb1 + kb2) * BK, ldb, bn / 2))) return rc;    // one CTA's half of the tile
  if ((rc = make_map(&mbl, B_lo, n, (kb1 + kb2) * BK, ldb, bn / 2))) return rc;
  Params p{};
  p.m = m; p.n = n; p.k1 = k1; p.k2 = k2; p.kb1 = kb1; p.k_blocks = kb1 + kb2; p.bn = bn;
  p.bias = bias; p.out1 = out1; p.out2 = out2; p.ld1 = ld1; p.ld2 = ld2; p.n1 = n1;
  p.addend = addend; p.ld_add = ld_add;
  static unsigned long long* dbg_buf = nullptr;
  if (getenv("EG_GEMM_RAW_DEBUG")) {                 // measurement aid: per-role wait / work cycles of CTA 0
    if (!dbg_buf) EG_CUDA(cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long)));
    EG_CUDA(cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(unsigned long long), s));
    p.dbg = dbg_buf;
  }
  static PerDeviceOnce attr_once;
  EG_SET_SMEM_ONCE(attr_once, EG_CUDA(cudaFuncSetAttribute(gemm_nt_raw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)));
  const unsigned grid = 2u * (unsigned)std::min<int64_t>(ceil_div(m, 2 * BM), kNumSMs / 2);
  gemm_nt_raw2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ma1, ma2, mbh, mbl, p);
  EG_LAUNCHED();
  if (p.dbg) {
    unsigned long long h[16];
    EG_CUDA(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, s));
    EG_CUDA(cudaStreamSynchronize(s));
    const double nk = (double)std::max<unsigned long long>(h[7], 1);
    fprintf(stderr, "[eagraft] gemm_nt_raw2 CTA 0, cycles per k-block (%llu k-blocks, bn %d): A producer waits %.0f, B producer waits %.0f | "
                    "A converter: waits raw %.0f, waits TMEM stage %.0f, works %.0f | issuer: waits A %.0f, waits B %.0f, issues %.0f; "
                    "waits accumulator %.0f per tile\n",
            h[7], bn, h[0] / nk, h[9] / nk, h[1] / nk, h[2] / nk, h[3] / nk, h[4] / nk, h[5] / nk, h[6] / nk,
            (double)h[8] / std::max(1.0, nk / p.k_blocks));
    fprintf(stderr, "[eagraft] gemm_nt_raw2 CTA 0: %llu SM cycles in %llu ns of the issuing thread's loop = %.0f MHz\n", h[10], h[11],
            h[11] ? 1e3 * (double)h[10] / (double)h[11] : 0.0);
  }
  return EG_OK;
}

}  // namespace eg

