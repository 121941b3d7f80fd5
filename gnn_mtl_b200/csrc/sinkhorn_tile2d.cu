// Scaling-domain Sinkhorn sweeps with the kernel matrix TILED IN TWO DIMENSIONS over the chip.
//
// Continuation of the on-chip solve of utils/ot_loss.py:53-55 (v = b / Kᵀu, u = a / Kv) for the reference's
// batch (3000 x 3000, config.py:30) after the log-domain warm-up of sinkhorn_onchip_kernel.  The row-block
// kernel (sinkhorn_onchip_scaling_kernel) needs two grid-wide barriers per sweep, because every CTA needs all of v
// and every column needs all CTAs.  Here the grid is NC thread-block clusters of 8 CTAs:
//     cluster p  <->  a block of rows,      CTA q of the cluster  <->  a slice of columns,
// and CTA (p, q) keeps the tile Kt[rows p, cols q] = exp2(LU_i + LV_j - M_ij/reg) on chip for the whole solve
// (6 rows per warp in registers, the rest in shared memory).  Per sweep
//   C  column partials of the tile -> global scratch                  (the only traffic that leaves the SM)
//      ONE grid barrier
//      every CTA of slice q adds the NC partials of ITS columns (fixed order) -> v for its own columns: no
//      second barrier and no broadcast, because a CTA only ever needs v on its own column slice
//   U  row partials of the tile -> pushed into the 8 CTAs of the cluster through distributed shared memory,
//      ONE cluster barrier, every CTA adds the 8 partials (fixed order) -> u for the cluster's rows
// All sums run in a fixed order (deterministic; CTAs that share a vector compute it bit-identically, which is
// what lets them take the fold-into-Kt decisions without talking to each other).
// Stop rule: the marginal error of utils/ot_loss.py:64-66 is accumulated during the sweep that computes Kᵀu and
// decided one barrier later; the iterate it refers to is kept as a back-up and restored when the rule fires.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sinkhorn_common.cuh"

namespace eg {

constexpr int kT2Threads = 512;
constexpr int kT2Warps = kT2Threads / 32;
constexpr int kT2CS = 8;            // CTAs per cluster = column slices
constexpr int kT2QG = 3;            // float4 column groups per lane -> at most 96 groups (384 columns) per CTA
constexpr int kT2RRW = 6;           // rows per warp kept in registers
constexpr int kT2MaxRowsPerWarp = 16;
constexpr int kT2MaxClusters = 32;
constexpr int kT2RedStride = kT2QG * 32 * 4;                         // floats per warp in the column-partial staging
constexpr int kT2RowStride = kT2Warps * kT2MaxRowsPerWarp + kT2MaxRowsPerWarp;   // floats per peer in the row-partial inbox
// (compile-time strides: the 16 + 8 addresses of the two fixed-order reductions become immediates, not registers)
constexpr long long kT2WaitClocks = 40000000ll;     // ~20 ms of SM clocks: bound of every wait on another CTA

__device__ int g_dev_tile2d_absorbs = 0;

struct T2Params {
  const float* M; int64_t I; int J; int64_t ld; double inv2;     // inv2 = log2(e) / reg
  const float* a; const float* b;
  float* log_u; float* log_v;               // in: warm-up potentials (natural log); out: accepted potentials
  const PersistState* warm;                 // state of the warm-up launch (nullable)
  int start_iter, max_iter; double stop_thr;
  uint32_t* part;                           // [2][kT2CS][nc][t2_words(gs)] sign-tagged words (zeroed before the launch)
  PersistState* st;
  int nc;                                   // clusters = row blocks
  int rows_per_cluster;                     // <= 16 * nrw
  int nrw;                                  // rows per warp (<= 16)
  int gs;                                   // float4 groups per column slice (<= 96)
  float absorb_log2; int force_fallback;
};
// words one CTA publishes per sweep: its column partials + one control word (total marginal error of the last check)
__host__ __device__ inline int t2_words(int /*gs*/) { return kT2QG * 32 * 4 + 4; }   // compile-time: poll addresses are immediates

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_idx() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// Remote store that also signals: the 4 bytes land in the peer's shared memory and its mbarrier's transaction count
// drops by 4 — the receiver waits for "all bytes of this sweep are here" on its OWN barrier, so the row exchange needs
// neither barrier.cluster nor the release fence in front of it (measured: ERRBAR + arrive + wait = 0.7 us per sweep).
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void t2_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void t2_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool t2_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// ---- tensor memory as the second tier of the tile -------------------------------------------------------------
// The rows of the tile that do not fit in registers used to live in shared memory: 21 LDS.128 per thread and phase =
// 0.68 us of shared-memory wavefronts per phase at 128 B/clk.  Tensor memory is idle in this kernel, is private to
// the CTA, and streams into registers at ~400 B/clk/SM with all 16 warps reading (tools/tmem_bw.cu), so each warp
// keeps its slow rows in ITS OWN window of TMEM (lanes 32*(warp%4).., 128 columns at 128*(warp/4)): lane = thread,
// column = 12 floats (3 float4 groups) per row.  Only tcgen05.ld / .st touch it; no other warp ever reads the window.
constexpr int kT2TmemColsPerRow = kT2QG * 4;
static_assert(kT2QG == 3 && (kT2MaxRowsPerWarp - kT2RRW + 1) / 2 * 2 * kT2TmemColsPerRow <= 128, "slow rows of a warp must fit its 128-column TMEM window");
__device__ __forceinline__ void tmem_ld_row_pair(uint32_t taddr, float (&r)[2 * kT2TmemColsPerRow]) {
  uint32_t x[24];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[8]), "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]) : "r"(taddr + 8));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[16]), "=r"(x[17]), "=r"(x[18]), "=r"(x[19]), "=r"(x[20]), "=r"(x[21]), "=r"(x[22]), "=r"(x[23]) : "r"(taddr + 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 24; ++k) r[k] = __uint_as_float(x[k]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               ::"r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Every exchanged number is a sum of non-negative terms, so its sign bit is free to carry the hand-shake: the word of
// sweep s lives in slot (s & 1) and is published with sign bit t2_phase(s); the previous occupant of the slot (sweep
// s - 2, or the zero fill before the launch) carries the opposite bit, and the occupant of sweep s + 2 cannot appear
// before every reader of sweep s has published its own sweep s + 1.  A reader therefore polls the data itself: no
// barrier, no fence, no second round trip, 4 bytes per number.  (Relaxed gpu-scope accesses are served at the L2.)
__device__ __forceinline__ uint32_t t2_phase(int sweep_from_start) { return (((uint32_t)sweep_from_start >> 1) & 1u) ^ 1u; }
__device__ __forceinline__ void st_signed(uint32_t* p, uint32_t phase, float v) {
  const uint32_t w = (__float_as_uint(v) & 0x7fffffffu) | (phase << 31);
  asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(w) : "memory");
}
// Polls N words `kT2PartStride` apart until each carries `phase`; returns their sum in index order (fixed order
// -> every CTA that adds the same words gets the same bits).  N is always 16: the rows of clusters that do not exist
// (NC..15) are published as zeros by the last cluster, so the loop needs no per-word predicate.  All N loads of a round issue back to back from one
// base register (immediate offsets); a round that finds a stale word is simply repeated (a word, once valid, stays
// valid until its reader has moved on).  Gives up (sum of what arrived) when `*give_up` is set or after
// kT2WaitClocks, and reports that through *timed_out.
constexpr int kT2PartStride = kT2QG * 32 * 4 + 4;
template <int N>
__device__ __forceinline__ float t2_poll_sum(const uint32_t* src, uint32_t phase, volatile int* give_up, bool* timed_out) {
  uint32_t w[N];
  const uint32_t want = phase << 31;
  long long t0 = 0;
  while (true) {
#pragma unroll
    for (int k = 0; k < N; ++k)
      asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(w[k]) : "l"(src + (size_t)k * kT2PartStride) : "memory");
    uint32_t bad = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) bad |= w[k] ^ want;
    if ((bad >> 31) == 0u) break;
    if (t0 == 0) t0 = clock64();
    if (*give_up != 0 || clock64() - t0 > kT2WaitClocks) {      // never hang: stale words count as 0
      *timed_out = true;
#pragma unroll
      for (int k = 0; k < N; ++k) if (((w[k] ^ want) >> 31) != 0u) w[k] = 0u;
      break;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < N; ++k) s += fabsf(__uint_as_float(w[k]));     // published values are >= 0
  return s;
}

// 16 values per lane -> every lane ends with the warp-wide sum of value (lane & 15): 15 + 1 shuffles.
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
    const bool hi = (lane & w) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = hi ? v[k + w] : v[k];
      const float send = hi ? v[k] : v[k + w];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  float t = v[0];
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t;
}

// Packed fp32 FMA (FFMA2): two lanes of a float2 per issue slot; with b = (u, u) ptxas uses the scalar-broadcast form.
__device__ __forceinline__ void fma2(float2& acc, const float2 a, const float2 b) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(acc.x), "+f"(acc.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}
struct Acc4 { float2 lo, hi; };     // one float4 accumulator as two packed halves
__device__ __forceinline__ void fma4s(Acc4& acc, const float4& k, float u) {      // acc += k * u
  const float2 uu = make_float2(u, u);
  fma2(acc.lo, make_float2(k.x, k.y), uu);
  fma2(acc.hi, make_float2(k.z, k.w), uu);
}
__device__ __forceinline__ void fma4v(float2& acc, const float4& k, const float4& v) {   // acc += (k.xy*v.xy) + (k.zw*v.zw)
  fma2(acc, make_float2(k.x, k.y), make_float2(v.x, v.y));
  fma2(acc, make_float2(k.z, k.w), make_float2(v.z, v.w));
}
__device__ __forceinline__ void scale4(float4& k, float u, const float4& v) {
  k.x *= u * v.x; k.y *= u * v.y; k.z *= u * v.z; k.w *= u * v.w;
}

struct T2Smem {            // offsets in floats into dynamic shared memory
  int LU, LV, bak_f, bak_g, u, a, v, b, red_c, rowpart, ctrl_in, flags, mbar, red_e, kt;
  size_t bytes;
};
__host__ __device__ inline T2Smem t2_carve(int nrw, int gs) {
  T2Smem s;
  const int rb = kT2Warps * nrw + kT2MaxRowsPerWarp;   // padded row count (+ slack for register rows past nrw)
  const int nc4 = gs * 4;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) / 4 * 4; return o; };
  s.LU = take(2 * rb); s.LV = take(2 * nc4);           // doubles: folded potentials keep their low bits
  s.bak_f = take(2 * rb); s.bak_g = take(2 * nc4);     // doubles
  s.red_e = take(2 * kT2Warps);                        // doubles
  s.u = take(rb); s.a = take(rb);
  s.v = take(nc4); s.b = take(nc4);
  s.red_c = take(kT2Warps * kT2RedStride);
  s.rowpart = take(2 * kT2CS * kT2RowStride);
  s.ctrl_in = take(2 * kT2CS);                         // [2][kT2CS] marginal-error partials pushed by the cluster peers
  s.flags = take(8);
  s.mbar = take(4);                                    // two 8-byte mbarriers: row-exchange inbox of even / odd sweeps
  s.kt = take(0);
  (void)nrw;
  s.bytes = sizeof(float) * (size_t)off;          // the tile itself lives in registers + tensor memory
  return s;
}

// TIMING = true (EG_PERSIST_TIMING set): CTA 0 accumulates per-phase globaltimer deltas; the shipping instantiation has
// none of that code.
template <bool TIMING>
__global__ void __cluster_dims__(kT2CS, 1, 1) __launch_bounds__(kT2Threads, 1)
sinkhorn_tile2d_kernel(const T2Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw_t2[];
  float* sm = reinterpret_cast<float*>(smem_raw_t2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank(), p = (int)cluster_idx();
  const int NC = P.nc;
  PersistState* st = P.st;
  // the warm-up launch already met the stop rule: its potentials are the answer
  if (P.warm != nullptr && P.warm->sweeps < P.start_iter) return;
  if (cluster_nctarank() != (uint32_t)kT2CS) {          // launched without its cluster shape: the host redoes the solve
    if (tid == 0) atomicMax(&st->fallback, 2);
    return;
  }

  const int nrw = P.nrw, SRW = nrw > kT2RRW ? nrw - kT2RRW : 0;
  const int Rb = P.rows_per_cluster;
  const int64_t row_base = (int64_t)p * Rb;
  const int R = (int)max((int64_t)0, min((int64_t)Rb, P.I - row_base));       // live rows of this cluster
  const int ng = P.J / 4, Gs = P.gs, g_base = q * Gs;
  const int Gq = max(0, min(Gs, ng - g_base));                                  // live column groups of this slice
  const int nc4 = Gs * 4, ncols = Gq * 4, W = t2_words(Gs);
  const int rb_pad = kT2Warps * nrw + kT2MaxRowsPerWarp;
  const T2Smem L = t2_carve(nrw, Gs);
  double* LU_s = reinterpret_cast<double*>(sm + L.LU); double* LV_s = reinterpret_cast<double*>(sm + L.LV);
  double* bak_f = reinterpret_cast<double*>(sm + L.bak_f); double* bak_g = reinterpret_cast<double*>(sm + L.bak_g);
  double* red_e = reinterpret_cast<double*>(sm + L.red_e);
  float* u_s = sm + L.u; float* a_s = sm + L.a; float* v_s = sm + L.v; float* b_s = sm + L.b;
  float* red_c = sm + L.red_c;
  float* rowpart = sm + L.rowpart;                       // [2][kT2CS][kT2RowStride]
  float* ctrl_in = sm + L.ctrl_in;                       // [2][kT2CS]
  int* flag_s = reinterpret_cast<int*>(sm + L.flags);    // [0] fold v, [1] fold u, [2] bad sum, [3] break code, [4] gave up waiting
  float* ctrl_f = reinterpret_cast<float*>(flag_s + 5);  // [5] total error of the last check (to publish), [6] incoming total
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(flag_s + 7);     // [7] base address of the TMEM allocation
  const int SRP = (SRW + 1) / 2;                          // slow rows are stored and read two at a time
  const float kAbsorb = P.absorb_log2;

  // ---- hand-over: potentials (log2 units), marginals, and the tile of Kt ------------------------------------
  for (int r = tid; r < rb_pad; r += kT2Threads) {
    const bool live = r < R;
    u_s[r] = 1.0f;
    LU_s[r] = live ? (double)P.log_u[row_base + r] * 1.4426950408889634074 : 0.0;
    a_s[r] = live ? P.a[row_base + r] : 0.f;
  }
  for (int c = tid; c < nc4; c += kT2Threads) {
    const bool live = c < ncols;
    v_s[c] = live ? 1.0f : 0.f;
    LV_s[c] = live ? (double)P.log_v[(int64_t)g_base * 4 + c] * 1.4426950408889634074 : 0.0;
    b_s[c] = live ? P.b[(int64_t)g_base * 4 + c] : 0.f;
  }
  if (tid < 8) flag_s[tid] = 0;
  const uint32_t mbar_saddr = (uint32_t)__cvta_generic_to_shared(sm + L.mbar);
  if (tid == 0) {
    t2_mbar_init(mbar_saddr, 1);
    t2_mbar_init(mbar_saddr + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {                             // whole tensor memory of this SM (one CTA per SM, nothing else uses it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync_all();                          // every peer's inbox barriers exist before anyone can signal them
  const uint32_t tmem_base = *tmem_slot;
  // this warp's window: lanes 32*(warp%4).. (the only lanes a warp may address), 128 columns at 128*(warp/4)
  const uint32_t tmem_w = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);

  int lg[kT2QG], lgc[kT2QG];
  bool gok[kT2QG];
#pragma unroll
  for (int g = 0; g < kT2QG; ++g) {
    lg[g] = lane + 32 * g;
    gok[g] = lg[g] < Gq;
    lgc[g] = min(lg[g], max(Gs, 1) - 1);                  // clamped: loads stay inside the slice
  }
  const int wrow0 = warp * nrw;                           // first local row of this warp
  auto build = [&](int lr, int g) -> float4 {
    float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lr < R && gok[g]) {
      const float4 m = *reinterpret_cast<const float4*>(P.M + (row_base + lr) * P.ld + 4 * (int64_t)(g_base + lg[g]));
      const double lu = LU_s[lr], i2 = P.inv2;
      const double* lv = LV_s + 4 * lg[g];
      // exponent in fp64 (|M/reg| reaches hundreds: an fp32 fma would leave ~1e-5 relative in the entry), then ex2
      k.x = ex2f((float)(lu + lv[0] - (double)m.x * i2));
      k.y = ex2f((float)(lu + lv[1] - (double)m.y * i2));
      k.z = ex2f((float)(lu + lv[2] - (double)m.z * i2));
      k.w = ex2f((float)(lu + lv[3] - (double)m.w * i2));
    }
    return k;
  };
  float4 kreg[kT2RRW][kT2QG];
#pragma unroll
  for (int r = 0; r < kT2RRW; ++r)
#pragma unroll
    for (int g = 0; g < kT2QG; ++g) kreg[r][g] = (r < nrw) ? build(wrow0 + r, g) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sr = 0; sr < 2 * SRP; ++sr)           // (the odd row out of the last pair is stored as zeros)
#pragma unroll
    for (int g = 0; g < kT2QG; ++g)
      tmem_st4(tmem_w + (uint32_t)(sr * kT2TmemColsPerRow + 4 * g),
               sr < SRW ? build(wrow0 + kT2RRW + sr, g) : make_float4(0.f, 0.f, 0.f, 0.f));
  tmem_wait_st();
  __syncthreads();

  int cpt = P.start_iter, sweeps = P.start_iter;
  double err = 1.0;
  bool stop_hit = false;
  int pending_cpt = -1;                          // sweep whose marginal-error check is decided one exchange later
  const bool timer = TIMING && (blockIdx.x == 0 && tid == 0);
  const uint32_t rowpart_saddr = (uint32_t)__cvta_generic_to_shared(rowpart);
  const uint32_t ctrl_saddr = (uint32_t)__cvta_generic_to_shared(ctrl_in);

  // Column exchange of one sweep: publish my partials (sign-tagged), then wait for every cluster's partials of MY
  // columns.  Returns the cluster-order sum for thread `tid` (< nc4); thread nc4 fetches the control word of cluster 0.
  auto exchange = [&](int sweep, float mine, bool publish_partials) -> float {
    const uint32_t phase = t2_phase(sweep - P.start_iter);
    uint32_t* base = P.part + ((size_t)((sweep & 1) * kT2CS + q) * ((NC + 15) & ~15)) * W;
    const int NCP = (NC + 15) & ~15;                       // rows polled per column
    if (tid < nc4) {
      if (publish_partials) {
        st_signed(base + (size_t)p * W + tid, phase, mine);
        if (p == NC - 1)                                   // the rows of the clusters that do not exist read as +0
          for (int pp = NC; pp < NCP; ++pp) st_signed(base + (size_t)pp * W + tid, phase, 0.f);
      }
    } else if (tid == nc4) st_signed(base + (size_t)p * W + nc4, phase, ctrl_f[0]);
    float s = 0.f;
    if ((publish_partials && tid < nc4) || tid == nc4) {
      bool timed_out = false;
      volatile int* give_up = reinterpret_cast<volatile int*>(flag_s) + 4;
      if (tid == nc4) {                                    // the control word comes from cluster 0 only
        s = t2_poll_sum<1>(base + tid, phase, give_up, &timed_out);
      } else {
        for (int pp = 0; pp < NCP; pp += 16)
          s += t2_poll_sum<16>(base + (size_t)pp * W + tid, phase, give_up, &timed_out);
      }
      if (timed_out && reinterpret_cast<volatile int*>(flag_s)[4] == 0) {      // never hang: the host redoes the solve
        reinterpret_cast<volatile int*>(flag_s)[4] = 1;
        atomicMax(&st->fallback, 2);
      }
    }
    return s;
  };

  // sweeps until the next marginal-error check: the first sweep >= start_iter with (cpt - 1) % 10 == 0 and cpt >= 1
  int until_check = (P.start_iter <= 1) ? (1 - P.start_iter) : ((10 - (P.start_iter - 1) % 10) % 10);
  for (cpt = P.start_iter; cpt < P.max_iter; ++cpt) {
    const int par = cpt & 1;
    const bool check = (until_check == 0);         // sweeps 1, 11, 21, ... (utils/ot_loss.py:63: cpt % 10 == 0 after the increment)
    until_check = check ? 9 : until_check - 1;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    if (TIMING && timer) t0 = gtime();
    // this sweep's inbox: 8 peers x (16 warps x nrw rows) floats, + 8 error shares on a checking sweep
    if (tid == 0) t2_mbar_expect(mbar_saddr + 8u * (uint32_t)par, 4u * (uint32_t)(kT2CS * kT2Warps * nrw + (check ? kT2CS : 0)));
    // ---- C: column partials of my tile (thread: 3 column groups x the warp's rows), packed FMAs ----------------
    {
      Acc4 acc[kT2QG];
#pragma unroll
      for (int g = 0; g < kT2QG; ++g) { acc[g].lo = make_float2(0.f, 0.f); acc[g].hi = make_float2(0.f, 0.f); }
      float ur[kT2RRW];
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r) ur[r] = u_s[wrow0 + r];
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r)
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) fma4s(acc[g], kreg[r][g], ur[r]);
      const float* uw = u_s + wrow0 + kT2RRW;
#pragma unroll
      for (int sp = 0; sp < (kT2MaxRowsPerWarp - kT2RRW) / 2; ++sp) {
        if (sp < SRP) {                          // warp-uniform; unrolled so the TMEM addresses are immediates
          float k[2 * kT2TmemColsPerRow];
          tmem_ld_row_pair(tmem_w + (uint32_t)(sp * 2 * kT2TmemColsPerRow), k);
          const float ua = uw[2 * sp], ub = uw[2 * sp + 1];    // (finite even for the padding row; its entries are 0)
#pragma unroll
          for (int g = 0; g < kT2QG; ++g) {
            fma4s(acc[g], make_float4(k[4 * g], k[4 * g + 1], k[4 * g + 2], k[4 * g + 3]), ua);
            fma4s(acc[g], make_float4(k[12 + 4 * g], k[13 + 4 * g], k[14 + 4 * g], k[15 + 4 * g]), ub);
          }
        }
      }
#pragma unroll
      for (int g = 0; g < kT2QG; ++g)
        if (lg[g] < Gs)                          // (dead groups: their tile entries are zero, so is the sum)
          reinterpret_cast<float4*>(red_c)[warp * (kT2RedStride / 4) + lg[g]] =
              make_float4(acc[g].lo.x, acc[g].lo.y, acc[g].hi.x, acc[g].hi.y);
    }
    __syncthreads();
    float mine = 0.f;
    if (tid < nc4) {
#pragma unroll
      for (int w2 = 0; w2 < kT2Warps; ++w2) mine += red_c[w2 * kT2RedStride + tid];
    }
    if (TIMING && timer) t1 = gtime();
    // ---- all clusters' partials of MY columns -> v on my slice (every CTA of slice q computes the same bits) ----
    const float s = exchange(cpt, mine, true);
    if (TIMING && timer) t2 = gtime();
    if (check) {   // back-up of the iterate the check refers to (before sweep cpt touches it), as full potentials
      for (int r = tid; r < rb_pad; r += kT2Threads) bak_f[r] = LU_s[r] + log2((double)u_s[r]);
    }
    {
      double d2 = 0.0;
      if (tid < nc4) {
        const bool live = tid < ncols;
        const float v_old = v_s[tid];
        if (check) {
          bak_g[tid] = live ? LV_s[tid] + log2((double)v_old) : 0.0;
          if (live && p == 0) {
            const double d = (double)(v_old * s) - (double)b_s[tid];
            d2 = d * d;
          }
        }
        if (live) {
          const float vn = __fdividef(b_s[tid], s);       // every CTA of the slice runs the same instructions: same bits
          if (!(s > 0.f) || !(vn < CUDART_INF_F) || !(vn > 0.f)) flag_s[2] = 1;
          if (fabsf(lg2f(vn)) > kAbsorb) flag_s[0] = 1;
          v_s[tid] = vn;
        }
      } else if (tid == nc4) {
        ctrl_f[1] = s;                           // total marginal error of the previous sweep's check (from cluster 0)
      }
      if (check && p == 0) {                     // CTA-uniform: every lane takes part; fixed order -> deterministic
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        if (lane == 0) red_e[warp] = d2;
      }
    }
    __syncthreads();
    if (TIMING && timer) t3 = gtime();
    if (pending_cpt >= 0) {                      // the check issued one sweep ago: same word in every CTA
      err = sqrt((double)ctrl_f[1]);
      const bool stop = !(err > P.stop_thr);
      if (stop) { stop_hit = true; break; }
      pending_cpt = -1;
    }
    if (flag_s[2] && tid == 0) atomicMax(&st->fallback, 1);
    if (check) pending_cpt = cpt;
    float4 vq[kT2QG];
#pragma unroll
    for (int g = 0; g < kT2QG; ++g) vq[g] = gok[g] ? reinterpret_cast<const float4*>(v_s)[lg[g]] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (flag_s[0]) {
      // fold v into the tile and the column potentials (every CTA of this slice takes the same decision)
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r)
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) scale4(kreg[r][g], 1.0f, vq[g]);
      for (int sp = 0; sp < SRP; ++sp) {
        float k[2 * kT2TmemColsPerRow];
        const uint32_t ta = tmem_w + (uint32_t)(sp * 2 * kT2TmemColsPerRow);
        tmem_ld_row_pair(ta, k);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int g = 0; g < kT2QG; ++g) {
            float4 kk = make_float4(k[12 * h + 4 * g], k[12 * h + 4 * g + 1], k[12 * h + 4 * g + 2], k[12 * h + 4 * g + 3]);
            scale4(kk, 1.0f, vq[g]);             // vq is 0 on dead groups, and so are their entries
            tmem_st4(ta + (uint32_t)(12 * h + 4 * g), kk);
          }
      }
      tmem_wait_st();
      if (tid < ncols) { LV_s[tid] += log2((double)v_s[tid]); v_s[tid] = 1.0f; }
      if (tid == 0) { flag_s[0] = 0; if (blockIdx.x == 0) atomicAdd(&g_dev_tile2d_absorbs, 1); }
#pragma unroll
      for (int g = 0; g < kT2QG; ++g) if (gok[g]) vq[g] = make_float4(1.f, 1.f, 1.f, 1.f);
      __syncthreads();
    }
    // ---- U: row partials of my tile, exchanged inside the cluster --------------------------------------------
    {
      float dots[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) dots[r] = 0.f;
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r) {
        float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) fma4v(d2, kreg[r][g], vq[g]);
        dots[r] = d2.x + d2.y;
      }
#pragma unroll
      for (int sp = 0; sp < (kT2MaxRowsPerWarp - kT2RRW) / 2; ++sp) {
        if (sp < SRP) {                          // warp-uniform
          float k[2 * kT2TmemColsPerRow];
          tmem_ld_row_pair(tmem_w + (uint32_t)(sp * 2 * kT2TmemColsPerRow), k);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = 0; g < kT2QG; ++g)
              fma4v(d2, make_float4(k[12 * h + 4 * g], k[12 * h + 4 * g + 1], k[12 * h + 4 * g + 2], k[12 * h + 4 * g + 3]), vq[g]);
            dots[kT2RRW + 2 * sp + h] = d2.x + d2.y;
          }
        }
      }
      const float tot = warp_transpose_sum16(dots, lane);
      if (TIMING && timer) t4 = gtime();
      const int r = lane & 15;
      const uint32_t inbox_bar = mbar_saddr + 8u * (uint32_t)par;
      if (r < nrw) {
        const uint32_t local = rowpart_saddr + 4u * (uint32_t)((par * kT2CS + q) * kT2RowStride + wrow0 + r);
        const int q0 = (lane >> 4) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          st_async_f32(mapa_shared(local, (uint32_t)(q0 + k)), tot, mapa_shared(inbox_bar, (uint32_t)(q0 + k)));
      }
      if (check && tid < kT2CS) {               // my share of the marginal error rides along (zero outside cluster 0)
        double e = 0.0;
        if (p == 0) {
#pragma unroll
          for (int w2 = 0; w2 < kT2Warps; ++w2) e += red_e[w2];
        }
        st_async_f32(mapa_shared(ctrl_saddr + 4u * (uint32_t)(par * kT2CS + q), (uint32_t)tid), (float)e,
                     mapa_shared(inbox_bar, (uint32_t)tid));
      }
    }
    // ---- wait for the 8 peers' row partials (and error shares) of this sweep in my own inbox ---------------------
    if (tid < rb_pad || tid == kT2Threads - 1) {
      const uint32_t bar = mbar_saddr + 8u * (uint32_t)par;
      const uint32_t parity = ((uint32_t)(cpt - P.start_iter) >> 1) & 1u;
      if (!t2_mbar_try(bar, parity)) {
        const long long tw = clock64();
        while (!t2_mbar_try(bar, parity)) {
          if (reinterpret_cast<volatile int*>(flag_s)[4] != 0 || clock64() - tw > kT2WaitClocks) {   // never hang
            reinterpret_cast<volatile int*>(flag_s)[4] = 1;
            atomicMax(&st->fallback, 2);
            break;
          }
        }
      }
    }
    if (tid < rb_pad) {
      const float* rp = rowpart + (par * kT2CS) * kT2RowStride + tid;
      float rs = 0.f;
#pragma unroll
      for (int k = 0; k < kT2CS; ++k) rs += rp[k * kT2RowStride];
      if (tid < R) {
        const float un = __fdividef(a_s[tid], rs);
        if (!(rs > 0.f) || !(un < CUDART_INF_F) || !(un > 0.f)) flag_s[2] = 1;
        if (fabsf(lg2f(un)) > kAbsorb) flag_s[1] = 1;
        u_s[tid] = un;
      }
    } else if (check && tid == kT2Threads - 1) {
      float e = 0.f;
#pragma unroll
      for (int k = 0; k < kT2CS; ++k) e += ctrl_in[par * kT2CS + k];
      ctrl_f[0] = e;                             // cluster 0: the whole squared error, published with the next sweep
    }
    __syncthreads();
    if (flag_s[1]) {
      // fold u into the tile and the row potentials (every CTA of this cluster takes the same decision)
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r) {
        const float us = u_s[wrow0 + r];
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) { kreg[r][g].x *= us; kreg[r][g].y *= us; kreg[r][g].z *= us; kreg[r][g].w *= us; }
      }
      for (int sp = 0; sp < SRP; ++sp) {
        float k[2 * kT2TmemColsPerRow];
        const uint32_t ta = tmem_w + (uint32_t)(sp * 2 * kT2TmemColsPerRow);
        tmem_ld_row_pair(ta, k);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float us = (2 * sp + h < SRW) ? u_s[wrow0 + kT2RRW + 2 * sp + h] : 0.f;
#pragma unroll
          for (int g = 0; g < kT2QG; ++g)
            tmem_st4(ta + (uint32_t)(12 * h + 4 * g), make_float4(k[12 * h + 4 * g] * us, k[12 * h + 4 * g + 1] * us,
                                                                 k[12 * h + 4 * g + 2] * us, k[12 * h + 4 * g + 3] * us));
        }
      }
      tmem_wait_st();
      __syncthreads();
      if (tid < R) { LU_s[tid] += log2((double)u_s[tid]); u_s[tid] = 1.0f; }
      if (tid == 0) flag_s[1] = 0;
      __syncthreads();
    }
    sweeps = cpt + 1;
    if (TIMING && timer) {
      t5 = gtime();
      st->t_phase[0] += t1 - t0; st->t_phase[1] += t2 - t1; st->t_phase[2] += t3 - t2;
      st->t_phase[3] += t4 - t3; st->t_phase[4] += t5 - t4;
    }
  }
  // a check issued in the very last sweep: its total travels with one more (control-only) exchange
  if (cpt >= P.max_iter && pending_cpt >= 0) {
    const float e = exchange(P.max_iter, 0.f, false);
    if (tid == nc4) ctrl_f[1] = e;
    __syncthreads();
    err = sqrt((double)ctrl_f[1]);
    if (!(err > P.stop_thr)) stop_hit = true;
  }
  if (flag_s[2] && tid == 0) atomicMax(&st->fallback, 1);
  __syncthreads();
  cluster_sync_all();                            // nobody leaves while a peer could still be writing into its inbox
  // ---- outputs (natural-log potentials, composed in fp64) ------------------------------------------------------
  const double ln2 = 0.69314718055994530942;
  if (q == 0) {
    for (int r = tid; r < R; r += kT2Threads) {
      const double f = stop_hit ? bak_f[r] : LU_s[r] + log2((double)u_s[r]);
      P.log_u[row_base + r] = (float)(f * ln2);
    }
  }
  if (p == 0) {
    for (int c = tid; c < ncols; c += kT2Threads) {
      const double g = stop_hit ? bak_g[c] : LV_s[c] + log2((double)v_s[c]);
      P.log_v[(int64_t)g_base * 4 + c] = (float)(g * ln2);
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    st->sweeps = stop_hit ? pending_cpt : sweeps;
    st->final_buf = 0;
    st->err = err;
    if (P.force_fallback) atomicMax(&st->fallback, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// The communication skeleton of one sweep with no mat-vec work: sign-tagged column partials through global memory
// (publish, poll NC words per column), then the row-partial push into the peers' inboxes (st.async + their mbarrier)
// and the wait on the own inbox.  bench.py times it to state the latency floor of this design next to the measured
// sweep time.
__global__ void __cluster_dims__(kT2CS, 1, 1) __launch_bounds__(kT2Threads, 1)
sinkhorn_tile2d_sync_floor_kernel(PersistState* st, uint32_t* part, int NC, int Gs, int nrw, int iters) {
  extern __shared__ __align__(16) unsigned char smem_raw_t2[];
  float* sm = reinterpret_cast<float*>(smem_raw_t2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank(), p = (int)cluster_idx();
  if (cluster_nctarank() != (uint32_t)kT2CS) { if (tid == 0) atomicMax(&st->fallback, 2); return; }
  const int nc4 = Gs * 4, rb_pad = kT2Warps * nrw + kT2MaxRowsPerWarp, W = t2_words(Gs);
  const T2Smem L = t2_carve(nrw, Gs);
  float* u_s = sm + L.u; float* v_s = sm + L.v; float* red_c = sm + L.red_c; float* rowpart = sm + L.rowpart;
  const uint32_t rowpart_saddr = (uint32_t)__cvta_generic_to_shared(rowpart);
  const uint32_t mbar_saddr = (uint32_t)__cvta_generic_to_shared(sm + L.mbar);
  for (int i = tid; i < kT2Warps * kT2RedStride; i += kT2Threads) red_c[i] = 1.0f;
  for (int i = tid; i < rb_pad; i += kT2Threads) u_s[i] = 1.0f;
  if (tid == 0) {
    t2_mbar_init(mbar_saddr, 1);
    t2_mbar_init(mbar_saddr + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    const uint32_t phase = t2_phase(it);
    if (tid == 0) t2_mbar_expect(mbar_saddr + 8u * (uint32_t)par, 4u * (uint32_t)(kT2CS * kT2Warps * nrw));
    __syncthreads();
    uint32_t* base = part + ((size_t)(par * kT2CS + q) * ((NC + 15) & ~15)) * W;
    if (tid < nc4) {
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kT2Warps; ++w2) s += red_c[w2 * kT2RedStride + tid];
      st_signed(base + (size_t)p * W + tid, phase, s * u_s[0]);
      const int NCP = (NC + 15) & ~15;
      if (p == NC - 1)
        for (int pp = NC; pp < NCP; ++pp) st_signed(base + (size_t)pp * W + tid, phase, 0.f);
      float tot = 0.f;
      bool timed_out = false;
      int never = 0;
      for (int pp = 0; pp < NCP; pp += 16)
        tot += t2_poll_sum<16>(base + (size_t)pp * W + tid, phase, &never, &timed_out);
      if (timed_out) atomicMax(&st->fallback, 2);
      v_s[tid] = 1.0f / tot;
    }
    __syncthreads();
    const float tot = v_s[lane];
    const int r = lane & 15;
    const uint32_t inbox_bar = mbar_saddr + 8u * (uint32_t)par;
    if (r < nrw) {
      const uint32_t local = rowpart_saddr + 4u * (uint32_t)((par * kT2CS + q) * kT2RowStride + warp * nrw + r);
      const int q0 = (lane >> 4) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        st_async_f32(mapa_shared(local, (uint32_t)(q0 + k)), tot, mapa_shared(inbox_bar, (uint32_t)(q0 + k)));
    }
    if (tid < kT2Warps * nrw) {
      const uint32_t parity = ((uint32_t)it >> 1) & 1u;
      const long long tw = clock64();
      while (!t2_mbar_try(inbox_bar, parity))
        if (clock64() - tw > kT2WaitClocks) { atomicMax(&st->fallback, 2); break; }
      const float* rp = rowpart + (par * kT2CS) * kT2RowStride + tid;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kT2CS; ++k) s += rp[k * kT2RowStride];
      u_s[tid] = 1.0f / s;
    }
  }
  __syncthreads();
  cluster_sync_all();
}

int sinkhorn_tile2d_absorbs_read() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_dev_tile2d_absorbs, sizeof(int));
  return v;
}

// Co-residency probe: the same launch shape (cluster of 8, 512 threads, the solver's dynamic shared memory) doing
// ONE grid barrier with a 5 ms budget.  If some cluster cannot be resident together with the others the barrier
// times out and *ok is cleared — the occupancy query is a necessary condition, this is the sufficient one.
__global__ void __cluster_dims__(kT2CS, 1, 1) __launch_bounds__(kT2Threads, 1)
t2_probe_kernel(unsigned int* counter, int* ok, unsigned int nblocks) {
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    const unsigned long long t0 = gtime();
    for (;;) {
      unsigned int seen;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen >= nblocks) break;
      if (gtime() - t0 > 5000000ull) { atomicExch(ok, 0); break; }
    }
  }
  __syncthreads();
  cluster_sync_all();
}

// Launch geometry: the fewest rows per warp for which the co-resident clusters cover all rows.
struct T2Geometry { int nc, nrw, gs, rows_per_cluster; size_t smem; };

// Clusters of 8 CTAs (one CTA per SM at the largest shared-memory request) that were SEEN running together on this
// device: starts from the occupancy query and walks down until the probe's grid barrier completes.  One-time host
// synchronisation per device; 0 = unknown, -1 = clusters unusable.
static std::atomic<int> g_t2_verified_clusters[64];
static int t2_verified_clusters(PersistState* scratch, cudaStream_t s, int* out) {
  int dev = 0, max_smem = 0;
  EG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { *out = 0; return EG_OK; }
  int cached = g_t2_verified_clusters[dev].load(std::memory_order_acquire);
  if (cached != 0) { *out = std::max(cached, 0); return EG_OK; }
  EG_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  auto kern = t2_probe_kernel;
  const size_t smem = (size_t)max_smem - 1024;
  EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchAttribute attrs[2];
  // The cluster shape is compiled into the kernels (__cluster_dims__), so a launch carries only the cooperative
  // attribute; the cluster attribute is passed to the occupancy query alone.  (A profiler that re-issues the launch
  // through the plain cooperative entry point then still gets clusters of 8.)
  attrs[0].id = cudaLaunchAttributeCooperative;
  attrs[0].val.cooperative = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = kT2CS; attrs[1].val.clusterDim.y = 1; attrs[1].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kT2Threads);
  cfg.stream = s;
  cfg.attrs = attrs;
  cfg.dynamicSmemBytes = smem;
  cfg.gridDim = dim3(kT2CS * kT2MaxClusters);
  int n = 0;
  cfg.numAttrs = 2;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  }
  n = std::min(n, kT2MaxClusters);
  int verified = -1;
  for (; n >= 4; --n) {
    EG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(PersistState), s));
    int one = 1;
    EG_CUDA(cudaMemcpyAsync(&scratch->sweeps, &one, sizeof(int), cudaMemcpyHostToDevice, s));
    cfg.gridDim = dim3((unsigned)(kT2CS * n));
    cfg.numAttrs = 1;
    unsigned int nblocks = (unsigned)(kT2CS * n);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, &scratch->barrier, &scratch->sweeps, nblocks);
    if (e != cudaSuccess) { cudaGetLastError(); continue; }          // cooperative launch refused: too many clusters
    int ok = 0;
    EG_CUDA(cudaMemcpyAsync(&ok, &scratch->sweeps, sizeof(int), cudaMemcpyDeviceToHost, s));
    EG_CUDA(cudaStreamSynchronize(s));
    if (getenv("EG_PERSIST_TIMING")) fprintf(stderr, "[eagraft] cluster probe: %d clusters of %d -> %s\n", n, kT2CS, ok ? "co-resident" : "NOT co-resident");
    if (ok) { verified = n; break; }
  }
  g_t2_verified_clusters[dev].store(verified, std::memory_order_release);
  *out = std::max(verified, 0);
  return EG_OK;
}

template <typename Kern>
static int t2_geometry(Kern kern, int64_t I, int64_t J, cudaStream_t s, PersistState* scratch, T2Geometry* g,
                       cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attrs) {
  g->nc = 0;
  int dev = 0, max_smem = 0, verified = 0;
  EG_CUDA(cudaGetDevice(&dev));
  EG_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int vrc = t2_verified_clusters(scratch, s, &verified);
  if (vrc) return vrc;
  if (verified < 4) return EG_OK;
  g->gs = (int)ceil_div(J / 4, (int64_t)kT2CS);
  // The cluster shape is compiled into the kernels (__cluster_dims__), so a launch carries only the cooperative
  // attribute; the cluster attribute is passed to the occupancy query alone.  (A profiler that re-issues the launch
  // through the plain cooperative entry point then still gets clusters of 8.)
  attrs[0].id = cudaLaunchAttributeCooperative;
  attrs[0].val.cooperative = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = kT2CS; attrs[1].val.clusterDim.y = 1; attrs[1].val.clusterDim.z = 1;
  *cfg = cudaLaunchConfig_t{};
  cfg->blockDim = dim3(kT2Threads);
  cfg->stream = s;
  cfg->attrs = attrs;
  for (int try_nrw = 1; try_nrw <= kT2MaxRowsPerWarp; ++try_nrw) {
    const T2Smem L = t2_carve(try_nrw, g->gs);
    if (L.bytes > (size_t)max_smem) break;
    EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
    cfg->gridDim = dim3(kT2CS * kT2MaxClusters);
    cfg->dynamicSmemBytes = L.bytes;
    int max_clusters = 0;
    cfg->numAttrs = 2;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, cfg) != cudaSuccess) {
      cudaGetLastError();
      cfg->numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, cfg) != cudaSuccess) { cudaGetLastError(); return EG_OK; }
    }
    max_clusters = std::min(std::min(max_clusters, kT2MaxClusters), verified);
    if (max_clusters < 4) continue;
    if ((int64_t)max_clusters * kT2Warps * try_nrw >= I) {
      g->nrw = try_nrw;
      g->nc = (int)std::min<int64_t>(max_clusters, ceil_div(I, (int64_t)kT2Warps * try_nrw));
      g->smem = L.bytes;
      break;
    }
  }
  if (g->nc == 0) return EG_OK;
  g->rows_per_cluster = (int)ceil_div(I, (int64_t)g->nc);
  cfg->gridDim = dim3((unsigned)(kT2CS * g->nc));
  cfg->dynamicSmemBytes = g->smem;
  cfg->numAttrs = 1;
  return EG_OK;
}

int sinkhorn_tile2d_sync_floor_launch(int64_t I, int64_t J, int iters, void* part_, size_t part_bytes,
                                      PersistState* st, cudaStream_t s, bool* launched) {
  uint32_t* part = reinterpret_cast<uint32_t*>(part_);
  *launched = false;
  if (J % 4 != 0 || J > 4 * kT2CS * 32 * kT2QG || iters <= 0) return EG_OK;
  T2Geometry g;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  auto kern = sinkhorn_tile2d_sync_floor_kernel;
  int rc = t2_geometry(kern, I, J, s, st, &g, &cfg, attrs);
  if (rc) return rc;
  const size_t words = (size_t)2 * kT2CS * ((g.nc + 15) & ~15) * t2_words(g.gs);
  if (g.nc == 0 || words * 4 > part_bytes) return EG_OK;
  EG_CUDA(cudaMemsetAsync(st, 0, sizeof(PersistState), s));
  EG_CUDA(cudaMemsetAsync(part, 0, words * 4, s));               // sign bit 0 everywhere = "not yet published"
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, st, part, g.nc, g.gs, g.nrw, iters);
  if (e != cudaSuccess) { cudaGetLastError(); return EG_OK; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  *launched = true;
  return EG_OK;
}

int sinkhorn_tile2d_launch(const float* M, int64_t I, int64_t J, int64_t ld, double inv_reg, const float* a,
                           const float* b, float* log_u, float* log_v, const PersistState* warm, int start_iter,
                           int max_iter, double stop_thr, void* part_, size_t part_bytes, PersistState* st,
                           float absorb_log2, int force_fallback, cudaStream_t s, bool* launched) {
  uint32_t* part = reinterpret_cast<uint32_t*>(part_);
  *launched = false;
  if (J % 4 != 0 || (reinterpret_cast<uintptr_t>(M) & 15) || (ld % 4) != 0 || J > 4 * kT2CS * 32 * kT2QG) return EG_OK;
  T2Geometry g;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  const bool timing = getenv("EG_PERSIST_TIMING") != nullptr;
  auto kern = timing ? sinkhorn_tile2d_kernel<true> : sinkhorn_tile2d_kernel<false>;
  int rc = t2_geometry(kern, I, J, s, st, &g, &cfg, attrs);
  if (rc) return rc;
  const size_t words = (size_t)2 * kT2CS * ((g.nc + 15) & ~15) * t2_words(g.gs);
  if (g.nc == 0 || words * 4 > part_bytes) return EG_OK;
  EG_CUDA(cudaMemsetAsync(st, 0, sizeof(PersistState), s));      // the probe may have used it
  EG_CUDA(cudaMemsetAsync(part, 0, words * 4, s));               // sign bit 0 everywhere = "not yet published"
  T2Params P;
  P.M = M; P.I = I; P.J = (int)J; P.ld = ld; P.inv2 = inv_reg * 1.4426950408889634074;
  P.a = a; P.b = b; P.log_u = log_u; P.log_v = log_v;
  P.warm = warm; P.start_iter = start_iter; P.max_iter = max_iter; P.stop_thr = stop_thr; P.part = part; P.st = st;
  P.nc = g.nc; P.rows_per_cluster = g.rows_per_cluster; P.nrw = g.nrw; P.gs = g.gs; P.absorb_log2 = absorb_log2;
  P.force_fallback = force_fallback;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) {          // cooperative + cluster launch refused: the caller takes the row-block kernel
    cudaGetLastError();
    if (getenv("EG_PERSIST_TIMING")) fprintf(stderr, "[eagraft] tile2d launch refused: %s\n", cudaGetErrorString(e));
    return EG_OK;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (getenv("EG_PERSIST_TIMING"))
    fprintf(stderr, "[eagraft] tile2d sinkhorn: %d clusters x %d CTAs, %d rows/cluster, %d rows/warp, %d groups/slice, %zu B smem\n",
            g.nc, kT2CS, g.rows_per_cluster, g.nrw, g.gs, g.smem);
  *launched = true;
  return EG_OK;
}

}  // namespace eg
