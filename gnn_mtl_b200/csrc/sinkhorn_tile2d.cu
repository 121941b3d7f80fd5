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

__device__ int g_dev_tile2d_absorbs = 0;

struct T2Params {
  const float* M; int64_t I; int J; int64_t ld; double inv2;     // inv2 = log2(e) / reg
  const float* a; const float* b;
  float* log_u; float* log_v;               // in: warm-up potentials (natural log); out: accepted potentials
  const PersistState* warm;                 // state of the warm-up launch (nullable)
  int start_iter, max_iter; double stop_thr;
  float* part;                              // [2][kT2CS][nc][gs * 4] column partials
  PersistState* st;
  int nc;                                   // clusters = row blocks
  int rows_per_cluster;                     // <= 16 * nrw
  int nrw;                                  // rows per warp (<= 16)
  int gs;                                   // float4 groups per column slice (<= 96)
  float absorb_log2; int force_fallback;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_relaxed_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier on a monotonically increasing counter (all CTAs co-resident: verified once per device by
// t2_probe_kernel, and the launch is cooperative).  The wait is bounded: after 20 ms without progress (a sweep
// takes microseconds) the CTA raises the sticky `fallback = 2`, every waiter sees it and leaves its sweep loop at
// the next check — the solve is then redone by the log-domain kernel instead of hanging the device.
__device__ __forceinline__ void t2_grid_barrier(PersistState* st, unsigned int& target, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&st->barrier) : "memory");
    unsigned int seen;
    unsigned int spins = 0;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&st->barrier) : "memory");
      if ((int)(seen - target) >= 0) break;
      if ((++spins & 63u) == 0) {
        if (ld_relaxed_s32(&st->fallback) >= 2) break;   // aborted solve: nobody waits any more
        if (clock64() - t0 > 40000000ll) {               // ~20 ms of SM clocks
          if (atomicMax(&st->fallback, 2) < 2) {         // first to give up: leave a trace for EG_PERSIST_TIMING
            st->t_phase[6] = ((unsigned long long)blockIdx.x << 32) | seen;
            st->t_phase[7] = target;
          }
          atomicMax(&st->flag_code, 0x7fffffff);       // "raised before every sweep": acted on at once
          break;
        }
      }
    }
  }
  __syncthreads();
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

__device__ __forceinline__ float dot4(const float4& k, const float4& v) {
  return (k.x * v.x + k.y * v.y) + (k.z * v.z + k.w * v.w);
}
__device__ __forceinline__ void fma4(float4& acc, const float4& k, float u) {
  acc.x = fmaf(k.x, u, acc.x); acc.y = fmaf(k.y, u, acc.y); acc.z = fmaf(k.z, u, acc.z); acc.w = fmaf(k.w, u, acc.w);
}
__device__ __forceinline__ void scale4(float4& k, float u, const float4& v) {
  k.x *= u * v.x; k.y *= u * v.y; k.z *= u * v.z; k.w *= u * v.w;
}

struct T2Smem {            // offsets in floats into dynamic shared memory
  int u, LU, a, v, LV, b, red_c, rowpart, flags, bak_f, bak_g, kt;
  size_t bytes;
};
__host__ __device__ inline T2Smem t2_carve(int nrw, int gs) {
  T2Smem s;
  const int rb = kT2Warps * nrw + kT2MaxRowsPerWarp;   // padded row count (+ slack for register rows past nrw)
  const int nc4 = gs * 4;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) / 4 * 4; return o; };
  s.u = take(rb); s.LU = take(rb); s.a = take(rb);
  s.v = take(nc4); s.LV = take(nc4); s.b = take(nc4);
  s.red_c = take(kT2Warps * nc4);
  s.rowpart = take(2 * kT2CS * rb);
  s.flags = take(8);
  s.bak_f = take(2 * rb);       // doubles
  s.bak_g = take(2 * nc4);      // doubles
  s.kt = take(0);
  const int srw = nrw > kT2RRW ? nrw - kT2RRW : 0;
  s.bytes = sizeof(float) * (size_t)off + sizeof(float4) * (size_t)kT2Warps * srw * gs;
  return s;
}

__global__ void __launch_bounds__(kT2Threads, 1)
sinkhorn_tile2d_kernel(const T2Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw_t2[];
  float* sm = reinterpret_cast<float*>(smem_raw_t2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank(), p = (int)cluster_idx();
  const int NC = P.nc;
  const unsigned int nb = (unsigned int)(NC * kT2CS);
  PersistState* st = P.st;
  // the warm-up launch already met the stop rule: its potentials are the answer
  if (P.warm != nullptr && P.warm->sweeps < P.start_iter) return;

  const int nrw = P.nrw, SRW = nrw > kT2RRW ? nrw - kT2RRW : 0;
  const int Rb = P.rows_per_cluster;
  const int64_t row_base = (int64_t)p * Rb;
  const int R = (int)max((int64_t)0, min((int64_t)Rb, P.I - row_base));       // live rows of this cluster
  const int ng = P.J / 4, Gs = P.gs, g_base = q * Gs;
  const int Gq = max(0, min(Gs, ng - g_base));                                  // live column groups of this slice
  const int nc4 = Gs * 4, ncols = Gq * 4;
  const int rb_pad = kT2Warps * nrw + kT2MaxRowsPerWarp;
  const T2Smem L = t2_carve(nrw, Gs);
  float* u_s = sm + L.u; float* LU_s = sm + L.LU; float* a_s = sm + L.a;
  float* v_s = sm + L.v; float* LV_s = sm + L.LV; float* b_s = sm + L.b;
  float* red_c = sm + L.red_c;
  float* rowpart = sm + L.rowpart;                       // [2][kT2CS][rb_pad]
  int* flag_s = reinterpret_cast<int*>(sm + L.flags);    // [0] fold v, [1] fold u, [2] bad sum, [3] break code
  double* bak_f = reinterpret_cast<double*>(sm + L.bak_f);
  double* bak_g = reinterpret_cast<double*>(sm + L.bak_g);
  float4* kt_s = reinterpret_cast<float4*>(sm + L.kt);   // [warp][SRW][Gs]
  const float kAbsorb = P.absorb_log2;

  // ---- hand-over: potentials (log2 units), marginals, and the tile of Kt ------------------------------------
  for (int r = tid; r < rb_pad; r += kT2Threads) {
    const bool live = r < R;
    u_s[r] = 1.0f;
    LU_s[r] = live ? P.log_u[row_base + r] * kLog2e : 0.f;
    a_s[r] = live ? P.a[row_base + r] : 0.f;
  }
  for (int c = tid; c < nc4; c += kT2Threads) {
    const bool live = c < ncols;
    v_s[c] = live ? 1.0f : 0.f;
    LV_s[c] = live ? P.log_v[(int64_t)g_base * 4 + c] * kLog2e : 0.f;
    b_s[c] = live ? P.b[(int64_t)g_base * 4 + c] : 0.f;
  }
  if (tid < 8) flag_s[tid] = 0;
  __syncthreads();

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
      const float4 lv = reinterpret_cast<const float4*>(LV_s)[lg[g]];
      const double lu = (double)LU_s[lr], i2 = P.inv2;
      // exponent in fp64 (|M/reg| reaches hundreds: an fp32 fma would leave ~1e-5 relative in the entry), then ex2
      k.x = ex2f((float)(lu + (double)lv.x - (double)m.x * i2));
      k.y = ex2f((float)(lu + (double)lv.y - (double)m.y * i2));
      k.z = ex2f((float)(lu + (double)lv.z - (double)m.z * i2));
      k.w = ex2f((float)(lu + (double)lv.w - (double)m.w * i2));
    }
    return k;
  };
  float4 kreg[kT2RRW][kT2QG];
#pragma unroll
  for (int r = 0; r < kT2RRW; ++r)
#pragma unroll
    for (int g = 0; g < kT2QG; ++g) kreg[r][g] = (r < nrw) ? build(wrow0 + r, g) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sr = 0; sr < SRW; ++sr)
#pragma unroll
    for (int g = 0; g < kT2QG; ++g)
      if (lg[g] < Gs) kt_s[(warp * SRW + sr) * Gs + lg[g]] = build(wrow0 + kT2RRW + sr, g);
  __syncthreads();

  unsigned int target = 0;
  int cpt = P.start_iter, sweeps = P.start_iter;
  double err = 1.0;
  bool stop_hit = false;
  int pending_slot = -1, pending_cpt = 0;      // a marginal-error check whose sum is complete after the next barrier
  const bool timer = (blockIdx.x == 0 && tid == 0);
  const uint32_t rowpart_saddr = (uint32_t)__cvta_generic_to_shared(rowpart);

  for (cpt = P.start_iter; cpt < P.max_iter; ++cpt) {
    const int par = cpt & 1;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    if (timer) t0 = gtime();
    // ---- C: column partials of my tile (thread: 3 column groups x the warp's rows) ---------------------------
    {
      float4 acc[kT2QG];
#pragma unroll
      for (int g = 0; g < kT2QG; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r) {
        const float ur = u_s[wrow0 + r];
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) fma4(acc[g], kreg[r][g], ur);
      }
      const float4* kw = kt_s + warp * SRW * Gs;
#pragma unroll 2
      for (int sr = 0; sr < SRW; ++sr) {
        const float ur = u_s[wrow0 + kT2RRW + sr];
        float4 k[kT2QG];
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) k[g] = kw[sr * Gs + lgc[g]];
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) fma4(acc[g], k[g], ur);
      }
#pragma unroll
      for (int g = 0; g < kT2QG; ++g)
        if (lg[g] < Gs) reinterpret_cast<float4*>(red_c)[warp * Gs + lg[g]] = gok[g] ? acc[g] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float* my_part = P.part + ((size_t)(par * kT2CS + q) * NC) * nc4;
    if (tid < nc4) {
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kT2Warps; ++w2) s += red_c[w2 * nc4 + tid];
      my_part[(size_t)p * nc4 + tid] = s;
    }
    if (timer) t1 = gtime();
    t2_grid_barrier(st, target, nb);
    if (timer) t2 = gtime();
    // ---- all clusters' partials of MY columns -> v on my slice (every CTA of slice q computes the same bits) ----
    const bool check = (cpt >= 1) && ((cpt - 1) % 10 == 0);
    const int slot = check ? ((cpt - 1) / 10) & 127 : 0;
    if (check) {   // back-up of the iterate the check refers to (before sweep cpt touches it), as full potentials
      for (int r = tid; r < rb_pad; r += kT2Threads) bak_f[r] = (double)LU_s[r] + log2((double)u_s[r]);
    }
    {
      float s = 0.f;
      double d2 = 0.0;
      if (tid < nc4) {
        // all partials requested before the first add (one L2 round trip, not NC of them); summed in cluster order
        const float* src = my_part + tid;
        for (int pp = 0; pp < NC; pp += 16) {
          float t[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) t[k] = (pp + k < NC) ? ld_relaxed_f32(src + (size_t)(pp + k) * nc4) : 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) s += t[k];
        }
        const bool live = tid < ncols;
        const float v_old = v_s[tid];
        if (check) {
          bak_g[tid] = live ? (double)LV_s[tid] + log2((double)v_old) : 0.0;
          if (live && p == 0) {
            const double d = (double)(v_old * s) - (double)b_s[tid];
            d2 = d * d;
          }
        }
        if (live) {
          const float vn = b_s[tid] / s;
          if (!(s > 0.f) || !(vn < CUDART_INF_F) || !(vn > 0.f)) flag_s[2] = 1;
          if (fabsf(lg2f(vn)) > kAbsorb) flag_s[0] = 1;
          v_s[tid] = vn;
        }
      }
      if (check && p == 0) {                 // CTA-uniform: every lane of every warp takes part (d2 = 0 past nc4)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        if (lane == 0 && d2 != 0.0) atomicAdd(&st->err2[slot], d2);
      }
      // flags raised during an EARLIER sweep are complete and identical for every CTA after this sweep's barrier
      if (tid == 0) {
        const int code = ld_relaxed_s32(&st->flag_code);
        int brk = 0;
        if (code != 0 && (0x7fffffff - code) < cpt) brk = 1;
        if (brk == 0 && pending_slot >= 0) {
          const double e2 = ld_relaxed_f64(&st->err2[pending_slot]);
          if (!(sqrt(e2) > P.stop_thr)) brk = 2;
        }
        flag_s[3] = brk;
      }
    }
    __syncthreads();
    if (timer) t3 = gtime();
    {
      const int brk = flag_s[3];
      if (pending_slot >= 0) {
        err = sqrt(ld_relaxed_f64(&st->err2[pending_slot]));     // complete: same value in every thread
        pending_slot = -1;
      }
      if (brk == 1) break;
      if (brk == 2) { stop_hit = true; break; }
      if (flag_s[2] && tid == 0) {
        atomicMax(&st->fallback, 1);
        atomicMax(&st->flag_code, 0x7fffffff - cpt);
      }
      if (check) { pending_slot = slot; pending_cpt = cpt; }
    }
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
      for (int sr = 0; sr < SRW; ++sr)
#pragma unroll
        for (int g = 0; g < kT2QG; ++g)
          if (gok[g]) scale4(kt_s[(warp * SRW + sr) * Gs + lg[g]], 1.0f, vq[g]);
      if (tid < ncols) { LV_s[tid] += log2f(v_s[tid]); v_s[tid] = 1.0f; }
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
      for (int r = 0; r < kT2RRW; ++r)
        dots[r] = (dot4(kreg[r][0], vq[0]) + dot4(kreg[r][1], vq[1])) + dot4(kreg[r][2], vq[2]);
      const float4* kw = kt_s + warp * SRW * Gs;
#pragma unroll
      for (int sr = 0; sr < kT2MaxRowsPerWarp - kT2RRW; ++sr) {
        if (sr < SRW) {
          float4 k[kT2QG];
#pragma unroll
          for (int g = 0; g < kT2QG; ++g) k[g] = kw[sr * Gs + lgc[g]];
          dots[kT2RRW + sr] = (dot4(k[0], vq[0]) + dot4(k[1], vq[1])) + dot4(k[2], vq[2]);
        }
      }
      const float tot = warp_transpose_sum16(dots, lane);
      if (timer) t4 = gtime();
      const int r = lane & 15;
      if (r < nrw) {
        const uint32_t local = rowpart_saddr + 4u * (uint32_t)((par * kT2CS + q) * rb_pad + wrow0 + r);
        const int q0 = (lane >> 4) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) st_cluster_f32(mapa_shared(local, (uint32_t)(q0 + k)), tot);
      }
    }
    cluster_sync_all();
    if (tid < rb_pad) {
      const float* rp = rowpart + (par * kT2CS) * rb_pad + tid;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kT2CS; ++k) s += rp[k * rb_pad];
      if (tid < R) {
        const float un = a_s[tid] / s;
        if (!(s > 0.f) || !(un < CUDART_INF_F) || !(un > 0.f)) flag_s[2] = 1;
        if (fabsf(lg2f(un)) > kAbsorb) flag_s[1] = 1;
        u_s[tid] = un;
      }
    }
    __syncthreads();
    if (flag_s[1]) {
      // fold u into the tile and the row potentials (every CTA of this cluster takes the same decision)
#pragma unroll
      for (int r = 0; r < kT2RRW; ++r) {
        const float ur = u_s[wrow0 + r];
#pragma unroll
        for (int g = 0; g < kT2QG; ++g) { kreg[r][g].x *= ur; kreg[r][g].y *= ur; kreg[r][g].z *= ur; kreg[r][g].w *= ur; }
      }
      for (int sr = 0; sr < SRW; ++sr) {
        const float ur = u_s[wrow0 + kT2RRW + sr];
#pragma unroll
        for (int g = 0; g < kT2QG; ++g)
          if (gok[g]) { float4& k = kt_s[(warp * SRW + sr) * Gs + lg[g]]; k.x *= ur; k.y *= ur; k.z *= ur; k.w *= ur; }
      }
      __syncthreads();
      if (tid < R) { LU_s[tid] += log2f(u_s[tid]); u_s[tid] = 1.0f; }
      if (tid == 0) flag_s[1] = 0;
      __syncthreads();
    }
    sweeps = cpt + 1;
    if (timer) {
      t5 = gtime();
      st->t_phase[0] += t1 - t0; st->t_phase[1] += t2 - t1; st->t_phase[2] += t3 - t2;
      st->t_phase[3] += t4 - t3; st->t_phase[4] += t5 - t4;
    }
  }
  // a check issued in the very last sweep: its sum is complete one barrier later
  if (cpt >= P.max_iter && pending_slot >= 0) {
    if (flag_s[2] && tid == 0) { atomicMax(&st->fallback, 1); }
    t2_grid_barrier(st, target, nb);
    err = sqrt(ld_relaxed_f64(&st->err2[pending_slot]));
    if (!(err > P.stop_thr)) stop_hit = true;
  } else if (cpt >= P.max_iter) {
    if (flag_s[2] && tid == 0) atomicMax(&st->fallback, 1);     // last sweep's sums: no later barrier publishes them
  }
  __syncthreads();
  // ---- outputs (natural-log potentials, composed in fp64) ------------------------------------------------------
  const double ln2 = 0.69314718055994530942;
  if (q == 0) {
    for (int r = tid; r < R; r += kT2Threads) {
      const double f = stop_hit ? bak_f[r] : (double)LU_s[r] + log2((double)u_s[r]);
      P.log_u[row_base + r] = (float)(f * ln2);
    }
  }
  if (p == 0) {
    for (int c = tid; c < ncols; c += kT2Threads) {
      const double g = stop_hit ? bak_g[c] : (double)LV_s[c] + log2((double)v_s[c]);
      P.log_v[(int64_t)g_base * 4 + c] = (float)(g * ln2);
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    st->sweeps = stop_hit ? pending_cpt : sweeps;
    st->final_buf = 0;
    st->err = err;
    if (P.force_fallback) atomicMax(&st->fallback, 1);
  }
}

// The communication skeleton of one sweep with no mat-vec work: column partials through global memory + the grid
// barrier + NC partial reads, then the row-partial push through distributed shared memory + the cluster barrier.
// bench.py times it to state the latency floor of this design next to the measured sweep time.
__global__ void __launch_bounds__(kT2Threads, 1)
sinkhorn_tile2d_sync_floor_kernel(PersistState* st, float* part, int NC, int Gs, int nrw, int iters) {
  extern __shared__ __align__(16) unsigned char smem_raw_t2[];
  float* sm = reinterpret_cast<float*>(smem_raw_t2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank(), p = (int)cluster_idx();
  const unsigned int nb = (unsigned int)(NC * kT2CS);
  const int nc4 = Gs * 4, rb_pad = kT2Warps * nrw + kT2MaxRowsPerWarp;
  const T2Smem L = t2_carve(nrw, Gs);
  float* u_s = sm + L.u; float* v_s = sm + L.v; float* red_c = sm + L.red_c; float* rowpart = sm + L.rowpart;
  const uint32_t rowpart_saddr = (uint32_t)__cvta_generic_to_shared(rowpart);
  for (int i = tid; i < kT2Warps * nc4; i += kT2Threads) red_c[i] = 1.0f;
  for (int i = tid; i < rb_pad; i += kT2Threads) u_s[i] = 1.0f;
  __syncthreads();
  unsigned int target = 0;
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    __syncthreads();
    float* my_part = part + ((size_t)(par * kT2CS + q) * NC) * nc4;
    if (tid < nc4) {
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kT2Warps; ++w2) s += red_c[w2 * nc4 + tid];
      my_part[(size_t)p * nc4 + tid] = s * u_s[0];
    }
    t2_grid_barrier(st, target, nb);
    if (tid < nc4) {
      float s = 0.f;
      for (int pp = 0; pp < NC; pp += 16) {
        float t[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) t[k] = (pp + k < NC) ? ld_relaxed_f32(my_part + (size_t)(pp + k) * nc4 + tid) : 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += t[k];
      }
      v_s[tid] = 1.0f / s;
    }
    __syncthreads();
    const float tot = v_s[lane];
    const int r = lane & 15;
    if (r < nrw) {
      const uint32_t local = rowpart_saddr + 4u * (uint32_t)((par * kT2CS + q) * rb_pad + warp * nrw + r);
      const int q0 = (lane >> 4) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) st_cluster_f32(mapa_shared(local, (uint32_t)(q0 + k)), tot);
    }
    cluster_sync_all();
    if (tid < kT2Warps * nrw) {
      const float* rp = rowpart + (par * kT2CS) * rb_pad + tid;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kT2CS; ++k) s += rp[k * rb_pad];
      u_s[tid] = 1.0f / s;
    }
  }
}

int sinkhorn_tile2d_absorbs_read() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_dev_tile2d_absorbs, sizeof(int));
  return v;
}

// Co-residency probe: the same launch shape (cluster of 8, 512 threads, the solver's dynamic shared memory) doing
// ONE grid barrier with a 5 ms budget.  If some cluster cannot be resident together with the others the barrier
// times out and *ok is cleared — the occupancy query is a necessary condition, this is the sufficient one.
__global__ void __launch_bounds__(kT2Threads, 1)
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
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kT2CS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kT2Threads);
  cfg.stream = s;
  cfg.attrs = attrs;
  cfg.dynamicSmemBytes = smem;
  cfg.gridDim = dim3(kT2CS * kT2MaxClusters);
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  n = std::min(n, kT2MaxClusters);
  int verified = -1;
  for (; n >= 4; --n) {
    EG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(PersistState), s));
    int one = 1;
    EG_CUDA(cudaMemcpyAsync(&scratch->sweeps, &one, sizeof(int), cudaMemcpyHostToDevice, s));
    cfg.gridDim = dim3((unsigned)(kT2CS * n));
    cfg.numAttrs = 2;
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
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kT2CS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
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
    cfg->numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, cfg) != cudaSuccess) { cudaGetLastError(); return EG_OK; }
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
  cfg->numAttrs = 2;
  return EG_OK;
}

int sinkhorn_tile2d_sync_floor_launch(int64_t I, int64_t J, int iters, float* part, size_t part_floats,
                                      PersistState* st, cudaStream_t s, bool* launched) {
  *launched = false;
  if (J % 4 != 0 || J > 4 * kT2CS * 32 * kT2QG || iters <= 0) return EG_OK;
  T2Geometry g;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  auto kern = sinkhorn_tile2d_sync_floor_kernel;
  int rc = t2_geometry(kern, I, J, s, st, &g, &cfg, attrs);
  if (rc) return rc;
  if (g.nc == 0 || (size_t)2 * kT2CS * g.nc * g.gs * 4 > part_floats) return EG_OK;
  EG_CUDA(cudaMemsetAsync(st, 0, sizeof(PersistState), s));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, st, part, g.nc, g.gs, g.nrw, iters);
  if (e != cudaSuccess) { cudaGetLastError(); return EG_OK; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  *launched = true;
  return EG_OK;
}

int sinkhorn_tile2d_launch(const float* M, int64_t I, int64_t J, int64_t ld, double inv_reg, const float* a,
                           const float* b, float* log_u, float* log_v, const PersistState* warm, int start_iter,
                           int max_iter, double stop_thr, float* part, size_t part_floats, PersistState* st,
                           float absorb_log2, int force_fallback, cudaStream_t s, bool* launched) {
  *launched = false;
  if (J % 4 != 0 || (reinterpret_cast<uintptr_t>(M) & 15) || (ld % 4) != 0 || J > 4 * kT2CS * 32 * kT2QG) return EG_OK;
  T2Geometry g;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  auto kern = sinkhorn_tile2d_kernel;
  int rc = t2_geometry(kern, I, J, s, st, &g, &cfg, attrs);
  if (rc) return rc;
  if (g.nc == 0 || (size_t)2 * kT2CS * g.nc * g.gs * 4 > part_floats) return EG_OK;
  EG_CUDA(cudaMemsetAsync(st, 0, sizeof(PersistState), s));      // the probe may have used it
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
