// (a) GCN / highway-GCN message passing:  out = epilogue(A · H)  over a CSR adjacency.
//
// Replaces torch.spmm + activation + highway blend of layers/layers.py:35-38,64-76.
// HBM-bound: per row the kernel streams (col, val) once, gathers the neighbour
// feature rows with 16-byte coalesced loads (a 300-float row is 75 float4 =
// 2.34 warp-wide requests) and writes the output row once; the epilogue reads
// gate_pre / x_res once.  One warp owns one row, so the reduction over
// neighbours stays in registers and the summation order is the CSR order
// (deterministic).  Hub rows of power-law graphs are cut into fixed-length
// segments handled by extra warps; their partial sums are combined, in order,
// by a second small launch that also applies the epilogue.
#include <algorithm>

#include "common.cuh"

namespace eg {

constexpr int kWarpsPerBlock = 8;   // scalar fallback kernel only

// Tuning knobs (defaults chosen from the ncu study in profiles/); eg_debug_set() overrides them.
int g_tune_unroll = 2;      // neighbour rows in flight per warp (x VPL float4 each).  0 = by row width (4 up to 64 float4 per
                            // row): wins on the benchmark graph at d = 128 (0.218 vs 0.257 ms, ~11 neighbours per row) but
                            // loses on the 10M-node power-law graph (5.78 vs 5.40 ms, ~5 per row), so 2 stays the default
int g_tune_warps = 4;       // warps (= rows) per CTA
int g_tune_spmm_persist = 0;  // eg_debug_set(6, n): n > 0 -> persistent pipelined SpMM with n CTAs per SM
int g_tune_spmm_slab = 0;   // eg_debug_set(14, v): v in {32, 64}: walk the feature columns in slabs of v float4 (one launch
                            // per slab, rows inner) so that a slab of H stays L2-resident; 0: as wide as the kernel allows
int g_tune_hints = 0;       // L2 eviction-priority hints (gathers evict_last, streams evict_first): no measured gain
int g_tune_spmm_dynamic = 0;  // eg_debug_set(17, 1): the persistent SpMM hands out rows through an atomic counter
int g_tune_spmm_bulk = 0;   // eg_debug_set(16, 1): neighbour rows fetched by cp.async.bulk into shared memory (spmm_bulk_kernel)

struct Epilogue {
  const float* gate_pre;
  const float* x_res;
  float* out;
  float* act_out;
  int act;
};

struct Policies {
  unsigned long long keep, stream;
  bool on;
};

__device__ __forceinline__ Policies make_policies(bool on) {
  Policies p;
  p.on = on;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  return p;
}

// feature-row gather: keep in L2 (rows are re-read by other warps)
__device__ __forceinline__ float4 ld_keep_f4(const float4* ptr, const Policies& p) {
  float4 r;
  if (p.on)
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr), "l"(p.keep));
  else
    r = __ldg(ptr);
  return r;
}
// touched-once data: do not displace the feature rows
__device__ __forceinline__ float4 ld_once_f4(const float4* ptr, const Policies& p) {
  float4 r;
  if (p.on)
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr), "l"(p.stream));
  else
    r = ld_stream_f4(ptr);
  return r;
}
__device__ __forceinline__ void st_once_f4(float4* ptr, const float4& v, const Policies& p) {
  if (p.on)
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(p.stream) : "memory");
  else
    st_stream_f4(ptr, v);
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ float4 apply_epilogue(const Epilogue& ep, float4 s, int64_t off4, const Policies& pol) {
  if (ep.act == EG_ACT_RELU) {
    s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f);
  }
  if (ep.act_out) st_once_f4(reinterpret_cast<float4*>(ep.act_out) + off4, s, pol);
  if (ep.gate_pre) {
    float4 g = ld_once_f4(reinterpret_cast<const float4*>(ep.gate_pre) + off4, pol);
    float4 x = ld_once_f4(reinterpret_cast<const float4*>(ep.x_res) + off4, pol);
    float t;
    t = sigmoidf_(g.x); s.x = t * s.x + (1.0f - t) * x.x;
    t = sigmoidf_(g.y); s.y = t * s.y + (1.0f - t) * x.y;
    t = sigmoidf_(g.z); s.z = t * s.z + (1.0f - t) * x.z;
    t = sigmoidf_(g.w); s.w = t * s.w + (1.0f - t) * x.w;
  }
  return s;
}

__device__ __forceinline__ float apply_epilogue1(const Epilogue& ep, float s, int64_t off) {
  if (ep.act == EG_ACT_RELU) s = fmaxf(s, 0.f);
  if (ep.act_out) ep.act_out[off] = s;
  if (ep.gate_pre) {
    float t = sigmoidf_(ep.gate_pre[off]);
    s = t * s + (1.0f - t) * ep.x_res[off];
  }
  return s;
}

// acc += sum_{i in [b,e)} val[i] * H[col[i], chunk], VPL float4 per lane, UNROLL rows in flight.
// `Hc` already points at the column chunk; `d4` is the full row stride in float4.
template <int VPL, int UNROLL>
__device__ __forceinline__ void gather_rows(const int32_t* __restrict__ col, const float* __restrict__ val,
                                            const float4* __restrict__ Hc, int d4, int d4_local, int b, int e,
                                            int lane, const Policies& pol, float4 (&acc)[VPL]) {
  for (int base = b; base < e; base += 32) {
    int idx = base + lane;
    int my_col = 0;
    float my_val = 0.f;
    if (idx < e) {
      my_col = ld_stream_i32(col + idx);
      my_val = ld_stream_f32(val + idx);
    }
    int cnt = min(32, e - base);
    for (int t = 0; t < cnt; t += UNROLL) {
      float4 x[UNROLL][VPL];
      float v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        int c = __shfl_sync(0xffffffffu, my_col, (t + u) & 31);
        float vv = __shfl_sync(0xffffffffu, my_val, (t + u) & 31);
        bool live = (t + u < cnt);
        v[u] = live ? vv : 0.f;
        const float4* rowp = Hc + (int64_t)c * d4;
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
          int k = lane + 32 * p;
          x[u][p] = (live && k < d4_local) ? ld_keep_f4(rowp + k, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
          acc[p].x = fmaf(v[u], x[u][p].x, acc[p].x);
          acc[p].y = fmaf(v[u], x[u][p].y, acc[p].y);
          acc[p].z = fmaf(v[u], x[u][p].z, acc[p].z);
          acc[p].w = fmaf(v[u], x[u][p].w, acc[p].w);
        }
      }
    }
  }
}

// Warps [0, n_rows): one short row each (rows longer than `thresh` are skipped here).
// Warps [n_rows, n_rows + n_seg): one segment of a long row each -> seg_scratch.
template <int VPL, int UNROLL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
spmm_vec_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const float* __restrict__ val, int64_t n_rows, const float* __restrict__ H, int d4,
                int chunk0, Epilogue ep, int thresh, const int32_t* __restrict__ seg_begin,
                const int32_t* __restrict__ seg_end, int64_t n_seg, float* __restrict__ seg_scratch,
                int hints) {
  int lane = threadIdx.x & 31;
  int64_t w = blockIdx.x * (int64_t)WARPS + (threadIdx.x >> 5);
  if (w >= n_rows + n_seg) return;
  const Policies pol = make_policies(hints != 0);
  const float4* Hc = reinterpret_cast<const float4*>(H) + chunk0;
  int d4_local = min(d4 - chunk0, 32 * VPL);
  float4 acc[VPL];
#pragma unroll
  for (int p = 0; p < VPL; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (w < n_rows) {
    int b = rowptr[w], e = rowptr[w + 1];
    if (e - b > thresh) return;
    gather_rows<VPL, UNROLL>(col, val, Hc, d4, d4_local, b, e, lane, pol, acc);
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
      int k = lane + 32 * p;
      if (k < d4_local) {
        int64_t off4 = w * d4 + chunk0 + k;
        float4 r = apply_epilogue(ep, acc[p], off4, pol);
        st_once_f4(reinterpret_cast<float4*>(ep.out) + off4, r, pol);
      }
    }
  } else {
    int64_t sidx = w - n_rows;
    gather_rows<VPL, UNROLL>(col, val, Hc, d4, d4_local, seg_begin[sidx], seg_end[sidx], lane, pol, acc);
    float4* dst = reinterpret_cast<float4*>(seg_scratch) + sidx * d4 + chunk0;
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
      int k = lane + 32 * p;
      if (k < d4_local) dst[k] = acc[p];
    }
  }
}

// ---- bulk-copy variant: the feature rows of a CSR row travel as cp.async.bulk copies into shared memory ---------
// The register-gather kernel above keeps 2 neighbour rows (3 KB) in flight per warp and ~25 warps per SM: 75 KB per SM
// against a ~1.7 us gather latency = the 6 TB/s of L2 -> SM traffic the captures show (Little's law, not a bandwidth
// limit).  Here lane 0 of a warp hands the copy engine one 16-byte-aligned row (d4 * 16 bytes) per neighbour, up to
// SLOTS rows ahead — for the benchmark graph's ~11 neighbours per row the WHOLE row is in flight at once — each landing
// in a shared-memory slot guarded by its own mbarrier (complete_tx); the warp then streams the slots through the FMAs
// in CSR order (same summation order, same bits).  Bytes in flight are bounded by shared memory (SLOTS x 1.2 KB per
// warp), not by registers.
__device__ __forceinline__ void bulk_row_to_smem(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst), b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bar_wait_parity(uint64_t* bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(b), "r"(parity) : "memory");
}

template <int VPL, int SLOTS>
__device__ __forceinline__ void gather_rows_bulk(const int32_t* __restrict__ col, const float* __restrict__ val,
                                                 const float4* __restrict__ Hc, int d4, int d4_local, int b, int e,
                                                 int lane, unsigned char* slots, uint32_t slot_bytes, uint64_t* bars,
                                                 uint32_t& ph, float4 (&acc)[VPL]) {
  for (int base = b; base < e; base += 32) {
    const int idx = base + lane;
    int my_col = 0;
    float my_val = 0.f;
    if (idx < e) {
      my_col = ld_stream_i32(col + idx);
      my_val = ld_stream_f32(val + idx);
    }
    const int cnt = min(32, e - base);
    // up to SLOTS rows of this chunk go out at once
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
      const int c = __shfl_sync(0xffffffffu, my_col, t);
      if (t < cnt && lane == 0) bulk_row_to_smem(slots + t * slot_bytes, Hc + (int64_t)c * d4, slot_bytes, bars + t);
    }
    int s = 0;
    for (int t = 0; t < cnt; ++t) {
      bar_wait_parity(bars + s, (ph >> s) & 1u);
      ph ^= 1u << s;
      const float v = __shfl_sync(0xffffffffu, my_val, t);
      const float4* row = reinterpret_cast<const float4*>(slots + s * slot_bytes);
#pragma unroll
      for (int p = 0; p < VPL; ++p) {
        const int k = lane + 32 * p;
        if (k < d4_local) {
          const float4 x = row[k];
          acc[p].x = fmaf(v, x.x, acc[p].x);
          acc[p].y = fmaf(v, x.y, acc[p].y);
          acc[p].z = fmaf(v, x.z, acc[p].z);
          acc[p].w = fmaf(v, x.w, acc[p].w);
        }
      }
      __syncwarp();                                           // every lane is done with the slot: it may be refilled
      const int tn = t + SLOTS;
      const int cn = __shfl_sync(0xffffffffu, my_col, tn & 31);
      if (tn < cnt && lane == 0) bulk_row_to_smem(slots + s * slot_bytes, Hc + (int64_t)cn * d4, slot_bytes, bars + s);
      if (++s == SLOTS) s = 0;
    }
  }
}

template <int VPL, int WARPS, int SLOTS>
__global__ void __launch_bounds__(WARPS * 32)
spmm_bulk_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                 int64_t n_rows, const float* __restrict__ H, int d4, int chunk0, Epilogue ep, int thresh,
                 const int32_t* __restrict__ seg_begin, const int32_t* __restrict__ seg_end, int64_t n_seg,
                 float* __restrict__ seg_scratch) {
  extern __shared__ __align__(128) unsigned char spmm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int d4_local = min(d4 - chunk0, 32 * VPL);
  const uint32_t slot_bytes = (uint32_t)d4_local * 16u;
  unsigned char* slots = spmm_smem + (size_t)wib * SLOTS * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(spmm_smem + (size_t)WARPS * SLOTS * slot_bytes) + wib * SLOTS;
  if (lane == 0) {
    for (int i = 0; i < SLOTS; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t w = blockIdx.x * (int64_t)WARPS + wib;
  if (w >= n_rows + n_seg) return;
  const Policies pol = make_policies(false);
  const float4* Hc = reinterpret_cast<const float4*>(H) + chunk0;
  float4 acc[VPL];
#pragma unroll
  for (int p = 0; p < VPL; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t ph = 0;
  if (w < n_rows) {
    const int b = rowptr[w], e = rowptr[w + 1];
    if (e - b > thresh) return;
    gather_rows_bulk<VPL, SLOTS>(col, val, Hc, d4, d4_local, b, e, lane, slots, slot_bytes, bars, ph, acc);
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
      const int k = lane + 32 * p;
      if (k < d4_local) {
        const int64_t off4 = w * d4 + chunk0 + k;
        const float4 r = apply_epilogue(ep, acc[p], off4, pol);
        st_once_f4(reinterpret_cast<float4*>(ep.out) + off4, r, pol);
      }
    }
  } else {
    const int64_t sidx = w - n_rows;
    gather_rows_bulk<VPL, SLOTS>(col, val, Hc, d4, d4_local, seg_begin[sidx], seg_end[sidx], lane, slots, slot_bytes, bars,
                                 ph, acc);
    float4* dst = reinterpret_cast<float4*>(seg_scratch) + sidx * d4 + chunk0;
#pragma unroll
    for (int p = 0; p < VPL; ++p) {
      const int k = lane + 32 * p;
      if (k < d4_local) dst[k] = acc[p];
    }
  }
}

// Persistent, software-pipelined variant: each warp walks work items w, w + stride, … and fetches the NEXT
// item's bounds and first (col, val) chunk while the current item's feature-row gathers are in flight, so the
// rowptr -> col/val -> feature-row dependency chain (three L2/HBM round trips for ~11 neighbours of work)
// is paid once per warp instead of once per row.
template <int VPL, int UNROLL>
__device__ __forceinline__ void gather_chunk(const float4* __restrict__ Hc, int d4, int d4_local, int lane, int cnt,
                                             int my_col, float my_val, const Policies& pol, float4 (&acc)[VPL]) {
  for (int t = 0; t < cnt; t += UNROLL) {
    float4 x[UNROLL][VPL];
    float v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int c = __shfl_sync(0xffffffffu, my_col, (t + u) & 31);
      float vv = __shfl_sync(0xffffffffu, my_val, (t + u) & 31);
      bool live = (t + u < cnt);
      v[u] = live ? vv : 0.f;
      const float4* rowp = Hc + (int64_t)c * d4;
#pragma unroll
      for (int p = 0; p < VPL; ++p) {
        int k = lane + 32 * p;
        x[u][p] = (live && k < d4_local) ? ld_keep_f4(rowp + k, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int p = 0; p < VPL; ++p) {
        acc[p].x = fmaf(v[u], x[u][p].x, acc[p].x);
        acc[p].y = fmaf(v[u], x[u][p].y, acc[p].y);
        acc[p].z = fmaf(v[u], x[u][p].z, acc[p].z);
        acc[p].w = fmaf(v[u], x[u][p].w, acc[p].w);
      }
    }
  }
}

template <int VPL, int UNROLL>
__global__ void __launch_bounds__(256)
spmm_persist_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const float* __restrict__ val, int64_t n_rows, const float* __restrict__ H, int d4,
                    int chunk0, Epilogue ep, int thresh, const int32_t* __restrict__ seg_begin,
                    const int32_t* __restrict__ seg_end, int64_t n_seg, float* __restrict__ seg_scratch,
                    int hints, unsigned long long* __restrict__ next_item) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * 8;
  const int64_t total = n_rows + n_seg;
  int64_t w = blockIdx.x * (int64_t)8 + (threadIdx.x >> 5);
  if (w >= total) return;
  // next_item != null: work items are handed out by an atomic counter (first-come first-served: no row-length
  // imbalance between warps) instead of the static round-robin; the counter starts at the number of warps in the grid
  auto grab = [&]() -> int64_t {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(next_item, 1ull);
    return (int64_t)__shfl_sync(0xffffffffu, v, 0);
  };
  int64_t w_after = next_item ? grab() : w + stride;        // the item after this one (fetched one item ahead)
  const Policies pol = make_policies(hints != 0);
  const float4* Hc = reinterpret_cast<const float4*>(H) + chunk0;
  const int d4_local = min(d4 - chunk0, 32 * VPL);
  auto bounds = [&](int64_t item, int& b, int& e) {
    if (item < n_rows) {
      b = __ldg(rowptr + item);
      e = __ldg(rowptr + item + 1);
      if (e - b > thresh) e = b - 1;            // long row: handled by its segments; marks "skip"
    } else {
      b = __ldg(seg_begin + (item - n_rows));
      e = __ldg(seg_end + (item - n_rows));
    }
  };
  int b, e;
  bounds(w, b, e);
  int my_col = 0;
  float my_val = 0.f;
  if (b + lane < e) { my_col = ld_stream_i32(col + b + lane); my_val = ld_stream_f32(val + b + lane); }
  while (true) {
    const int64_t wn = w_after;
    const bool has_next = wn < total;
    if (has_next) w_after = next_item ? grab() : wn + stride;
    int nb = 0, ne = 0;
    if (has_next) bounds(wn, nb, ne);                        // independent loads, issued before the gathers
    float4 acc[VPL];
#pragma unroll
    for (int p = 0; p < VPL; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool skip = e < b;
    if (!skip) {
      gather_chunk<VPL, UNROLL>(Hc, d4, d4_local, lane, min(32, e - b), my_col, my_val, pol, acc);
      for (int base = b + 32; base < e; base += 32) {
        int c2 = 0;
        float v2 = 0.f;
        if (base + lane < e) { c2 = ld_stream_i32(col + base + lane); v2 = ld_stream_f32(val + base + lane); }
        gather_chunk<VPL, UNROLL>(Hc, d4, d4_local, lane, min(32, e - base), c2, v2, pol, acc);
      }
    }
    int ncol = 0;
    float nval = 0.f;
    if (has_next && nb + lane < ne) { ncol = ld_stream_i32(col + nb + lane); nval = ld_stream_f32(val + nb + lane); }
    if (!skip) {
      if (w < n_rows) {
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
          int k = lane + 32 * p;
          if (k < d4_local) {
            int64_t off4 = w * d4 + chunk0 + k;
            float4 r = apply_epilogue(ep, acc[p], off4, pol);
            st_once_f4(reinterpret_cast<float4*>(ep.out) + off4, r, pol);
          }
        }
      } else {
        float4* dst = reinterpret_cast<float4*>(seg_scratch) + (w - n_rows) * d4 + chunk0;
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
          int k = lane + 32 * p;
          if (k < d4_local) dst[k] = acc[p];
        }
      }
    }
    if (!has_next) break;
    w = wn; b = nb; e = ne; my_col = ncol; my_val = nval;
  }
}

// Finish long rows: sum their segment partials in segment order, then epilogue.
__global__ void spmm_long_finish_kernel(const int32_t* __restrict__ long_rows,
                                        const int32_t* __restrict__ long_first, int64_t n_long,
                                        const float* __restrict__ seg_scratch, int d, Epilogue ep) {
  int64_t r = blockIdx.x;
  if (r >= n_long) return;
  int row = long_rows[r];
  int s0 = long_first[r], s1 = long_first[r + 1];
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = 0.f;
    for (int s = s0; s < s1; ++s) acc += seg_scratch[(int64_t)s * d + k];
    int64_t off = (int64_t)row * d + k;
    ep.out[off] = apply_epilogue1(ep, acc, off);
  }
}

// Generic-d fallback (d % 4 != 0): scalar lanes, any d.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
spmm_scalar_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                   const float* __restrict__ val, int64_t n_rows, const float* __restrict__ H, int d,
                   Epilogue ep, int thresh, const int32_t* __restrict__ seg_begin,
                   const int32_t* __restrict__ seg_end, int64_t n_seg, float* __restrict__ seg_scratch) {
  int lane = threadIdx.x & 31;
  int64_t w = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5);
  if (w >= n_rows + n_seg) return;
  int b, e;
  bool is_seg = w >= n_rows;
  if (!is_seg) {
    b = rowptr[w]; e = rowptr[w + 1];
    if (e - b > thresh) return;
  } else {
    b = seg_begin[w - n_rows]; e = seg_end[w - n_rows];
  }
  for (int k0 = 0; k0 < d; k0 += 32) {
    int k = k0 + lane;
    float acc = 0.f;
    for (int i = b; i < e; ++i) {
      int c = col[i];
      float v = val[i];
      if (k < d) acc = fmaf(v, __ldg(H + (int64_t)c * d + k), acc);
    }
    if (k < d) {
      if (!is_seg) {
        int64_t off = w * d + k;
        ep.out[off] = apply_epilogue1(ep, acc, off);
      } else {
        seg_scratch[(w - n_rows) * d + k] = acc;
      }
    }
  }
}

// ---- element-wise backward of the epilogue ------------------------------------
__global__ void epilogue_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ a,
                                    const float* __restrict__ gate_pre, const float* __restrict__ x_res,
                                    int64_t n, int act, float* __restrict__ dS, float* __restrict__ d_gate,
                                    float* __restrict__ d_xres) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = dout[i];
    float av = a ? a[i] : 0.f;
    float t = 1.0f;
    if (gate_pre) {
      t = sigmoidf_(gate_pre[i]);
      float x = x_res[i];
      if (d_gate) d_gate[i] = g * (av - x) * (t * (1.0f - t));
      if (d_xres) d_xres[i] = g * (1.0f - t);
    }
    float ds = g * t;
    if (act == EG_ACT_RELU && !(av > 0.f)) ds = 0.f;
    dS[i] = ds;
  }
}

// float4 version: all operands 16-byte aligned, n a multiple of 4 (the layers' [rows, 300] arrays).  Seven streams of
// 240 MB at the benchmark size: HBM-bound, every load issued before the first use.
__global__ void __launch_bounds__(256) epilogue_bwd_vec4_kernel(const float4* __restrict__ dout, const float4* __restrict__ a,
                                                                const float4* __restrict__ gate_pre,
                                                                const float4* __restrict__ x_res, int64_t n4, int act,
                                                                float4* __restrict__ dS, float4* __restrict__ d_gate,
                                                                float4* __restrict__ d_xres) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g4 = __ldcs(dout + i);
    const float4 a4 = a ? __ldcs(a + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), x4 = p4;
    if (gate_pre) { p4 = __ldcs(gate_pre + i); x4 = __ldcs(x_res + i); }
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
    const float pre[4] = {p4.x, p4.y, p4.z, p4.w}, x[4] = {x4.x, x4.y, x4.z, x4.w};
    float ds[4], dg[4], dx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = 1.0f;
      dg[e] = dx[e] = 0.f;
      if (gate_pre) {
        t = sigmoidf_(pre[e]);
        dg[e] = g[e] * (av[e] - x[e]) * (t * (1.0f - t));
        dx[e] = g[e] * (1.0f - t);
      }
      ds[e] = g[e] * t;
      if (act == EG_ACT_RELU && !(av[e] > 0.f)) ds[e] = 0.f;
    }
    if (gate_pre && d_gate) d_gate[i] = make_float4(dg[0], dg[1], dg[2], dg[3]);
    if (gate_pre && d_xres) d_xres[i] = make_float4(dx[0], dx[1], dx[2], dx[3]);
    dS[i] = make_float4(ds[0], ds[1], ds[2], ds[3]);
  }
}

template <int VPL, int UNROLL, int WARPS>
static int launch_vec3(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_rows,
                       const float* H, int d4, int chunk0, const Epilogue& ep, int thresh,
                       const int32_t* seg_begin, const int32_t* seg_end, int64_t n_seg, float* seg_scratch,
                       cudaStream_t s) {
  int64_t warps = n_rows + n_seg;
  if (g_tune_spmm_bulk > 0) {
    constexpr int BW = 4, BS = 12;
    const int d4_local = std::min(d4 - chunk0, 32 * VPL);
    const size_t smem = (size_t)BW * BS * d4_local * 16 + (size_t)BW * BS * 8;
    auto kern = spmm_bulk_kernel<VPL, BW, BS>;
    static PerDeviceOnce once;
    EG_SET_SMEM_ONCE(once, EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BW * BS * (32 * VPL * 16 + 8))));
    kern<<<(unsigned)ceil_div(warps, BW), BW * 32, smem, s>>>(rowptr, col, val, n_rows, H, d4, chunk0, ep, thresh, seg_begin,
                                                              seg_end, n_seg, seg_scratch);
    EG_LAUNCHED();
    return EG_OK;
  }
  if (g_tune_spmm_persist > 0) {
    int64_t want = ceil_div(warps, 8);
    unsigned pgrid = (unsigned)std::min<int64_t>(want, (int64_t)kNumSMs * g_tune_spmm_persist);
    unsigned long long* counter = nullptr;
    if (g_tune_spmm_dynamic) {
      static unsigned long long* dev_counter[64] = {};
      int dev = 0;
      EG_CUDA(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64) return EG_ERR_UNSUPPORTED;
      if (!dev_counter[dev]) EG_CUDA(cudaMalloc(&dev_counter[dev], sizeof(unsigned long long)));
      counter = dev_counter[dev];
      const unsigned long long first = (unsigned long long)pgrid * 8ull;
      EG_CUDA(cudaMemcpyAsync(counter, &first, sizeof(first), cudaMemcpyHostToDevice, s));
    }
    spmm_persist_kernel<VPL, UNROLL><<<pgrid, 256, 0, s>>>(rowptr, col, val, n_rows, H, d4, chunk0, ep, thresh,
                                                           seg_begin, seg_end, n_seg, seg_scratch, g_tune_hints, counter);
    EG_LAUNCHED();
    return EG_OK;
  }
  unsigned grid = (unsigned)ceil_div(warps, WARPS);
  spmm_vec_kernel<VPL, UNROLL, WARPS><<<grid, WARPS * 32, 0, s>>>(rowptr, col, val, n_rows, H, d4, chunk0, ep,
                                                                   thresh, seg_begin, seg_end, n_seg,
                                                                   seg_scratch, g_tune_hints);
  EG_LAUNCHED();
  return EG_OK;
}

template <int VPL>
static int launch_vec(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_rows,
                      const float* H, int d4, int chunk0, const Epilogue& ep, int thresh,
                      const int32_t* seg_begin, const int32_t* seg_end,
                      int64_t n_seg, float* seg_scratch, cudaStream_t s) {
  const int unroll = g_tune_unroll > 0 ? g_tune_unroll : (VPL <= 2 ? 4 : 2);
#define EG_SPMM_CASE(U, W)                                                                              \
  if (unroll == U && g_tune_warps == W)                                                          \
    return launch_vec3<VPL, U, W>(rowptr, col, val, n_rows, H, d4, chunk0, ep, thresh, seg_begin, seg_end, \
                                  n_seg, seg_scratch, s);
  EG_SPMM_CASE(4, 8) EG_SPMM_CASE(2, 8) EG_SPMM_CASE(4, 4) EG_SPMM_CASE(2, 4) EG_SPMM_CASE(1, 8) EG_SPMM_CASE(8, 4)
#undef EG_SPMM_CASE
  return launch_vec3<VPL, 2, 4>(rowptr, col, val, n_rows, H, d4, chunk0, ep, thresh, seg_begin, seg_end, n_seg,
                                seg_scratch, s);
}

}  // namespace eg

extern "C" {

int eg_spmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, const float* H,
            int d, int act, const float* gate_pre, const float* x_res, float* out, float* act_out,
            int long_row_threshold, const int32_t* seg_row, const int32_t* seg_begin,
            const int32_t* seg_end, int64_t n_seg, const int32_t* long_rows, const int32_t* long_first,
            int64_t n_long, float* seg_scratch, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || d <= 0 || !rowptr || !out || !H) return EG_ERR_INVALID;
  if ((gate_pre == nullptr) != (x_res == nullptr)) return EG_ERR_INVALID;
  if (act != EG_ACT_IDENTITY && act != EG_ACT_RELU) return EG_ERR_INVALID;
  if (n_seg < 0 || n_long < 0) return EG_ERR_INVALID;
  if (n_seg > 0 && (!seg_begin || !seg_end || !long_rows || !long_first || !seg_scratch || n_long == 0))
    return EG_ERR_INVALID;
  if (n_rows == 0) return EG_OK;
  if (long_row_threshold <= 0 || n_seg == 0) long_row_threshold = 0x7fffffff;
  cudaStream_t s = as_stream(stream_);
  Epilogue ep{gate_pre, x_res, out, act_out, act};
  bool vec_ok = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(out) |
                                   reinterpret_cast<uintptr_t>(gate_pre) | reinterpret_cast<uintptr_t>(x_res) |
                                   reinterpret_cast<uintptr_t>(act_out) | reinterpret_cast<uintptr_t>(seg_scratch)) % 16 == 0);
  if (vec_ok) {
    int d4 = d / 4;
    for (int chunk0 = 0; chunk0 < d4;) {
      int rem = d4 - chunk0;
      if (g_tune_spmm_slab == 32 || g_tune_spmm_slab == 64) rem = std::min(rem, g_tune_spmm_slab);
      int rc;
      if (rem <= 32) { rc = launch_vec<1>(rowptr, col, val, n_rows, H, d4, chunk0, ep, long_row_threshold, seg_begin, seg_end, n_seg, seg_scratch, s); chunk0 += 32; }
      else if (rem <= 64) { rc = launch_vec<2>(rowptr, col, val, n_rows, H, d4, chunk0, ep, long_row_threshold, seg_begin, seg_end, n_seg, seg_scratch, s); chunk0 += 64; }
      else if (rem <= 96) { rc = launch_vec<3>(rowptr, col, val, n_rows, H, d4, chunk0, ep, long_row_threshold, seg_begin, seg_end, n_seg, seg_scratch, s); chunk0 += 96; }
      else { rc = launch_vec<4>(rowptr, col, val, n_rows, H, d4, chunk0, ep, long_row_threshold, seg_begin, seg_end, n_seg, seg_scratch, s); chunk0 += 128; }
      if (rc != EG_OK) return rc;
    }
  } else {
    int64_t warps = n_rows + n_seg;
    spmm_scalar_kernel<<<(unsigned)ceil_div(warps, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
        rowptr, col, val, n_rows, H, d, ep, long_row_threshold, seg_begin, seg_end, n_seg, seg_scratch);
    EG_LAUNCHED();
  }
  if (n_seg > 0) {
    spmm_long_finish_kernel<<<(unsigned)n_long, 128, 0, s>>>(long_rows, long_first, n_long, seg_scratch, d, ep);
    EG_LAUNCHED();
  }
  return EG_OK;
}

int eg_epilogue_bwd(const float* dout, const float* a, const float* gate_pre, const float* x_res,
                    int64_t n_elem, int act, float* dS, float* d_gate, float* d_xres, eg_stream_t stream_) {
  using namespace eg;
  if (n_elem < 0 || !dout || !dS) return EG_ERR_INVALID;
  if (act == EG_ACT_RELU && !a) return EG_ERR_INVALID;
  if (gate_pre && (!x_res || !a)) return EG_ERR_INVALID;
  if (n_elem == 0) return EG_OK;
  const uintptr_t align = (uintptr_t)dout | (uintptr_t)a | (uintptr_t)gate_pre | (uintptr_t)x_res | (uintptr_t)dS |
                          (uintptr_t)d_gate | (uintptr_t)d_xres;
  if (n_elem % 4 == 0 && (align & 15) == 0) {
    const int64_t n4 = n_elem / 4;
    const int64_t blocks4 = std::min<int64_t>(ceil_div(n4, (int64_t)256), (int64_t)kNumSMs * 8);
    epilogue_bwd_vec4_kernel<<<(unsigned)blocks4, 256, 0, as_stream(stream_)>>>(
        reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(gate_pre),
        reinterpret_cast<const float4*>(x_res), n4, act, reinterpret_cast<float4*>(dS), reinterpret_cast<float4*>(d_gate),
        reinterpret_cast<float4*>(d_xres));
    EG_LAUNCHED();
    return EG_OK;
  }
  int64_t blocks = ceil_div(n_elem, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  epilogue_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(dout, a, gate_pre, x_res, n_elem, act,
                                                                        dS, d_gate, d_xres);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
