// (b) Log-domain Sinkhorn half-sweeps on a MATERIALISED cost matrix.
//
// Restates the scaling iteration of utils/ot_loss.py:53-55 (v = b / Kᵀu, u = a / Kv,
// K = exp(-M/reg)) and of SinkhornOT/sinkhorn_loss.py:197-201 with log u, log v:
//     log v_j = log b_j - LSE_i( log u_i - M_ij / reg )
//     log u_i = log a_i - LSE_j( log v_j - M_ij / reg )
// so K is never formed and nothing under/overflows.  Each half-sweep streams the
// cost once (HBM/L2-bound: I·J·sizeof(T) bytes); the column half-sweep runs on a
// transposed copy made once per solve so both directions read rows.
// fp32 and fp64 instantiations; the fp64 one reproduces the reference's float64
// arithmetic to rounding.
#include <math_constants.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sinkhorn_common.cuh"

namespace eg {

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
  static __device__ __forceinline__ float log_(float x) { return logf(x); }
  static __device__ __forceinline__ float ninf() { return -CUDART_INF_F; }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double ninf() { return -CUDART_INF; }
};

// Online (max, sum-of-exp) accumulator.
template <typename T>
struct Lse {
  T m, s;
  __device__ __forceinline__ void init() { m = Num<T>::ninf(); s = T(0); }
  // Branch-free (the solver kernels are instruction-issue bound; divergent `if (z > m)` costs more than one
  // extra exp): new max, rescale the running sum, add the four terms.  -inf inputs contribute exp(-inf) = 0;
  // an all -inf state keeps m = -inf, s = 0 (the reference point is clamped so that -inf - ref stays -inf).
  __device__ __forceinline__ void push4(T z0, T z1, T z2, T z3) {
    const T mx = fmax(fmax(fmax(z0, z1), fmax(z2, z3)), m);
    const T ref = (mx == Num<T>::ninf()) ? T(0) : mx;
    s = s * Num<T>::exp_(m - ref) +
        ((Num<T>::exp_(z0 - ref) + Num<T>::exp_(z1 - ref)) + (Num<T>::exp_(z2 - ref) + Num<T>::exp_(z3 - ref)));
    m = mx;
  }
  __device__ __forceinline__ void push1(T z) {
    const T mx = fmax(z, m);
    const T ref = (mx == Num<T>::ninf()) ? T(0) : mx;
    s = s * Num<T>::exp_(m - ref) + Num<T>::exp_(z - ref);
    m = mx;
  }
  __device__ __forceinline__ void merge(T om, T os) {
    const T mx = fmax(m, om);
    const T ref = (mx == Num<T>::ninf()) ? T(0) : mx;
    s = s * Num<T>::exp_(m - ref) + os * Num<T>::exp_(om - ref);
    m = mx;
  }
  __device__ __forceinline__ T value() const { return (s > T(0)) ? m + Num<T>::log_(s) : Num<T>::ninf(); }
};

// Warp all-reduce of (max, sum) pairs: max first (no transcendental), ONE rescale per lane, then a plain
// sum — 1 exp per lane instead of 2 per shuffle level.
template <typename T>
__device__ __forceinline__ void warp_merge(Lse<T>& a) {
  T mx = a.m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const T ref = (mx == Num<T>::ninf()) ? T(0) : mx;
  T sum = a.s * Num<T>::exp_(a.m - ref);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  a.m = mx;
  a.s = sum;
}

// One CTA of kRowThreads per row when rows are long, one warp per row otherwise.
template <typename T, int WARPS_PER_ROW>
__global__ void __launch_bounds__(256)
lse_rows_kernel(const T* __restrict__ M, int64_t n_rows, int64_t n_cols, int64_t ld, T inv_reg,
                const T* __restrict__ pot_in, const T* __restrict__ logw, T* __restrict__ pot_out,
                T* __restrict__ lse_out) {
  constexpr int kWarps = 8;
  constexpr int kRowsPerBlock = kWarps / WARPS_PER_ROW;
  __shared__ T sm_m[kWarps], sm_s[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = warp % WARPS_PER_ROW;          // which slice of the row this warp covers
  const int64_t row = (int64_t)blockIdx.x * kRowsPerBlock + warp / WARPS_PER_ROW;
  Lse<T> acc;
  acc.init();
  if (row < n_rows) {
    const T* mrow = M + row * ld;
    constexpr int V = 16 / sizeof(T);              // elements per 16-byte vector
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(mrow) | reinterpret_cast<uintptr_t>(pot_in)) & 15) == 0;
    const int64_t n_vec = vec_ok ? n_cols / V : 0;
    const int stride = 32 * WARPS_PER_ROW;
    if constexpr (sizeof(T) == 4) {
      const float4* m4 = reinterpret_cast<const float4*>(mrow);
      const float4* p4 = reinterpret_cast<const float4*>(pot_in);
      constexpr int U = 4;   // 16-byte loads in flight per lane before any math
      for (int64_t v0 = sub * 32 + lane; v0 < n_vec; v0 += (int64_t)stride * U) {
        float4 mv[U], pv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int64_t v = v0 + (int64_t)u * stride;
          if (v < n_vec) {
            mv[u] = ld_stream_f4(m4 + v);
            pv[u] = __ldg(p4 + v);
          } else {
            mv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            pv[u] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          acc.push4(fmaf(-mv[u].x, inv_reg, pv[u].x), fmaf(-mv[u].y, inv_reg, pv[u].y),
                    fmaf(-mv[u].z, inv_reg, pv[u].z), fmaf(-mv[u].w, inv_reg, pv[u].w));
      }
    } else {
      const double2* m2 = reinterpret_cast<const double2*>(mrow);
      const double2* p2 = reinterpret_cast<const double2*>(pot_in);
      constexpr int U = 4;
      for (int64_t v0 = sub * 32 + lane; v0 < n_vec; v0 += (int64_t)stride * U) {
        double2 mv[U], pv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int64_t v = v0 + (int64_t)u * stride;
          if (v < n_vec) { mv[u] = m2[v]; pv[u] = p2[v]; }
          else { mv[u] = make_double2(0.0, 0.0); pv[u] = make_double2(-CUDART_INF, -CUDART_INF); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          // pot - M/reg (division by reg folded into inv_reg)
          acc.push1(pv[u].x - mv[u].x * inv_reg);
          acc.push1(pv[u].y - mv[u].y * inv_reg);
        }
      }
    }
    for (int64_t j = n_vec * V + sub * 32 + lane; j < n_cols; j += stride)
      acc.push1(pot_in[j] - mrow[j] * inv_reg);
  }
  warp_merge(acc);
  if constexpr (WARPS_PER_ROW > 1) {
    if (lane == 0) { sm_m[warp] = acc.m; sm_s[warp] = acc.s; }
    __syncthreads();
    if (sub == 0 && lane == 0) {
      for (int w = 1; w < WARPS_PER_ROW; ++w) acc.merge(sm_m[warp + w], sm_s[warp + w]);
    }
  }
  if (row < n_rows && sub == 0 && lane == 0) {
    T l = acc.value();
    if (lse_out) lse_out[row] = l;
    if (pot_out) pot_out[row] = (logw ? logw[row] : T(0)) - l;
  }
}

template <typename T>
__global__ void transpose_kernel(const T* __restrict__ src, int64_t n_rows, int64_t n_cols, int64_t ld_src,
                                 T* __restrict__ dst, int64_t ld_dst) {
  __shared__ T tile[32][33];
  int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int64_t rr = r0 + r, cc = c0 + threadIdx.x;
    if (rr < n_rows && cc < n_cols) tile[r][threadIdx.x] = src[rr * ld_src + cc];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int64_t cc = c0 + r, rr = r0 + threadIdx.x;   // dst row = src col
    if (cc < n_cols && rr < n_rows) dst[cc * ld_dst + rr] = tile[threadIdx.x][r];
  }
}

// Plan statistics over a 32-row x 256-column tile; thread = column.
template <typename T>
__global__ void __launch_bounds__(256)
plan_kernel(const T* __restrict__ M, int64_t n_rows, int64_t n_cols, int64_t ld, T inv_reg,
            const T* __restrict__ f, const T* __restrict__ g, T* __restrict__ P, int64_t ldP,
            double* __restrict__ loss, T* __restrict__ row_sum, T* __restrict__ col_sum) {
  __shared__ double red[8];
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool ok = j < n_cols;
  const T gj = ok ? g[j] : T(0);
  const int rows = (int)min((int64_t)32, n_rows - i0);
  double my_loss = 0.0;
  T my_col = T(0);
  for (int r = 0; r < rows; ++r) {
    T p = T(0);
    if (ok) {
      T m = M[(i0 + r) * ld + j];
      p = Num<T>::exp_(f[i0 + r] + gj - m * inv_reg);
      if (P) P[(i0 + r) * ldP + j] = p;
      my_loss += (double)p * (double)m;
      my_col += p;
    }
    if (row_sum) {
      T rs = warp_sum(p);
      if (lane == 0 && rs != T(0)) atomicAdd(&row_sum[i0 + r], rs);
    }
  }
  if (col_sum && ok) atomicAdd(&col_sum[j], my_col);
  if (loss) {
    my_loss = warp_sum(my_loss);
    if (lane == 0) red[warp] = my_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      atomicAdd(loss, t);
    }
  }
}

template <typename T>
__global__ void log_kernel(const T* __restrict__ x, int64_t n, T* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = Num<T>::log_(x[i]);
}
template <typename T>
__global__ void fill_kernel(T* __restrict__ out, int64_t n, T value) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = value;
}

// err^2 = sum_j (exp(log_v_j + lse_j) - b_j)^2  — the column-marginal violation of
// utils/ot_loss.py:64-66, from quantities the next column half-sweep already produced.
template <typename T>
__global__ void marginal_err_kernel(const T* __restrict__ log_v, const T* __restrict__ col_lse,
                                    const T* __restrict__ b, int64_t n, double* __restrict__ err2) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    double m = (double)Num<T>::exp_(log_v[j] + col_lse[j]) - (double)b[j];
    acc += m * m;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(err2, t);
  }
}

template <typename T>
static int lse_dense_t(const T* M, int64_t n_rows, int64_t n_cols, int64_t ld, double inv_reg, const T* pot_in,
                       const T* logw, T* pot_out, T* lse_out, cudaStream_t s) {
  // pick warps-per-row so that >= ~3 warps/SM-slot-quarter are resident and each lane still
  // sees a few 16-byte vectors: rows*wpr >= 148*64 when the row is long enough
  const int64_t vec_per_row = n_cols / (16 / (int64_t)sizeof(T));
  int wpr = 1;
  while (wpr < 8 && n_rows * wpr < (int64_t)kNumSMs * 64 && vec_per_row / (32 * wpr * 2) >= 2) wpr *= 2;
  if (wpr == 1)
    lse_rows_kernel<T, 1><<<(unsigned)ceil_div(n_rows, 8), 256, 0, s>>>(M, n_rows, n_cols, ld, (T)inv_reg, pot_in,
                                                                         logw, pot_out, lse_out);
  else if (wpr == 2)
    lse_rows_kernel<T, 2><<<(unsigned)ceil_div(n_rows, 4), 256, 0, s>>>(M, n_rows, n_cols, ld, (T)inv_reg, pot_in,
                                                                         logw, pot_out, lse_out);
  else if (wpr == 4)
    lse_rows_kernel<T, 4><<<(unsigned)ceil_div(n_rows, 2), 256, 0, s>>>(M, n_rows, n_cols, ld, (T)inv_reg, pot_in,
                                                                         logw, pot_out, lse_out);
  else
    lse_rows_kernel<T, 8><<<(unsigned)n_rows, 256, 0, s>>>(M, n_rows, n_cols, ld, (T)inv_reg, pot_in, logw,
                                                           pot_out, lse_out);
  EG_LAUNCHED();
  return EG_OK;
}

template <typename T>
static int transpose_t(const T* src, int64_t n_rows, int64_t n_cols, int64_t ld_src, T* dst, int64_t ld_dst,
                       cudaStream_t s) {
  int64_t gy = ceil_div(n_rows, 32);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(n_cols, 32), (unsigned)gy), block(32, 8);
  transpose_kernel<T><<<grid, block, 0, s>>>(src, n_rows, n_cols, ld_src, dst, ld_dst);
  EG_LAUNCHED();
  return EG_OK;
}

template <typename T>
static int plan_t(const T* M, int64_t n_rows, int64_t n_cols, int64_t ld, double inv_reg, const T* f, const T* g,
                  T* P, int64_t ldP, double* loss, T* row_sum, T* col_sum, cudaStream_t s) {
  int64_t gy = ceil_div(n_rows, 32);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;
  if (loss) EG_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), s));
  if (row_sum) EG_CUDA(cudaMemsetAsync(row_sum, 0, sizeof(T) * (size_t)n_rows, s));
  if (col_sum) EG_CUDA(cudaMemsetAsync(col_sum, 0, sizeof(T) * (size_t)n_cols, s));
  dim3 grid((unsigned)ceil_div(n_cols, 256), (unsigned)gy);
  plan_kernel<T><<<grid, 256, 0, s>>>(M, n_rows, n_cols, ld, (T)inv_reg, f, g, P, ldP, loss, row_sum, col_sum);
  EG_LAUNCHED();
  return EG_OK;
}


// ---- whole solve in ONE persistent cooperative kernel -----------------------------------------
// The streaming path above costs a launch (and a cold start) per half-sweep: 2000 launches for the
// reference's default 1000 sweeps.  Here one CTA per SM owns a contiguous block of rows for the whole
// solve; per sweep
//   C: partial column log-sum-exp over the CTA's rows           (thread = 16-byte column group)
//   -- grid barrier --
//   R: CTA c merges the partials of its slice of columns -> log v (+ marginal error on check sweeps)
//   -- grid barrier --
//   U: row update for the CTA's rows against the new log v      (warp = row)
// so M is read twice per sweep (from L2 when it fits), nothing else leaves the SM, and the only
// global synchronisation is two barriers per sweep.  Same init, order and stopping rule as
// utils/ot_loss.py:38-70.

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblocks) {
  __syncthreads();   // all of this CTA's writes are ordered before thread 0's release below (CTA-scope barrier)
  if (threadIdx.x == 0) {
    target += nblocks;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };
__device__ __forceinline__ void unpack(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }

constexpr int kPersistThreads = 1024;

template <typename T>
__global__ void __launch_bounds__(kPersistThreads, 1)
sinkhorn_persistent_kernel(const T* __restrict__ M, int64_t I, int64_t J, int64_t ld, T inv_reg,
                           const T* __restrict__ log_a, const T* __restrict__ log_b, const T* __restrict__ b,
                           int max_iter, double stop_thr, T* __restrict__ part_m, T* __restrict__ part_s,
                           T* __restrict__ lv_buf /* [2][J] */, T* __restrict__ log_u_out, T* __restrict__ log_v_out,
                           PersistState* __restrict__ st, int rows_resident) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::N;
  constexpr int kWarps = kPersistThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw_p[];
  const int nb = gridDim.x, cta = blockIdx.x;
  const int64_t rows_per = (I + nb - 1) / nb;
  const int64_t row0 = min(I, (int64_t)cta * rows_per);
  const int R = (int)(min(I, row0 + rows_per) - row0);
  const int64_t cols_per = (J + nb - 1) / nb;
  const int64_t col0 = min(J, (int64_t)cta * cols_per), col1 = min(J, col0 + cols_per);
  // shared memory: log v [J] | my log u [rows_per] | cross-warp merge scratch [2][kWarps][32] | resident rows of M
  T* lv_s = reinterpret_cast<T*>(smem_raw_p);
  T* lu_s = lv_s + J;
  T* red_m = lu_s + ((rows_per + 3) / 4) * 4;
  T* red_s = red_m + kWarps * 32;
  Vec* m_res = reinterpret_cast<Vec*>(red_s + kWarps * 32);          // [S][J / V], 16-byte aligned by construction
  const int S = min(R, rows_resident);
  __shared__ double err_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t groups = J / V;                                       // J % V == 0 is a launch precondition
  const int64_t ldv = ld / V;
  unsigned int target = 0;

  for (int r = tid; r < R; r += kPersistThreads) lu_s[r] = (T)(-log((double)I));
  // rows [0, S) of my block stay in shared memory for the whole solve; the rest stream from L2 every phase
  for (int r = 0; r < S; ++r) {
    const Vec* src = reinterpret_cast<const Vec*>(M + (row0 + r) * ld);
    for (int64_t g = tid; g < groups; g += kPersistThreads) m_res[(int64_t)r * groups + g] = src[g];
  }
  __syncthreads();

  int cpt = 0, sweeps = 0, final_buf = 0;
  double err = 1.0;
  bool stopped = false;
  const bool timer = (cta == 0 && tid == 0);
  for (cpt = 0; cpt < max_iter; ++cpt) {
    const int cur = cpt & 1, nxt = cur ^ 1;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (timer) t0 = gtime();
    // ---- C: partial column LSE over my rows (thread = 16-byte column group, V independent chains) -----
    const int ngroups = (int)groups;
    for (int g = tid; g < ngroups; g += kPersistThreads) {
      Lse<T> acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v].init();
      const Vec* gbase = reinterpret_cast<const Vec*>(M + row0 * ld) + g;
      const Vec* sbase = m_res + g;
      for (int r0 = 0; r0 < R; r0 += 4) {
        Vec mv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u;
          mv[u] = (r < S) ? sbase[r * ngroups] : (r < R ? gbase[(int64_t)r * ldv] : Vec());
        }
        T z[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const T lu = (r0 + u < R) ? lu_s[r0 + u] : Num<T>::ninf();
          T e[V];
          unpack(mv[u], e);
#pragma unroll
          for (int v = 0; v < V; ++v) z[u][v] = lu - e[v] * inv_reg;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v].push4(z[0][v], z[1][v], z[2][v], z[3][v]);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        part_m[(int64_t)cta * J + (int64_t)g * V + v] = acc[v].m;
        part_s[(int64_t)cta * J + (int64_t)g * V + v] = acc[v].s;
      }
    }
    if (timer) t1 = gtime();
    grid_barrier(&st->barrier, target, nb);
    if (timer) t2 = gtime();
    // ---- R: merge all CTAs' partials for my slice of columns -> log v --------------------------------
    // lane = column inside the slice (contiguous floats: coalesced), warp = strided subset of source CTAs
    const bool check = (cpt >= 1) && ((cpt - 1) % 10 == 0);
    const int slot = check ? ((cpt - 1) / 10) & 127 : 0;
    if (tid == 0) err_s = 0.0;
    for (int64_t c0 = col0; c0 < col1; c0 += 32) {
      const int64_t j = c0 + lane;
      Lse<T> acc;
      acc.init();
      if (j < col1) {
        // values written by other SMs since this SM last read them: read at L2, not a stale L1 line
        T pm[5], ps[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int pidx = warp + kWarps * k;
          pm[k] = (pidx < nb) ? __ldcg(part_m + (int64_t)pidx * J + j) : Num<T>::ninf();
          ps[k] = (pidx < nb) ? __ldcg(part_s + (int64_t)pidx * J + j) : T(0);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) acc.merge(pm[k], ps[k]);
        for (int pidx = warp + kWarps * 5; pidx < nb; pidx += kWarps)
          acc.merge(__ldcg(part_m + (int64_t)pidx * J + j), __ldcg(part_s + (int64_t)pidx * J + j));
      }
      red_m[warp * 32 + lane] = acc.m;
      red_s[warp * 32 + lane] = acc.s;
      __syncthreads();
      {
        // warp w finishes column c0 + w: lanes hold the 32 per-warp partials, 5 shuffle-merge steps
        const int64_t jc = c0 + warp;
        Lse<T> tot;
        tot.m = red_m[lane * 32 + warp];
        tot.s = red_s[lane * 32 + warp];
        warp_merge(tot);
        if (lane == 0 && jc < col1) {
          const T lse = tot.value();
          if (check) {
            const double d = (double)Num<T>::exp_(__ldcg(lv_buf + (int64_t)cur * J + jc) + lse) - (double)b[jc];
            atomicAdd(&err_s, d * d);
          }
          lv_buf[(int64_t)nxt * J + jc] = log_b[jc] - lse;
        }
      }
      __syncthreads();
    }
    if (check && tid == 0 && err_s != 0.0) atomicAdd(&st->err2[slot], err_s);
    if (timer) t3 = gtime();
    grid_barrier(&st->barrier, target, nb);
    if (timer) t4 = gtime();
    if (check) {
      double e2;
      asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(e2) : "l"(&st->err2[slot]) : "memory");
      err = sqrt(e2);
      if (!(err > stop_thr)) { stopped = true; final_buf = cur; break; }   // sweep cpt never happened
    }
    // ---- U: row update for my rows with the new log v (warp = row, 4 independent chains per lane) --------
    {
      const T* src = lv_buf + (int64_t)nxt * J;
      for (int64_t j0 = tid; j0 < J; j0 += (int64_t)kPersistThreads * 4) {
        T val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t j = j0 + (int64_t)u * kPersistThreads;
          val[u] = (j < J) ? __ldcg(src + j) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t j = j0 + (int64_t)u * kPersistThreads;
          if (j < J) lv_s[j] = val[u];
        }
      }
    }
    __syncthreads();
    {
      const Vec* pv = reinterpret_cast<const Vec*>(lv_s);
      for (int r = warp; r < R; r += kWarps) {
        const Vec* mrow = (r < S) ? (m_res + (int64_t)r * groups) : reinterpret_cast<const Vec*>(M + (row0 + r) * ld);
        Lse<T> acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u].init();
        const int ng = (int)groups;
        for (int g0 = lane; g0 < ng; g0 += 32 * 4) {
          Vec mv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int g = g0 + 32 * u;
            mv[u] = (g < ng) ? mrow[g] : Vec();
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int g = g0 + 32 * u;
            if (g < ng) {
              T e[V], l[V];
              unpack(mv[u], e);
              unpack(pv[g], l);
              if constexpr (V == 4)
                acc[u].push4(l[0] - e[0] * inv_reg, l[1] - e[1] * inv_reg, l[2] - e[2] * inv_reg, l[3] - e[3] * inv_reg);
              else {
                acc[u].push1(l[0] - e[0] * inv_reg);
                acc[u].push1(l[1] - e[1] * inv_reg);
              }
            }
          }
        }
        acc[0].merge(acc[1].m, acc[1].s);
        acc[2].merge(acc[3].m, acc[3].s);
        acc[0].merge(acc[2].m, acc[2].s);
        warp_merge(acc[0]);
        if (lane == 0) lu_s[r] = log_a[row0 + r] - acc[0].value();
      }
    }
    __syncthreads();
    sweeps = cpt + 1;
    if (timer) {
      st->t_phase[0] += t1 - t0; st->t_phase[1] += t2 - t1; st->t_phase[2] += t3 - t2;
      st->t_phase[3] += t4 - t3; st->t_phase[4] += gtime() - t4;
    }
  }
  if (!stopped) final_buf = max_iter & 1;
  // ---- outputs ---------------------------------------------------------------------------------
  for (int r = tid; r < R; r += kPersistThreads) log_u_out[row0 + r] = lu_s[r];
  for (int64_t j = col0 + tid; j < col1; j += kPersistThreads) log_v_out[j] = __ldcg(lv_buf + (int64_t)final_buf * J + j);
  if (cta == 0 && tid == 0) { st->sweeps = sweeps; st->final_buf = final_buf; st->err = err; }
}


// ---- fully on-chip variant for the reference's batch size (3000 x 3000 fp32, `config.py:30`) -----------
// Same sweep structure as sinkhorn_persistent_kernel, but the CTA's whole row block stays on the SM:
// S rows in shared memory + up to kOcRR rows in registers (thread t keeps its own 16-byte column groups of
// those rows), so phases C and U never wait on L2.  512 threads (128 registers each), 2 column groups per
// thread => J <= 4096.
constexpr int kOcThreads = 512;
constexpr int kOcWarps = kOcThreads / 32;
constexpr int kOcQ = 2;     // column groups (float4) per thread
constexpr int kOcRR = 5;    // rows kept in registers (phase C is written for exactly 4 + 1)
static_assert(kOcRR == 5, "phase C unrolls the register rows as 4 + 1");

// (max, sum) pair in log2 units; merge = one ex2 per operand.
struct Lse2 {
  float m, s;
  __device__ __forceinline__ void init() { m = -CUDART_INF_F; s = 0.f; }
  __device__ __forceinline__ void merge(float om, float os) {
    const float mx = fmaxf(m, om);
    const float ref = (mx == -CUDART_INF_F) ? 0.f : mx;
    s = s * ex2f(m - ref) + os * ex2f(om - ref);
    m = mx;
  }
  __device__ __forceinline__ float value() const { return (s > 0.f) ? m + lg2f(s) : -CUDART_INF_F; }
};
__device__ __forceinline__ void warp_merge2(Lse2& a) {
  float mx = a.m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float ref = (mx == -CUDART_INF_F) ? 0.f : mx;
  float sum = a.s * ex2f(a.m - ref);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  a.m = mx;
  a.s = sum;
}

__device__ int g_dev_sinkhorn_redos = 0;

__global__ void __launch_bounds__(kOcThreads, 1)
sinkhorn_onchip_kernel(const float* __restrict__ M, int64_t I, int J, int64_t ld, float inv_reg,
                       const float* __restrict__ log_a, const float* __restrict__ log_b, const float* __restrict__ b,
                       int max_iter, double stop_thr, float* __restrict__ part_m, float* __restrict__ part_s,
                       float* __restrict__ lv_buf, float* __restrict__ log_u_out, float* __restrict__ log_v_out,
                       PersistState* __restrict__ st, int S_max, const int* __restrict__ run_if) {
  using T = float;
  extern __shared__ __align__(16) unsigned char smem_raw_o[];
  const int nb = gridDim.x, cta = blockIdx.x;
  // conditional launch (the redo after a scaling-domain solve that left the fp32 range): the flag was written by
  // an earlier kernel on the same stream, every CTA reads the same value
  if (run_if != nullptr) {
    if (*run_if == 0) return;
    if (cta == 0 && threadIdx.x == 0) atomicAdd(&g_dev_sinkhorn_redos, 1);
  }
  const int rows_per = (int)((I + nb - 1) / nb);
  const int64_t row0 = min(I, (int64_t)cta * rows_per);
  const int R = (int)(min(I, row0 + rows_per) - row0);
  const int S = min(R, S_max);                              // rows [0,S) in smem, [S,R) in registers (R-S <= kOcRR)
  const int cols_per = (J + nb - 1) / nb;
  const int col0 = min(J, cta * cols_per), col1 = min(J, col0 + cols_per);
  const int ng = J / 4;
  float* lv_s = reinterpret_cast<float*>(smem_raw_o);      // [J]
  float* lu_s = lv_s + J;                                   // [32]
  float* red_m = lu_s + 32;                                 // [kOcWarps][32]
  float* red_s = red_m + kOcWarps * 32;
  float* rr_m = red_s + kOcWarps * 32;                      // [kOcWarps][kOcRR]
  float* rr_s = rr_m + kOcWarps * 8;
  float4* m_res = reinterpret_cast<float4*>(rr_s + kOcWarps * 8);   // [S][ng]
  __shared__ double err_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned int target = 0;
  const float inv2 = inv_reg * kLog2e;       // all potentials inside this kernel are in log2 units

  int gq[kOcQ];
  bool gok[kOcQ];
#pragma unroll
  for (int q = 0; q < kOcQ; ++q) { gq[q] = tid + kOcThreads * q; gok[q] = gq[q] < ng; }
  float4 mreg[kOcRR][kOcQ];
#pragma unroll
  for (int rr = 0; rr < kOcRR; ++rr)
#pragma unroll
    for (int q = 0; q < kOcQ; ++q) {
      const bool live = gok[q] && (S + rr < R);
      mreg[rr][q] = live ? reinterpret_cast<const float4*>(M + (row0 + S + rr) * ld)[gq[q]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  for (int r = 0; r < S; ++r) {
    const float4* src = reinterpret_cast<const float4*>(M + (row0 + r) * ld);
    for (int g = tid; g < ng; g += kOcThreads) m_res[r * ng + g] = src[g];
  }
  for (int r = tid; r < 32; r += kOcThreads) lu_s[r] = (r < R) ? (float)(-log2((double)I)) : -CUDART_INF_F;
  __syncthreads();

  int cpt = 0, sweeps = 0, final_buf = 0;
  double err = 1.0;
  bool stopped = false;
  const bool timer = (cta == 0 && tid == 0);
  for (cpt = 0; cpt < max_iter; ++cpt) {
    const int cur = cpt & 1, nxt = cur ^ 1;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (timer) t0 = gtime();
    // ---- C: partial column LSE over my rows (two passes over on-chip data: max, then one exp2 each) ----
    // everything in log2 units: z2 = (log u_i - M_ij/reg) * log2(e)
#pragma unroll
    for (int q = 0; q < kOcQ; ++q) {
      if (!gok[q]) continue;
      const float4* sbase = m_res + gq[q];
      // ONE pass over the on-chip rows, 4 rows per step per column: running max, one rescale, 4 ex2
      float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      auto step4 = [&](const float (&z)[4][4]) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float m4 = fmaxf(fmaxf(fmaxf(z[0][v], z[1][v]), fmaxf(z[2][v], z[3][v])), mx[v]);
          const float ref = (m4 == -CUDART_INF_F) ? 0.f : m4;
          sum[v] = sum[v] * ex2f(mx[v] - ref) +
                   ((ex2f(z[0][v] - ref) + ex2f(z[1][v] - ref)) + (ex2f(z[2][v] - ref) + ex2f(z[3][v] - ref)));
          mx[v] = m4;
        }
      };
      for (int r0 = 0; r0 < S; r0 += 4) {
        float z[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u;
          const float4 mv = (r < S) ? sbase[r * ng] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float lu = (r < S) ? lu_s[r] : -CUDART_INF_F;
          z[u][0] = fmaf(-mv.x, inv2, lu); z[u][1] = fmaf(-mv.y, inv2, lu);
          z[u][2] = fmaf(-mv.z, inv2, lu); z[u][3] = fmaf(-mv.w, inv2, lu);
        }
        step4(z);
      }
      {
        float z[4][4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float lu = lu_s[S + rr];                    // -inf beyond R (lu_s is padded to 32 entries)
          z[rr][0] = fmaf(-mreg[rr][q].x, inv2, lu); z[rr][1] = fmaf(-mreg[rr][q].y, inv2, lu);
          z[rr][2] = fmaf(-mreg[rr][q].z, inv2, lu); z[rr][3] = fmaf(-mreg[rr][q].w, inv2, lu);
        }
        step4(z);
        const float lu = lu_s[S + 4];
        const float z4[4] = {fmaf(-mreg[4][q].x, inv2, lu), fmaf(-mreg[4][q].y, inv2, lu),
                             fmaf(-mreg[4][q].z, inv2, lu), fmaf(-mreg[4][q].w, inv2, lu)};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float m4 = fmaxf(z4[v], mx[v]);
          const float ref = (m4 == -CUDART_INF_F) ? 0.f : m4;
          sum[v] = sum[v] * ex2f(mx[v] - ref) + ex2f(z4[v] - ref);
          mx[v] = m4;
        }
      }
      reinterpret_cast<float4*>(part_m + (int64_t)cta * J)[gq[q]] = make_float4(mx[0], mx[1], mx[2], mx[3]);
      reinterpret_cast<float4*>(part_s + (int64_t)cta * J)[gq[q]] = make_float4(sum[0], sum[1], sum[2], sum[3]);
    }
    if (timer) t1 = gtime();
    grid_barrier(&st->barrier, target, nb);
    if (timer) t2 = gtime();
    // ---- R: merge all CTAs' partials for my slice of columns -> log v ---------------------------------
    const bool check = (cpt >= 1) && ((cpt - 1) % 10 == 0);
    const int slot = check ? ((cpt - 1) / 10) & 127 : 0;
    if (tid == 0) err_s = 0.0;
    for (int c0 = col0; c0 < col1; c0 += 32) {
      const int j = c0 + lane;
      Lse2 acc;
      acc.init();
      if (j < col1) {
        float pm[10], ps[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const int pidx = warp + kOcWarps * k;
          pm[k] = (pidx < nb) ? __ldcg(part_m + (int64_t)pidx * J + j) : -CUDART_INF_F;
          ps[k] = (pidx < nb) ? __ldcg(part_s + (int64_t)pidx * J + j) : 0.f;
        }
        // max first, then one ex2 per partial (no dependent merge chain)
        float mx = pm[0];
#pragma unroll
        for (int k = 1; k < 10; ++k) mx = fmaxf(mx, pm[k]);
        const float ref = (mx == -CUDART_INF_F) ? 0.f : mx;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 10; ++k) sum += ps[k] * ex2f(pm[k] - ref);
        acc.m = mx;
        acc.s = sum;
      }
      red_m[warp * 32 + lane] = acc.m;
      red_s[warp * 32 + lane] = acc.s;
      __syncthreads();
      {
        // warp w (and w + 16) finishes column c0 + w: lanes 0..15 hold the per-warp partials
        for (int cw = warp; cw < 32; cw += kOcWarps) {
          const int jc = c0 + cw;
          Lse2 tot;
          tot.m = (lane < kOcWarps) ? red_m[lane * 32 + cw] : -CUDART_INF_F;
          tot.s = (lane < kOcWarps) ? red_s[lane * 32 + cw] : 0.f;
          warp_merge2(tot);
          if (lane == 0 && jc < col1) {
            const float lse2 = tot.value();
            if (check) {
              const double d = (double)ex2f(__ldcg(lv_buf + (int64_t)cur * J + jc) + lse2) - (double)b[jc];
              atomicAdd(&err_s, d * d);
            }
            lv_buf[(int64_t)nxt * J + jc] = log_b[jc] * kLog2e - lse2;
          }
        }
      }
      __syncthreads();
    }
    if (check && tid == 0 && err_s != 0.0) atomicAdd(&st->err2[slot], err_s);
    if (timer) t3 = gtime();
    grid_barrier(&st->barrier, target, nb);
    if (timer) t4 = gtime();
    if (check) {
      double e2;
      asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(e2) : "l"(&st->err2[slot]) : "memory");
      err = sqrt(e2);
      if (!(err > stop_thr)) { stopped = true; final_buf = cur; break; }
    }
    // ---- U: row update for my rows with the new log v ---------------------------------------------------
    {
      const float4* src = reinterpret_cast<const float4*>(lv_buf + (int64_t)nxt * J);
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (gok[0]) v0 = __ldcg(src + gq[0]);
      if (gok[1]) v1 = __ldcg(src + gq[1]);
      if (timer) st->t_phase[5] += gtime() - t4;
      if (gok[0]) reinterpret_cast<float4*>(lv_s)[gq[0]] = v0;
      if (gok[1]) reinterpret_cast<float4*>(lv_s)[gq[1]] = v1;
      // register rows: this thread's columns of rows S..R-1 (it already holds the matching log v values)
#pragma unroll
      for (int rr = 0; rr < kOcRR; ++rr) {
        float z[8];
        z[0] = gok[0] ? fmaf(-mreg[rr][0].x, inv2, v0.x) : -CUDART_INF_F;
        z[1] = gok[0] ? fmaf(-mreg[rr][0].y, inv2, v0.y) : -CUDART_INF_F;
        z[2] = gok[0] ? fmaf(-mreg[rr][0].z, inv2, v0.z) : -CUDART_INF_F;
        z[3] = gok[0] ? fmaf(-mreg[rr][0].w, inv2, v0.w) : -CUDART_INF_F;
        z[4] = gok[1] ? fmaf(-mreg[rr][1].x, inv2, v1.x) : -CUDART_INF_F;
        z[5] = gok[1] ? fmaf(-mreg[rr][1].y, inv2, v1.y) : -CUDART_INF_F;
        z[6] = gok[1] ? fmaf(-mreg[rr][1].z, inv2, v1.z) : -CUDART_INF_F;
        z[7] = gok[1] ? fmaf(-mreg[rr][1].w, inv2, v1.w) : -CUDART_INF_F;
        float mx = fmaxf(fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), fmaxf(fmaxf(z[4], z[5]), fmaxf(z[6], z[7])));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float ref = (mx == -CUDART_INF_F) ? 0.f : mx;
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += ex2f(z[e] - ref);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) { rr_m[warp * 8 + rr] = mx; rr_s[warp * 8 + rr] = sum; }
      }
    }
    if (timer) st->t_phase[6] += gtime() - t4;
    __syncthreads();
    {
      const float4* pv = reinterpret_cast<const float4*>(lv_s);
      for (int r = warp; r < S; r += kOcWarps) {
        const float4* mrow = m_res + r * ng;
        // ONE pass (shared-memory bandwidth is what binds here: every warp streams its row and the whole log v):
        // 16 elements per step, one rescale of the running sum per step, then 16 ex2.
        float run_m = -CUDART_INF_F, run_s = 0.f;
        for (int g0 = lane; g0 < ng; g0 += 128) {
          float z[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int g = g0 + 32 * u;
            if (g < ng) {
              const float4 e = mrow[g], l = pv[g];
              z[4 * u + 0] = fmaf(-e.x, inv2, l.x); z[4 * u + 1] = fmaf(-e.y, inv2, l.y);
              z[4 * u + 2] = fmaf(-e.z, inv2, l.z); z[4 * u + 3] = fmaf(-e.w, inv2, l.w);
            } else {
              z[4 * u + 0] = z[4 * u + 1] = z[4 * u + 2] = z[4 * u + 3] = -CUDART_INF_F;
            }
          }
          float mxl = run_m;
#pragma unroll
          for (int e = 0; e < 16; ++e) mxl = fmaxf(mxl, z[e]);
          const float refl = (mxl == -CUDART_INF_F) ? 0.f : mxl;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            a0 += ex2f(z[e] - refl); a1 += ex2f(z[e + 1] - refl); a2 += ex2f(z[e + 2] - refl); a3 += ex2f(z[e + 3] - refl);
          }
          run_s = run_s * ex2f(run_m - refl) + ((a0 + a1) + (a2 + a3));
          run_m = mxl;
        }
        float mx = run_m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float ref = (mx == -CUDART_INF_F) ? 0.f : mx;
        const float s0 = run_s * ex2f(run_m - ref), s1 = 0.f;
        float sum = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) lu_s[r] = log_a[row0 + r] * kLog2e - (mx + lg2f(sum));
      }
      if (timer) st->t_phase[7] += gtime() - t4;
      if (warp == kOcWarps - 1) {                            // finish the register rows: 16 per-warp partials each
#pragma unroll
        for (int rr = 0; rr < kOcRR; ++rr) {
          Lse2 tot;
          tot.m = (lane < kOcWarps) ? rr_m[lane * 8 + rr] : -CUDART_INF_F;
          tot.s = (lane < kOcWarps) ? rr_s[lane * 8 + rr] : 0.f;
          warp_merge2(tot);
          if (lane == 0 && S + rr < R) lu_s[S + rr] = log_a[row0 + S + rr] * kLog2e - tot.value();
        }
      }
    }
    __syncthreads();
    sweeps = cpt + 1;
    if (timer) {
      st->t_phase[0] += t1 - t0; st->t_phase[1] += t2 - t1; st->t_phase[2] += t3 - t2;
      st->t_phase[3] += t4 - t3; st->t_phase[4] += gtime() - t4;
    }
  }
  if (!stopped) final_buf = max_iter & 1;
  for (int r = tid; r < R; r += kOcThreads) log_u_out[row0 + r] = lu_s[r] * kLn2;
  for (int j = col0 + tid; j < col1; j += kOcThreads) log_v_out[j] = __ldcg(lv_buf + (int64_t)final_buf * J + j) * kLn2;
  if (cta == 0 && tid == 0) { st->sweeps = sweeps; st->final_buf = final_buf; st->err = err; }
}

// ---- scaling-domain continuation of the on-chip solve ---------------------------------------------------
// After a few log-domain sweeps (sinkhorn_onchip_kernel) the potentials are close enough that the classic
// Sinkhorn-Knopp scaling form is safe in fp32 *relative to them*: with Kt_ij = exp2(LU_i + LV_j - M_ij/reg)
// held on chip (<= a_i, rows sum to a_i at the hand-over) the sweep becomes two mat-vecs
//     v_j = b_j / sum_i Kt_ij u_i,      u_i = a_i / sum_j Kt_ij v_j
// with NO exponential: one FMA per entry per pass instead of FMA + MUFU.ex2, one float of partials per column
// instead of a (max, sum) pair.  The accepted iterate is (LU + log2 u, LV + log2 v), identical in exact
// arithmetic to the log-domain recursion, same stop rule.  Safety: if |log2 u| or |log2 v| drifts past 32 every
// CTA folds u, v into Kt and the potentials (absorb_req); if a sum is 0 / inf / nan the kernel raises `fallback`
// and the host redoes the solve with the log-domain kernel.  All reductions run in a fixed order
// (deterministic).

// 8 values per lane -> lane l ends with the warp-wide sum of value (l & 7): 7 + 2 shuffles instead of 8 x 5
__device__ __forceinline__ float warp_transpose_sum8(float (&v)[8], int lane) {
#pragma unroll
  for (int w = 4; w >= 1; w >>= 1) {
    const bool hi = (lane & w) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = hi ? v[k + w] : v[k];
      const float send = hi ? v[k] : v[k + w];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  float t = v[0];
  t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t;
}

constexpr int kScSlots = 40;   // per-row slots: [0, 8) register rows (RR used), [8, 40) shared-memory rows

template <int RR>   // rows kept in registers (the rest of the CTA's row block lives in shared memory)
__global__ void __launch_bounds__(kOcThreads, 1)
sinkhorn_onchip_scaling_kernel(const float* __restrict__ M, int64_t I, int J, int64_t ld, float inv_reg,
                               const float* __restrict__ a, const float* __restrict__ b,
                               const float* __restrict__ log_u_in, const float* __restrict__ log_v_in,
                               int start_iter, int max_iter, double stop_thr, float* __restrict__ part,
                               float* __restrict__ v_buf, float* __restrict__ log_u_out,
                               float* __restrict__ log_v_out, PersistState* __restrict__ st, int S_max,
                               float kAbsorbLog2) {
  extern __shared__ __align__(16) unsigned char smem_raw_o[];
  const int nb = gridDim.x, cta = blockIdx.x;
  const int rows_per = (int)((I + nb - 1) / nb);
  const int64_t row0 = min(I, (int64_t)cta * rows_per);
  const int R = (int)(min(I, row0 + rows_per) - row0);
  const int S = min(R, S_max);                              // rows [0,S) in smem (S <= 32), [S,R) in registers
  const int cols_per = (J + nb - 1) / nb;                   // <= 32 (host checks)
  const int col0 = min(J, cta * cols_per), col1 = min(J, col0 + cols_per);
  const int ncols = col1 - col0;
  const int ng = J / 4;
  float* u_s = reinterpret_cast<float*>(smem_raw_o);        // [40] current scaling, by slot
  float* LU_s = u_s + kScSlots;                             // [40] log2 potential already folded into Kt, by slot
  float* a_s = LU_s + kScSlots;                             // [40] row marginals, by slot
  float* red = a_s + kScSlots;                              // [kOcWarps][40]
  int* flag_s = reinterpret_cast<int*>(red + kOcWarps * kScSlots);   // [4]
  float4* kt = reinterpret_cast<float4*>(flag_s + 4);       // [S][ng]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned int target = 0;
  const float inv2 = inv_reg * kLog2e;

  int gq[kOcQ];
  bool gok[kOcQ];
#pragma unroll
  for (int q = 0; q < kOcQ; ++q) { gq[q] = tid + kOcThreads * q; gok[q] = gq[q] < ng; }
  // slot -> local row: slot < 8: register row S + slot (live iff slot < RR and S + slot < R); else smem row slot - 8
  const int my_row = (tid < 8) ? S + tid : tid - 8;
  const bool my_live = (tid < kScSlots) && ((tid < 8) ? (tid < RR && my_row < R) : (my_row < S));
  if (tid < kScSlots) {
    u_s[tid] = 1.0f;
    LU_s[tid] = my_live ? log_u_in[row0 + my_row] * kLog2e : 0.f;
    a_s[tid] = my_live ? a[row0 + my_row] : 0.f;
  }
  if (tid < 4) flag_s[tid] = 0;
  __syncthreads();
  {
    float4 lvq[kOcQ];
#pragma unroll
    for (int q = 0; q < kOcQ; ++q) {
      lvq[q] = gok[q] ? reinterpret_cast<const float4*>(log_v_in)[gq[q]] : make_float4(0.f, 0.f, 0.f, 0.f);
      lvq[q].x *= kLog2e; lvq[q].y *= kLog2e; lvq[q].z *= kLog2e; lvq[q].w *= kLog2e;
    }
    for (int r = 0; r < S; ++r) {
      const float4* src = reinterpret_cast<const float4*>(M + (row0 + r) * ld);
      const float lu = LU_s[8 + r];
#pragma unroll
      for (int q = 0; q < kOcQ; ++q)
        if (gok[q]) {
          const float4 m = src[gq[q]];
          kt[r * ng + gq[q]] = make_float4(ex2f(fmaf(-m.x, inv2, lu + lvq[q].x)), ex2f(fmaf(-m.y, inv2, lu + lvq[q].y)),
                                           ex2f(fmaf(-m.z, inv2, lu + lvq[q].z)), ex2f(fmaf(-m.w, inv2, lu + lvq[q].w)));
        }
    }
  }
  float4 kreg[RR][kOcQ];
#pragma unroll
  for (int rr = 0; rr < RR; ++rr)
#pragma unroll
    for (int q = 0; q < kOcQ; ++q) {
      kreg[rr][q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gok[q] && (S + rr < R)) {
        const float4 m = reinterpret_cast<const float4*>(M + (row0 + S + rr) * ld)[gq[q]];
        float4 lv = reinterpret_cast<const float4*>(log_v_in)[gq[q]];
        const float lu = LU_s[rr];
        kreg[rr][q] = make_float4(ex2f(fmaf(-m.x, inv2, lu + lv.x * kLog2e)), ex2f(fmaf(-m.y, inv2, lu + lv.y * kLog2e)),
                                  ex2f(fmaf(-m.z, inv2, lu + lv.z * kLog2e)), ex2f(fmaf(-m.w, inv2, lu + lv.w * kLog2e)));
      }
    }
  // owned columns: warp 0, lane l < ncols
  float LV_own = 0.f, v_own = 1.0f, b_own = 0.f;
  double LU_fold = 0.0, LV_fold = 0.0;                      // what the fold steps have moved into Kt since the hand-over
  if (warp == 0 && lane < ncols) { LV_own = log_v_in[col0 + lane] * kLog2e; b_own = b[col0 + lane]; }
  __syncthreads();

  int cpt = start_iter, sweeps = start_iter;
  double err = 1.0;
  // raised in phase U, published in the next phase R (between the same two grid barriers as every other flag
  // write, so that all CTAs read one value)
  bool u_big = false, u_bad = false;
  const bool timer = (cta == 0 && tid == 0);
  const int gqs0 = gok[0] ? gq[0] : 0, gqs1 = gok[1] ? gq[1] : 0;
  for (cpt = start_iter; cpt < max_iter; ++cpt) {
    const int nxt = (cpt & 1) ^ 1;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (timer) t0 = gtime();
    // ---- C: partial column sums over my rows ---------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < kOcQ; ++q) {
      if (!gok[q]) continue;
      const float4* sbase = kt + gq[q];
      float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
      int r = 0;
      for (; r + 3 < S; r += 4) {
        const float4 k0 = sbase[r * ng], k1 = sbase[(r + 1) * ng], k2 = sbase[(r + 2) * ng], k3 = sbase[(r + 3) * ng];
        const float u0 = u_s[8 + r], u1 = u_s[9 + r], u2 = u_s[10 + r], u3 = u_s[11 + r];
        acc0.x = fmaf(k0.x, u0, acc0.x); acc0.y = fmaf(k0.y, u0, acc0.y); acc0.z = fmaf(k0.z, u0, acc0.z); acc0.w = fmaf(k0.w, u0, acc0.w);
        acc1.x = fmaf(k1.x, u1, acc1.x); acc1.y = fmaf(k1.y, u1, acc1.y); acc1.z = fmaf(k1.z, u1, acc1.z); acc1.w = fmaf(k1.w, u1, acc1.w);
        acc0.x = fmaf(k2.x, u2, acc0.x); acc0.y = fmaf(k2.y, u2, acc0.y); acc0.z = fmaf(k2.z, u2, acc0.z); acc0.w = fmaf(k2.w, u2, acc0.w);
        acc1.x = fmaf(k3.x, u3, acc1.x); acc1.y = fmaf(k3.y, u3, acc1.y); acc1.z = fmaf(k3.z, u3, acc1.z); acc1.w = fmaf(k3.w, u3, acc1.w);
      }
      for (; r < S; ++r) {
        const float4 k0 = sbase[r * ng];
        const float u0 = u_s[8 + r];
        acc0.x = fmaf(k0.x, u0, acc0.x); acc0.y = fmaf(k0.y, u0, acc0.y); acc0.z = fmaf(k0.z, u0, acc0.z); acc0.w = fmaf(k0.w, u0, acc0.w);
      }
#pragma unroll
      for (int rr = 0; rr < RR; ++rr) {
        const float u0 = u_s[rr];
        acc1.x = fmaf(kreg[rr][q].x, u0, acc1.x); acc1.y = fmaf(kreg[rr][q].y, u0, acc1.y);
        acc1.z = fmaf(kreg[rr][q].z, u0, acc1.z); acc1.w = fmaf(kreg[rr][q].w, u0, acc1.w);
      }
      reinterpret_cast<float4*>(part + (int64_t)cta * J)[gq[q]] =
          make_float4(acc0.x + acc1.x, acc0.y + acc1.y, acc0.z + acc1.z, acc0.w + acc1.w);
    }
    if (timer) t1 = gtime();
    grid_barrier(&st->barrier, target, nb);
    if (timer) t2 = gtime();
    // ---- R: total column sums for my slice of columns -> v ---------------------------------------------
    const bool check = (cpt >= 1) && ((cpt - 1) % 10 == 0);
    const int slot = check ? ((cpt - 1) / 10) & 127 : 0;
    {
      float sum = 0.f;
      if (lane < ncols) {
        float pv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const int pidx = warp + kOcWarps * k;
          pv[k] = (pidx < nb) ? __ldcg(part + (int64_t)pidx * J + col0 + lane) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) sum += pv[k];
      }
      red[warp * kScSlots + lane] = sum;
    }
    __syncthreads();
    float v_cand = 1.0f;
    if (warp == 0) {
      double d2 = 0.0;
      bool bad = u_bad, big = u_big;
      if (lane < ncols) {
        float tot = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kOcWarps; ++w2) tot += red[w2 * kScSlots + lane];
        if (check) {
          const double d = (double)(v_own * tot) - (double)b_own;
          d2 = d * d;
        }
        v_cand = b_own / tot;
        bad = bad || !(tot > 0.f) || !(v_cand < CUDART_INF_F) || !(v_cand > 0.f);
        big = big || fabsf(lg2f(v_cand)) > kAbsorbLog2;
        v_buf[(int64_t)nxt * J + col0 + lane] = v_cand;
      }
      if (check) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        if (lane == 0 && d2 != 0.0) atomicAdd(&st->err2[slot], d2);
      }
      if (__any_sync(0xffffffffu, bad) && lane == 0) st->fallback = 1;
      if (__any_sync(0xffffffffu, big) && lane == 0) st->absorb_req = cpt + 1;
    }
    if (timer) t3 = gtime();
    grid_barrier(&st->barrier, target, nb);
    if (timer) t4 = gtime();
    // the new v (written by its owners before the barrier) is requested first so that its L2 round trip overlaps
    // the flag / error reads below
    float4 vq[kOcQ];
    {
      const float4* src = reinterpret_cast<const float4*>(v_buf + (int64_t)nxt * J);
#pragma unroll
      for (int q = 0; q < kOcQ; ++q) {
        vq[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gok[q])
          asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(vq[q].x), "=f"(vq[q].y), "=f"(vq[q].z), "=f"(vq[q].w) : "l"(src + gq[q]) : "memory");
      }
    }
    int fb, areq;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(fb) : "l"(&st->fallback) : "memory");
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(areq) : "l"(&st->absorb_req) : "memory");
    double e2 = 0.0;
    if (check) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(e2) : "l"(&st->err2[slot]) : "memory");
    if (fb) break;
    if (check) {
      err = sqrt(e2);
      if (!(err > stop_thr)) break;                    // keep (u, v_own): sweep cpt never happened
    }
    v_own = v_cand;
    const bool absorb = (areq == cpt + 1);
    if (timer) st->t_phase[5] += gtime() - t4;
    // ---- U: row sums with the new v -> u -----------------------------------------------------------------
    // register rows, then the shared-memory rows in batches of 8 (two blocks of 4 rows: 8 LDS.128 in flight,
    // unconditional with a clamped row; dead column groups carry v = 0), each batch reduced across the warp
    // with the 8-value transpose
    {
      // all dot products first (register rows + shared-memory rows 0..23, clamped / zero-selected, so the loads
      // are unconditional and deep in flight), then the four 8-value transposes back to back so their shuffle
      // latencies overlap; rows 24..31 (only when S > 24) go through a separate, uniform-branch batch
      float pr[8], pb[2][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pr[k] = 0.f;
#pragma unroll
      for (int rr = 0; rr < RR; ++rr)
#pragma unroll
        for (int q = 0; q < kOcQ; ++q)
          pr[rr] += (kreg[rr][q].x * vq[q].x + kreg[rr][q].y * vq[q].y) + (kreg[rr][q].z * vq[q].z + kreg[rr][q].w * vq[q].w);
      if (timer) st->t_phase[6] += gtime() - t4;
      const int last_row = max(S, 1) - 1;          // a CTA past the end of the matrix has S == 0: stay inside kt
      auto rows4 = [&](int r0, float (&dst)[8], int off) {
        float4 k0[4];
        float dot[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) k0[u] = kt[min(r0 + u, last_row) * ng + gqs0];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          dot[u] = (k0[u].x * vq[0].x + k0[u].y * vq[0].y) + (k0[u].z * vq[0].z + k0[u].w * vq[0].w);
#pragma unroll
        for (int u = 0; u < 4; ++u) k0[u] = kt[min(r0 + u, last_row) * ng + gqs1];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dot[u] += (k0[u].x * vq[1].x + k0[u].y * vq[1].y) + (k0[u].z * vq[1].z + k0[u].w * vq[1].w);
          dst[off + u] = (r0 + u < S) ? dot[u] : 0.f;
        }
      };
#pragma unroll
      for (int batch = 0; batch < 2; ++batch) {
        rows4(batch * 8, pb[batch], 0);
        rows4(batch * 8 + 4, pb[batch], 4);
      }
      const float tr = warp_transpose_sum8(pr, lane);
      const float t0 = warp_transpose_sum8(pb[0], lane);
      const float t1 = warp_transpose_sum8(pb[1], lane);
      if (lane < 8) {
        red[warp * kScSlots + lane] = tr;
        red[warp * kScSlots + 8 + lane] = t0;
        red[warp * kScSlots + 16 + lane] = t1;
      }
      // rows 16..31 in two more batches, each behind a uniform branch (19 rows at J = 3000: one of them, 3 rows live)
#pragma unroll
      for (int batch = 2; batch < 4; ++batch) {
        if (batch * 8 < S) {
          rows4(batch * 8, pr, 0);
          if (batch * 8 + 4 < S) {
            rows4(batch * 8 + 4, pr, 4);
          } else {
#pragma unroll
            for (int k = 4; k < 8; ++k) pr[k] = 0.f;
          }
          const float t2 = warp_transpose_sum8(pr, lane);
          if (lane < 8) red[warp * kScSlots + 8 + batch * 8 + lane] = t2;
        } else if (lane < 8) {
          red[warp * kScSlots + 8 + batch * 8 + lane] = 0.f;
        }
      }
    }
    if (timer) st->t_phase[7] += gtime() - t4;
    __syncthreads();
    if (tid < kScSlots) {
      float tot = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kOcWarps; ++w2) tot += red[w2 * kScSlots + tid];
      if (my_live) {
        const float u = a_s[tid] / tot;
        if (!(tot > 0.f) || !(u < CUDART_INF_F) || !(u > 0.f)) atomicOr(&flag_s[0], 1);
        if (fabsf(lg2f(u)) > kAbsorbLog2) atomicOr(&flag_s[1], 1);
        u_s[tid] = u;
      }
    }
    __syncthreads();
    if (warp == 0) {                                   // only warp 0 (phase R, epilogue) consumes these
      u_bad = u_bad || (flag_s[0] != 0);
      u_big = (flag_s[1] != 0);
      __syncwarp();
      if (lane == 0) flag_s[1] = 0;                    // next writers are two grid barriers away
    }
    if (absorb) {
      // fold u, v into the on-chip kernel entries and the potentials; scalings restart from 1
#pragma unroll
      for (int q = 0; q < kOcQ; ++q) {
        if (!gok[q]) continue;
#pragma unroll
        for (int rr = 0; rr < RR; ++rr) {
          const float u0 = u_s[rr];
          kreg[rr][q].x *= u0 * vq[q].x; kreg[rr][q].y *= u0 * vq[q].y;
          kreg[rr][q].z *= u0 * vq[q].z; kreg[rr][q].w *= u0 * vq[q].w;
        }
        for (int r = 0; r < S; ++r) {
          const float u0 = u_s[8 + r];
          float4 k4 = kt[r * ng + gq[q]];
          k4.x *= u0 * vq[q].x; k4.y *= u0 * vq[q].y; k4.z *= u0 * vq[q].z; k4.w *= u0 * vq[q].w;
          kt[r * ng + gq[q]] = k4;
        }
      }
      __syncthreads();
      // folded amounts accumulate in fp64 (an fp32 potential of O(100) has an ulp of 1.5e-5: a fold per sweep would
      // random-walk past the 1e-4 parity bar)
      if (my_live) { LU_fold += (double)log2f(u_s[tid]); u_s[tid] = 1.0f; }
      if (warp == 0 && lane < ncols) { LV_fold += (double)log2f(v_own); v_own = 1.0f; }
      u_big = false;
      if (cta == 0 && tid == 0) st->absorbs += 1;
      __syncthreads();
    }
    sweeps = cpt + 1;
    if (timer) {
      st->t_phase[0] += t1 - t0; st->t_phase[1] += t2 - t1; st->t_phase[2] += t3 - t2;
      st->t_phase[3] += t4 - t3; st->t_phase[4] += gtime() - t4;
    }
  }
  __syncthreads();
  if (u_bad && tid == 0) st->fallback = 1;             // last sweep's row sums: no later phase R to publish it
  if (my_live) log_u_out[row0 + my_row] = (float)(((double)LU_s[tid] + LU_fold + (double)log2f(u_s[tid])) * 0.69314718055994530942);
  if (warp == 0 && lane < ncols)
    log_v_out[col0 + lane] = (float)(((double)LV_own + LV_fold + (double)log2f(v_own)) * 0.69314718055994530942);
  if (cta == 0 && tid == 0) { st->sweeps = sweeps; st->final_buf = 0; st->err = err; }
}

template <typename T>
struct SolveWs {
  T *log_a, *log_b, *lv_alt, *col_lse;
  double* err2;
  T *part_m, *part_s, *lv_buf;     // persistent-kernel scratch: [SMs, J] x2, [2, J]
  PersistState* state;
  PersistState* state2;            // tile kernel (scaling-domain continuation)
  PersistState* state3;            // conditional log-domain redo
  size_t total;
};
template <typename T>
static SolveWs<T> carve_solve(void* ws, int64_t I, int64_t J) {
  SolveWs<T> w{};
  char* p = reinterpret_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes); return q; };
  w.log_a = (T*)take(sizeof(T) * (size_t)I);
  w.log_b = (T*)take(sizeof(T) * (size_t)J);
  w.lv_alt = (T*)take(sizeof(T) * (size_t)J);
  w.col_lse = (T*)take(sizeof(T) * (size_t)J);
  w.err2 = (double*)take(sizeof(double));
  w.part_m = (T*)take(sizeof(T) * (size_t)kNumSMs * (size_t)J);
  w.part_s = (T*)take(sizeof(T) * (size_t)kNumSMs * (size_t)J);
  w.lv_buf = (T*)take(sizeof(T) * 2 * (size_t)J);
  w.state = (PersistState*)take(sizeof(PersistState));
  w.state2 = (PersistState*)take(sizeof(PersistState));
  w.state3 = (PersistState*)take(sizeof(PersistState));
  w.total = off;
  return w;
}

int g_tune_persistent = 1;   // eg_debug_set(3, 0) forces the streaming path
int g_tune_resident = 1;     // eg_debug_set(4, 0) keeps no rows of M in shared memory
int g_tune_onchip = 1;       // eg_debug_set(5, 0) disables the fully on-chip fp32 kernel
int g_tune_scaling = 1;      // eg_debug_set(7, 0) keeps the whole on-chip solve in the log domain
int g_tune_absorb_milli = 32000;   // eg_debug_set(10, x): fold u, v into the kernel once |log2| exceeds x / 1000
int g_tune_force_fallback = 0;     // eg_debug_set(11, 1): treat every scaling-domain solve as failed (tests the redo path)
int g_tune_tile2d = 1;              // eg_debug_set(12, 0): row-block scaling kernel instead of the 2-D tiled one
int g_sinkhorn_fallbacks = 0;
int g_sinkhorn_absorbs = 0;

int sinkhorn_redo_count() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_dev_sinkhorn_redos, sizeof(int));
  return g_sinkhorn_fallbacks + v;
}
int sinkhorn_absorb_count() { return g_sinkhorn_absorbs + sinkhorn_tile2d_absorbs_read(); }
constexpr int kWarmSweeps = 4;   // log-domain sweeps before the scaling-domain kernel takes over

// One cooperative launch for the whole solve when the shape allows it.
template <typename T>
static int sinkhorn_persistent_t(const T* M, int64_t I, int64_t J, double reg, const T* a, const T* b, int max_iter,
                                 double stop_thr, T* log_u, T* log_v, SolveWs<T>& w, int* h_sweeps, double* h_err,
                                 cudaStream_t s, bool* used) {
  *used = false;
  constexpr int V = VecOf<T>::N;
  if (!g_tune_persistent || J % V != 0 || (reinterpret_cast<uintptr_t>(M) & 15) || max_iter > 1270) return EG_OK;
  int dev = 0, coop = 0, sms = 0;
  EG_CUDA(cudaGetDevice(&dev));
  EG_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  EG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!coop || sms > kNumSMs) return EG_OK;
  const int64_t rows_per = ceil_div(I, (int64_t)sms);
  int max_smem = 0;
  EG_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if constexpr (sizeof(T) == 4) {
    // whole row block on chip (shared memory + registers)?  Needs J <= 4096 and rows_per <= S + kOcRR, nb <= 160.
    const size_t oc_fixed = sizeof(float) * (size_t)(J + 32 + 2 * 32 * kOcWarps + 2 * 8 * kOcWarps) + 16;
    const int64_t s_fit = ((int64_t)max_smem - 1024 - (int64_t)oc_fixed) / (int64_t)(sizeof(float) * (size_t)J);
    const int64_t s_need = std::max<int64_t>(rows_per - kOcRR, std::min<int64_t>(rows_per, kOcWarps));
    if (g_tune_onchip && J <= 4 * kOcThreads * kOcQ && rows_per <= 32 && sms <= 160 && s_need <= s_fit &&
        rows_per - s_need <= kOcRR && rows_per - s_need >= 0) {
      int S_max = (int)s_need;
      const size_t smem = oc_fixed + sizeof(float) * (size_t)J * (size_t)S_max;
      auto kern = sinkhorn_onchip_kernel;
      EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      EG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kOcThreads, smem));
      if (per_sm >= 1) {
        const int TB = 256;
        log_kernel<T><<<(unsigned)ceil_div(I, TB), TB, 0, s>>>(a, I, w.log_a); EG_LAUNCHED();
        log_kernel<T><<<(unsigned)ceil_div(J, TB), TB, 0, s>>>(b, J, w.log_b); EG_LAUNCHED();
        float inv_reg = (float)(1.0 / reg);
        int64_t ld = J;
        int Ji = (int)J;
        PersistState host_state;
        // one cooperative launch of the log-domain kernel; `run_if` != null makes it conditional on a device flag
        auto launch_log = [&](int iters, PersistState* state, const int* run_if) -> int {
          fill_kernel<T><<<(unsigned)ceil_div(J, TB), TB, 0, s>>>(w.lv_buf, J, (T)(-log2((double)J))); EG_LAUNCHED();
          EG_CUDA(cudaMemsetAsync(state, 0, sizeof(PersistState), s));
          void* args[] = {(void*)&M, (void*)&I, (void*)&Ji, (void*)&ld, (void*)&inv_reg, (void*)&w.log_a,
                          (void*)&w.log_b, (void*)&b, (void*)&iters, (void*)&stop_thr, (void*)&w.part_m,
                          (void*)&w.part_s, (void*)&w.lv_buf, (void*)&log_u, (void*)&log_v, (void*)&state,
                          (void*)&S_max, (void*)&run_if};
          EG_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)sms), dim3(kOcThreads), args, smem, s));
          g_launches.fetch_add(1, std::memory_order_relaxed);
          return EG_OK;
        };
        auto run_log = [&](int iters) -> int {
          int rc = launch_log(iters, w.state, nullptr);
          if (rc) return rc;
          EG_CUDA(cudaMemcpyAsync(&host_state, w.state, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
          EG_CUDA(cudaStreamSynchronize(s));
          return EG_OK;
        };
        auto report = [&](const char* what) {
          if (!getenv("EG_PERSIST_TIMING") || host_state.sweeps <= 0) return;
          const PersistState& full = host_state;
          fprintf(stderr, "[eagraft] on-chip sinkhorn %s (S=%d): %d sweeps, %d absorptions, fallback %d; CTA0 us/sweep: C %.2f | bar %.2f | R %.2f | bar %.2f | U %.2f (U marks: %.2f, %.2f, %.2f)\n",
                  what, S_max, full.sweeps, full.absorbs, full.fallback, full.t_phase[0] / 1e3 / full.sweeps,
                  full.t_phase[1] / 1e3 / full.sweeps, full.t_phase[2] / 1e3 / full.sweeps,
                  full.t_phase[3] / 1e3 / full.sweeps, full.t_phase[4] / 1e3 / full.sweeps,
                  full.t_phase[5] / 1e3 / full.sweeps, full.t_phase[6] / 1e3 / full.sweeps,
                  full.t_phase[7] / 1e3 / full.sweeps);
        };
        // ---- default: warm-up -> 2-D tiled scaling kernel -> conditional redo, with NO host synchronisation unless
        // the caller's stop rule needs the sweep count / error back (stop_thr >= 0)
        bool warm_done = false;
        static const bool env_tile2d_off = getenv("EG_TILE2D") != nullptr && getenv("EG_TILE2D")[0] == '0';
        if (g_tune_scaling && g_tune_tile2d && !env_tile2d_off && max_iter >= kWarmSweeps + 8) {
          int rc = launch_log(kWarmSweeps, w.state, nullptr);
          if (rc) return rc;
          EG_CUDA(cudaMemsetAsync(w.state2, 0, sizeof(PersistState), s));
          bool launched = false;
          rc = sinkhorn_tile2d_launch(M, I, J, ld, 1.0 / reg, a, b, log_u, log_v, w.state, kWarmSweeps, max_iter,
                                      stop_thr, w.part_m, sizeof(float) * (size_t)kNumSMs * (size_t)J, w.state2,
                                      (float)g_tune_absorb_milli / 1000.0f, g_tune_force_fallback, s, &launched);
          if (rc) return rc;
          if (launched) {
            rc = launch_log(max_iter, w.state3, &w.state2->fallback);
            if (rc) return rc;
            const bool timing = getenv("EG_PERSIST_TIMING") != nullptr;
            if (stop_thr >= 0.0 || timing) {
              PersistState hs[3];
              EG_CUDA(cudaMemcpyAsync(&hs[0], w.state, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
              EG_CUDA(cudaMemcpyAsync(&hs[1], w.state2, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
              EG_CUDA(cudaMemcpyAsync(&hs[2], w.state3, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
              EG_CUDA(cudaStreamSynchronize(s));
              const PersistState& fin = hs[0].sweeps < kWarmSweeps ? hs[0] : (hs[1].fallback ? hs[2] : hs[1]);
              if (h_sweeps) *h_sweeps = fin.sweeps;
              if (h_err) *h_err = fin.err;
              if (timing) {
                const double n = std::max(1, hs[1].sweeps - kWarmSweeps);
                fprintf(stderr, "[eagraft] tile2d sinkhorn: warm-up %d sweeps; %d sweeps, fallback %d, redo sweeps %d; CTA0 us/sweep: C %.2f | column exchange %.2f | v %.2f | U dots %.2f | row exchange + u %.2f\n",
                        hs[0].sweeps, hs[1].sweeps, hs[1].fallback, hs[2].sweeps, hs[1].t_phase[0] / 1e3 / n,
                        hs[1].t_phase[1] / 1e3 / n, hs[1].t_phase[2] / 1e3 / n, hs[1].t_phase[3] / 1e3 / n,
                        hs[1].t_phase[4] / 1e3 / n);
                if (hs[1].fallback >= 2) fprintf(stderr, "[eagraft] tile2d: a wait on another CTA timed out\n");
              }
            } else {
              // every sweep runs and nothing is read back: the whole solve stays asynchronous
              if (h_sweeps) *h_sweeps = max_iter;
              if (h_err) *h_err = -1.0;
            }
            *used = true;
            return EG_OK;
          }
          // tile kernel not available for this shape / device: the warm-up result is reused below
          EG_CUDA(cudaMemcpyAsync(&host_state, w.state, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
          EG_CUDA(cudaStreamSynchronize(s));
          warm_done = true;
        }
        bool done = false;
        if (g_tune_scaling && max_iter >= kWarmSweeps + 8 && ceil_div(J, (int64_t)sms) <= 32) {
          int rc = warm_done ? EG_OK : run_log(kWarmSweeps);
          if (rc) return rc;
          report("log-domain warm-up");
          if (host_state.sweeps < kWarmSweeps) {
            done = true;                                  // the stop rule fired inside the warm-up
          } else {
            // hand-over copies: the scaling kernel reads every CTA's potentials while others may already write theirs
            EG_CUDA(cudaMemcpyAsync(w.log_a, log_u, sizeof(T) * (size_t)I, cudaMemcpyDeviceToDevice, s));
            EG_CUDA(cudaMemcpyAsync(w.lv_alt, log_v, sizeof(T) * (size_t)J, cudaMemcpyDeviceToDevice, s));
            EG_CUDA(cudaMemsetAsync(w.state, 0, sizeof(PersistState), s));
            // this kernel's fixed shared memory is smaller than the log-domain one's, so more rows fit and fewer
            // have to live in registers
            const size_t fixed2 = sizeof(float) * (size_t)(3 * kScSlots + kScSlots * kOcWarps + 4);
            const int64_t fit2 = std::min<int64_t>(32, ((int64_t)max_smem - 1024 - (int64_t)fixed2) /
                                                           (int64_t)(sizeof(float) * (size_t)J));
            int S2 = (int)std::min<int64_t>(rows_per, fit2);
            const bool rr2 = rows_per - S2 <= 2;
            if (!rr2) S2 = S_max;                       // fall back to the 5-register-row split
            auto kern2 = rr2 ? sinkhorn_onchip_scaling_kernel<2> : sinkhorn_onchip_scaling_kernel<kOcRR>;
            const size_t smem2 = fixed2 + sizeof(float) * (size_t)J * (size_t)S2;
            EG_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            int start = kWarmSweeps;
            float absorb_log2 = (float)g_tune_absorb_milli / 1000.0f;
            void* args2[] = {(void*)&M, (void*)&I, (void*)&Ji, (void*)&ld, (void*)&inv_reg, (void*)&a, (void*)&b,
                             (void*)&w.log_a, (void*)&w.lv_alt, (void*)&start, (void*)&max_iter, (void*)&stop_thr,
                             (void*)&w.part_m, (void*)&w.lv_buf, (void*)&log_u, (void*)&log_v, (void*)&w.state,
                             (void*)&S2, (void*)&absorb_log2};
            EG_CUDA(cudaLaunchCooperativeKernel((void*)kern2, dim3((unsigned)sms), dim3(kOcThreads), args2, smem2, s));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            EG_CUDA(cudaMemcpyAsync(&host_state, w.state, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
            EG_CUDA(cudaStreamSynchronize(s));
            report("scaling-domain");
            g_sinkhorn_absorbs = host_state.absorbs;
            if (host_state.fallback || g_tune_force_fallback) {
              ++g_sinkhorn_fallbacks;
              // w.log_a was used as a hand-over buffer: restore log a for the log-domain kernel
              log_kernel<T><<<(unsigned)ceil_div(I, TB), TB, 0, s>>>(a, I, w.log_a); EG_LAUNCHED();
            } else {
              done = true;
            }
          }
        }
        if (!done) {
          int rc = run_log(max_iter);
          if (rc) return rc;
          report("log-domain");
        }
        if (h_sweeps) *h_sweeps = host_state.sweeps;
        if (h_err) *h_err = host_state.err;
        *used = true;
        return EG_OK;
      }
    }
  }
  const size_t fixed = sizeof(T) * (size_t)(J + (rows_per + 3) / 4 * 4 + 2 * 32 * (kPersistThreads / 32)) + 16;
  if (fixed > 160 * 1024) return EG_OK;
  // as many rows of the CTA's block as fit stay resident in shared memory for the whole solve
  int rows_resident = (int)std::min<int64_t>(rows_per, ((int64_t)max_smem - 1024 - (int64_t)fixed) /
                                                         (int64_t)(sizeof(T) * (size_t)J));
  if (rows_resident < 0 || !g_tune_resident) rows_resident = 0;
  const size_t smem = fixed + sizeof(T) * (size_t)J * (size_t)rows_resident;
  auto kern = sinkhorn_persistent_kernel<T>;
  EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  EG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPersistThreads, smem));
  if (per_sm < 1) return EG_OK;
  const int TB = 256;
  log_kernel<T><<<(unsigned)ceil_div(I, TB), TB, 0, s>>>(a, I, w.log_a); EG_LAUNCHED();
  log_kernel<T><<<(unsigned)ceil_div(J, TB), TB, 0, s>>>(b, J, w.log_b); EG_LAUNCHED();
  fill_kernel<T><<<(unsigned)ceil_div(J, TB), TB, 0, s>>>(w.lv_buf, J, (T)(-log((double)J))); EG_LAUNCHED();
  EG_CUDA(cudaMemsetAsync(w.state, 0, sizeof(PersistState), s));
  T inv_reg = (T)(1.0 / reg);
  int64_t ld = J;
  void* args[] = {(void*)&M, (void*)&I, (void*)&J, (void*)&ld, (void*)&inv_reg, (void*)&w.log_a, (void*)&w.log_b,
                  (void*)&b, (void*)&max_iter, (void*)&stop_thr, (void*)&w.part_m, (void*)&w.part_s,
                  (void*)&w.lv_buf, (void*)&log_u, (void*)&log_v, (void*)&w.state, (void*)&rows_resident};
  EG_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)sms), dim3(kPersistThreads), args, smem, s));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  PersistState host_state;
  EG_CUDA(cudaMemcpyAsync(&host_state, w.state, sizeof(int) * 3 + sizeof(double) + 4, cudaMemcpyDeviceToHost, s));
  EG_CUDA(cudaStreamSynchronize(s));
  if (h_sweeps) *h_sweeps = host_state.sweeps;
  if (h_err) *h_err = host_state.err;
  if (getenv("EG_PERSIST_TIMING")) {
    PersistState full;
    cudaMemcpy(&full, w.state, sizeof(PersistState), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[eagraft] persistent sinkhorn: %d sweeps; CTA0 us/sweep: C %.2f | bar %.2f | R %.2f | bar %.2f | U %.2f\n",
            full.sweeps, full.t_phase[0] / 1e3 / full.sweeps, full.t_phase[1] / 1e3 / full.sweeps,
            full.t_phase[2] / 1e3 / full.sweeps, full.t_phase[3] / 1e3 / full.sweeps, full.t_phase[4] / 1e3 / full.sweeps);
  }
  *used = true;
  return EG_OK;
}

template <typename T>
static int sinkhorn_dense_t(const T* M, int64_t I, int64_t J, double reg, const T* a, const T* b, int max_iter,
                            double stop_thr, T* Mt, T* log_u, T* log_v, void* ws, size_t ws_bytes, int* h_sweeps,
                            double* h_err, cudaStream_t s) {
  SolveWs<T> w = carve_solve<T>(ws, I, J);
  if (ws_bytes < w.total) return EG_ERR_WORKSPACE;
  {
    bool used = false;
    int prc = sinkhorn_persistent_t<T>(M, I, J, reg, a, b, max_iter, stop_thr, log_u, log_v, w, h_sweeps, h_err, s,
                                       &used);
    if (prc) return prc;
    if (used) return EG_OK;
  }
  if (Mt == nullptr) return EG_ERR_WORKSPACE;      // only this streaming path needs the transposed copy
  const double inv_reg = 1.0 / reg;
  const int TB = 256;
  int rc = transpose_t<T>(M, I, J, J, Mt, I, s);
  if (rc) return rc;
  log_kernel<T><<<(unsigned)ceil_div(I, TB), TB, 0, s>>>(a, I, w.log_a); EG_LAUNCHED();
  log_kernel<T><<<(unsigned)ceil_div(J, TB), TB, 0, s>>>(b, J, w.log_b); EG_LAUNCHED();
  fill_kernel<T><<<(unsigned)ceil_div(I, TB), TB, 0, s>>>(log_u, I, (T)(-log((double)I))); EG_LAUNCHED();
  fill_kernel<T><<<(unsigned)ceil_div(J, TB), TB, 0, s>>>(log_v, J, (T)(-log((double)J))); EG_LAUNCHED();
  T* lv_cur = log_v;     // holds the accepted log v
  T* lv_new = w.lv_alt;
  double err = 1.0;
  int sweeps = 0;
  bool stopped = false;
  for (int cpt = 0; cpt < max_iter; ++cpt) {
    // column half-sweep: log v_j = log b_j - LSE_i(log u_i - M_ij/reg)
    rc = lse_dense_t<T>(Mt, J, I, I, inv_reg, log_u, w.log_b, lv_new, w.col_lse, s);
    if (rc) return rc;
    if (cpt >= 1 && (cpt - 1) % 10 == 0) {
      // marginal error of sweep cpt-1: v_old ∘ (Kᵀ u) vs b — Kᵀu is exp(col_lse) just computed
      EG_CUDA(cudaMemsetAsync(w.err2, 0, sizeof(double), s));
      marginal_err_kernel<T><<<64, TB, 0, s>>>(lv_cur, w.col_lse, b, J, w.err2); EG_LAUNCHED();
      double e2 = 0.0;
      EG_CUDA(cudaMemcpyAsync(&e2, w.err2, sizeof(double), cudaMemcpyDeviceToHost, s));
      EG_CUDA(cudaStreamSynchronize(s));
      err = sqrt(e2);
      if (!(err > stop_thr)) { stopped = true; break; }   // keep (log_u, lv_cur): sweep cpt never happened
    }
    T* tmp = lv_cur; lv_cur = lv_new; lv_new = tmp;
    // row half-sweep: log u_i = log a_i - LSE_j(log v_j - M_ij/reg)
    rc = lse_dense_t<T>(M, I, J, J, inv_reg, lv_cur, w.log_a, log_u, (T*)nullptr, s);
    if (rc) return rc;
    sweeps = cpt + 1;
  }
  (void)stopped;
  if (lv_cur != log_v) EG_CUDA(cudaMemcpyAsync(log_v, lv_cur, sizeof(T) * (size_t)J, cudaMemcpyDeviceToDevice, s));
  EG_CUDA(cudaStreamSynchronize(s));
  if (h_sweeps) *h_sweeps = sweeps;
  if (h_err) *h_err = err;
  return EG_OK;
}

}  // namespace eg

extern "C" {

int eg_lse_dense(int dtype, const void* M, int64_t n_rows, int64_t n_cols, int64_t ld, double inv_reg,
                 const void* pot_in, const void* logw, void* pot_out, void* lse_out, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || ld < n_cols || (dtype != 0 && dtype != 1)) return EG_ERR_INVALID;
  if (n_rows == 0) return EG_OK;
  if (!M || !pot_in || (!pot_out && !lse_out)) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (dtype == 0)
    return lse_dense_t<float>((const float*)M, n_rows, n_cols, ld, inv_reg, (const float*)pot_in,
                              (const float*)logw, (float*)pot_out, (float*)lse_out, s);
  return lse_dense_t<double>((const double*)M, n_rows, n_cols, ld, inv_reg, (const double*)pot_in,
                             (const double*)logw, (double*)pot_out, (double*)lse_out, s);
}

int eg_transpose(int dtype, const void* src, int64_t n_rows, int64_t n_cols, int64_t ld_src, void* dst,
                 int64_t ld_dst, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || ld_src < n_cols || ld_dst < n_rows || (dtype != 0 && dtype != 1))
    return EG_ERR_INVALID;
  if (n_rows == 0 || n_cols == 0) return EG_OK;
  if (!src || !dst) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (dtype == 0) return transpose_t<float>((const float*)src, n_rows, n_cols, ld_src, (float*)dst, ld_dst, s);
  return transpose_t<double>((const double*)src, n_rows, n_cols, ld_src, (double*)dst, ld_dst, s);
}

int eg_plan_dense(int dtype, const void* M, int64_t n_rows, int64_t n_cols, int64_t ld, double inv_reg,
                  const void* f, const void* g, void* P, int64_t ldP, double* loss, void* row_sum,
                  void* col_sum, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || ld < n_cols || (dtype != 0 && dtype != 1)) return EG_ERR_INVALID;
  if (P && ldP < n_cols) return EG_ERR_INVALID;
  if (n_rows == 0 || n_cols == 0) return EG_OK;
  if (!M || !f || !g) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (dtype == 0)
    return plan_t<float>((const float*)M, n_rows, n_cols, ld, inv_reg, (const float*)f, (const float*)g,
                         (float*)P, ldP, loss, (float*)row_sum, (float*)col_sum, s);
  return plan_t<double>((const double*)M, n_rows, n_cols, ld, inv_reg, (const double*)f, (const double*)g,
                        (double*)P, ldP, loss, (double*)row_sum, (double*)col_sum, s);
}

int eg_sinkhorn_sync_floor(int64_t n_rows, int64_t n_cols, int iters, void* ws, size_t ws_bytes, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows <= 0 || n_cols <= 0 || iters <= 0 || !ws) return EG_ERR_INVALID;
  SolveWs<float> w = carve_solve<float>(ws, n_rows, n_cols);
  if (ws_bytes < w.total) return EG_ERR_WORKSPACE;
  bool launched = false;
  int rc = sinkhorn_tile2d_sync_floor_launch(n_rows, n_cols, iters, w.part_m, sizeof(float) * (size_t)kNumSMs * (size_t)n_cols, w.state2,
                                             as_stream(stream_), &launched);
  if (rc) return rc;
  return launched ? EG_OK : EG_ERR_UNSUPPORTED;
}

size_t eg_sinkhorn_dense_workspace_bytes(int dtype, int64_t n_rows, int64_t n_cols) {
  if (n_rows <= 0 || n_cols <= 0) return 0;
  return dtype == 0 ? eg::carve_solve<float>(nullptr, n_rows, n_cols).total
                    : eg::carve_solve<double>(nullptr, n_rows, n_cols).total;
}

int eg_sinkhorn_dense(int dtype, const void* M, int64_t n_rows, int64_t n_cols, double reg, const void* a,
                      const void* b, int max_iter, double stop_thr, void* Mt, void* log_u, void* log_v, void* ws,
                      size_t ws_bytes, int* h_sweeps, double* h_err, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows <= 0 || n_cols <= 0 || !(reg > 0.0) || max_iter < 0 || (dtype != 0 && dtype != 1))
    return EG_ERR_INVALID;
  if (!M || !a || !b || !log_u || !log_v || !ws) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (dtype == 0)
    return sinkhorn_dense_t<float>((const float*)M, n_rows, n_cols, reg, (const float*)a, (const float*)b,
                                   max_iter, stop_thr, (float*)Mt, (float*)log_u, (float*)log_v, ws, ws_bytes,
                                   h_sweeps, h_err, s);
  return sinkhorn_dense_t<double>((const double*)M, n_rows, n_cols, reg, (const double*)a, (const double*)b,
                                  max_iter, stop_thr, (double*)Mt, (double*)log_u, (double*)log_v, ws, ws_bytes,
                                  h_sweeps, h_err, s);
}

}  // extern "C"
