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

#include "common.cuh"

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
  __device__ __forceinline__ void push4(T z0, T z1, T z2, T z3) {
    T mx = fmax(fmax(z0, z1), fmax(z2, z3));
    if (mx > m) { s *= Num<T>::exp_(m - mx); m = mx; }   // exp(-inf) = 0 on the first push
    if (m > Num<T>::ninf())
      s += Num<T>::exp_(z0 - m) + Num<T>::exp_(z1 - m) + Num<T>::exp_(z2 - m) + Num<T>::exp_(z3 - m);
  }
  __device__ __forceinline__ void push1(T z) {
    if (z > m) { s *= Num<T>::exp_(m - z); m = z; }
    if (m > Num<T>::ninf()) s += Num<T>::exp_(z - m);
  }
  __device__ __forceinline__ void merge(T om, T os) {
    T mx = fmax(m, om);
    if (mx > Num<T>::ninf()) {
      s = s * Num<T>::exp_(m - mx) + os * Num<T>::exp_(om - mx);
      m = mx;
    }
  }
  __device__ __forceinline__ T value() const { return (s > T(0)) ? m + Num<T>::log_(s) : Num<T>::ninf(); }
};

template <typename T>
__device__ __forceinline__ void warp_merge(Lse<T>& a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T om = __shfl_xor_sync(0xffffffffu, a.m, o);
    T os = __shfl_xor_sync(0xffffffffu, a.s, o);
    a.merge(om, os);
  }
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

template <typename T>
struct SolveWs {
  T *log_a, *log_b, *lv_alt, *col_lse;
  double* err2;
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
  w.total = off;
  return w;
}

template <typename T>
static int sinkhorn_dense_t(const T* M, int64_t I, int64_t J, double reg, const T* a, const T* b, int max_iter,
                            double stop_thr, T* Mt, T* log_u, T* log_v, void* ws, size_t ws_bytes, int* h_sweeps,
                            double* h_err, cudaStream_t s) {
  SolveWs<T> w = carve_solve<T>(ws, I, J);
  if (ws_bytes < w.total) return EG_ERR_WORKSPACE;
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
  if (!M || !a || !b || !Mt || !log_u || !log_v || !ws) return EG_ERR_INVALID;
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
