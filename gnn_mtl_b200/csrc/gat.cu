// GAT edge-softmax aggregation: the sparse attention of layers/att_layers.py:29-61
// as ONE gather kernel — the per-edge weights are never materialised in the forward pass:
//     t_ij = s1_i + s2_j,   w_ij = exp(-leakyrelu_alpha(t_ij)),   W_i = sum_j w_ij,
//     y_i  = (sum_j w_ij h_j) / W_i                  (j over the stored columns of row i; A's values are unused)
// with s1 = h·a[:, :D]ᵀ, s2 = h·a[:, D:]ᵀ computed by the caller.  Same machinery as the SpMM (warp = row,
// neighbour rows gathered coalesced, CSR-order summation); lanes hold ceil(d/32) scalars each so the
// reference's head width d = 300/4 = 75 needs no padding.
// Backward: one row-wise kernel produces p_ij = w_ij / W_i per edge, ds1, ds2; dh = Pᵀ·dy then runs on the
// ordinary SpMM over CSR(Aᵀ) with the p values permuted into transposed order.
#include "common.cuh"

namespace eg {

__device__ __forceinline__ float gat_weight(float t, float alpha) {
  return __expf(-(t > 0.f ? t : alpha * t));
}

template <int S>
__global__ void __launch_bounds__(128)
gat_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
               const float* __restrict__ h, int d, const float* __restrict__ s1, const float* __restrict__ s2,
               float alpha, const float* __restrict__ edge_scale, float* __restrict__ out,
               float* __restrict__ wsum) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  const float si = s1[row];
  float acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0.f;
  float wtot = 0.f;
  for (int base = b; base < e; base += 32) {
    const int idx = base + lane;
    int my_col = 0;
    float my_w = 0.f;
    if (idx < e) {
      my_col = ld_stream_i32(col + idx);
      my_w = gat_weight(si + __ldg(s2 + my_col), alpha);
    }
    wtot += my_w;                                        // the row sum is taken BEFORE edge dropout (:47 vs :50)
    if (edge_scale && idx < e) my_w *= ld_stream_f32(edge_scale + idx);
    const int cnt = min(32, e - base);
    for (int t = 0; t < cnt; t += 2) {
      float x[2][S], w[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = __shfl_sync(0xffffffffu, my_col, (t + u) & 31);
        const float ww = __shfl_sync(0xffffffffu, my_w, (t + u) & 31);
        const bool live = t + u < cnt;
        w[u] = live ? ww : 0.f;
        const float* rp = h + (int64_t)c * d;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int k = lane + 32 * s;
          x[u][s] = (live && k < d) ? __ldg(rp + k) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(w[u], x[u][s], acc[s]);
    }
  }
  wtot = warp_sum(wtot);
  const float inv = wtot > 0.f ? 1.0f / wtot : 0.f;      // isolated row: no neighbours -> 0 (the reference gives NaN)
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int k = lane + 32 * s;
    if (k < d) out[row * d + k] = acc[s] * inv;
  }
  if (lane == 0 && wsum) wsum[row] = wtot;
}

template <int S>
__global__ void __launch_bounds__(128)
gat_bwd_edges_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                     const float* __restrict__ h, int d, const float* __restrict__ s1, const float* __restrict__ s2,
                     float alpha, const float* __restrict__ edge_scale, const float* __restrict__ y,
                     const float* __restrict__ wsum, const float* __restrict__ dy, float* __restrict__ p_edge,
                     float* __restrict__ ds1, float* __restrict__ ds2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  const float si = s1[row];
  const float W = wsum[row];
  const float invW = W > 0.f ? 1.0f / W : 0.f;
  float g[S];
  float c_part = 0.f;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int k = lane + 32 * s;
    g[s] = (k < d) ? dy[row * d + k] : 0.f;
    c_part = fmaf(g[s], (k < d) ? y[row * d + k] : 0.f, c_part);
  }
  const float c_i = warp_sum(c_part);                    // dy_i · y_i
  float ds1_acc = 0.f;
  for (int base = b; base < e; base += 32) {
    const int idx = base + lane;
    int my_col = 0;
    float my_t = 0.f;
    if (idx < e) { my_col = ld_stream_i32(col + idx); my_t = si + __ldg(s2 + my_col); }
    const int cnt = min(32, e - base);
    float my_dot = 0.f;                                  // lane t ends up holding dy_i · h_{col t}
    for (int t = 0; t < cnt; t += 2) {
      float part[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = __shfl_sync(0xffffffffu, my_col, (t + u) & 31);
        const bool live = t + u < cnt;
        const float* rp = h + (int64_t)c * d;
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int k = lane + 32 * s;
          acc = fmaf(g[s], (live && k < d) ? __ldg(rp + k) : 0.f, acc);
        }
        part[u] = acc;
      }
      const float d0 = warp_sum(part[0]), d1 = warp_sum(part[1]);
      if (lane == t) my_dot = d0;
      if (lane == t + 1) my_dot = d1;
    }
    if (idx < e) {
      const float w = gat_weight(my_t, alpha);
      const float m = edge_scale ? ld_stream_f32(edge_scale + idx) : 1.0f;
      const float dw = (m * my_dot - c_i) * invW;        // d loss / d w_ij
      const float dt = dw * w * (my_t > 0.f ? -1.0f : -alpha);
      p_edge[idx] = w * m * invW;
      ds1_acc += dt;
      atomicAdd(ds2 + my_col, dt);
    }
  }
  ds1_acc = warp_sum(ds1_acc);
  if (lane == 0) ds1[row] = ds1_acc;
}

__global__ void permute_edges_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                                     float* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

template <int S>
static int gat_fwd_s(const int32_t* rowptr, const int32_t* col, int64_t n, const float* h, int d, const float* s1,
                     const float* s2, float alpha, const float* edge_scale, float* out, float* wsum, cudaStream_t s) {
  gat_fwd_kernel<S><<<(unsigned)ceil_div(n, 4), 128, 0, s>>>(rowptr, col, n, h, d, s1, s2, alpha, edge_scale, out,
                                                              wsum);
  EG_LAUNCHED();
  return EG_OK;
}
template <int S>
static int gat_bwd_s(const int32_t* rowptr, const int32_t* col, int64_t n, const float* h, int d, const float* s1,
                     const float* s2, float alpha, const float* edge_scale, const float* y, const float* wsum,
                     const float* dy, float* p_edge, float* ds1, float* ds2, cudaStream_t s) {
  gat_bwd_edges_kernel<S><<<(unsigned)ceil_div(n, 4), 128, 0, s>>>(rowptr, col, n, h, d, s1, s2, alpha, edge_scale, y,
                                                                    wsum, dy, p_edge, ds1, ds2);
  EG_LAUNCHED();
  return EG_OK;
}

#define EG_GAT_DISPATCH(FN, ...)                         \
  do {                                                   \
    const int slots = (d + 31) / 32;                     \
    if (slots <= 1) return FN<1>(__VA_ARGS__);           \
    if (slots <= 2) return FN<2>(__VA_ARGS__);           \
    if (slots <= 3) return FN<3>(__VA_ARGS__);           \
    if (slots <= 4) return FN<4>(__VA_ARGS__);           \
    if (slots <= 6) return FN<6>(__VA_ARGS__);           \
    if (slots <= 10) return FN<10>(__VA_ARGS__);         \
    if (slots <= 16) return FN<16>(__VA_ARGS__);         \
    return EG_ERR_UNSUPPORTED;                           \
  } while (0)

}  // namespace eg

extern "C" {

int eg_gat_fwd(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* h, int d, const float* s1,
               const float* s2, float alpha, const float* edge_scale, float* out, float* wsum,
               eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || d <= 0) return EG_ERR_INVALID;
  if (n_rows == 0) return EG_OK;
  if (!rowptr || !h || !s1 || !s2 || !out) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  EG_GAT_DISPATCH(gat_fwd_s, rowptr, col, n_rows, h, d, s1, s2, alpha, edge_scale, out, wsum, s);
}

int eg_gat_bwd_edges(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols, const float* h, int d,
                     const float* s1, const float* s2, float alpha, const float* edge_scale, const float* y,
                     const float* wsum, const float* dy, float* p_edge, float* ds1, float* ds2,
                     eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || d <= 0) return EG_ERR_INVALID;
  if (n_rows == 0) return EG_OK;
  if (!rowptr || !h || !s1 || !s2 || !y || !wsum || !dy || !p_edge || !ds1 || !ds2) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  EG_CUDA(cudaMemsetAsync(ds2, 0, sizeof(float) * (size_t)n_cols, s));
  EG_GAT_DISPATCH(gat_bwd_s, rowptr, col, n_rows, h, d, s1, s2, alpha, edge_scale, y, wsum, dy, p_edge, ds1, ds2, s);
}

int eg_permute_edges(const float* src, const int32_t* perm, int64_t n, float* dst, eg_stream_t stream_) {
  using namespace eg;
  if (n < 0) return EG_ERR_INVALID;
  if (n == 0) return EG_OK;
  if (!src || !perm || !dst) return EG_ERR_INVALID;
  permute_edges_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream_)>>>(src, perm, n, dst);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
