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

// Hub rows (longer than `thresh`) are cut into segments exactly as in the SpMM (spmm.cu): warps
// [0, n_rows) take one short row each, warps [n_rows, n_rows + n_seg) one segment each, writing the
// un-normalised partial sum (d floats) and the partial weight sum to seg_scratch; gat_long_finish_kernel
// adds the partials in segment order and normalises.
struct GatSegments {
  int thresh;
  const int32_t* seg_row;
  const int32_t* seg_begin;
  const int32_t* seg_end;
  int64_t n_seg;
};

template <int S, int U>
__global__ void __launch_bounds__(128)
gat_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
               const float* __restrict__ h, int d, const float* __restrict__ s1, const float* __restrict__ s2,
               float alpha, const float* __restrict__ edge_scale, float* __restrict__ out,
               float* __restrict__ wsum, GatSegments sg, float* __restrict__ seg_scratch) {
  const int lane = threadIdx.x & 31;
  const int64_t item = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (item >= n_rows + sg.n_seg) return;
  int64_t row;
  int b, e;
  const bool is_seg = item >= n_rows;
  if (!is_seg) {
    row = item;
    b = rowptr[row]; e = rowptr[row + 1];
    if (e - b > sg.thresh) return;
  } else {
    row = sg.seg_row[item - n_rows];
    b = sg.seg_begin[item - n_rows]; e = sg.seg_end[item - n_rows];
  }
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
    for (int t = 0; t < cnt; t += U) {
      float x[U][S], w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
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
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(w[u], x[u][s], acc[s]);
    }
  }
  wtot = warp_sum(wtot);
  if (is_seg) {
    const int64_t sidx = item - n_rows;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int k = lane + 32 * s;
      if (k < d) seg_scratch[sidx * d + k] = acc[s];
    }
    if (lane == 0) seg_scratch[sg.n_seg * d + sidx] = wtot;
    return;
  }
  const float inv = wtot > 0.f ? 1.0f / wtot : 0.f;      // isolated row: no neighbours -> 0 (the reference gives NaN)
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int k = lane + 32 * s;
    if (k < d) out[row * d + k] = acc[s] * inv;
  }
  if (lane == 0 && wsum) wsum[row] = wtot;
}

__global__ void gat_long_finish_kernel(const int32_t* __restrict__ long_rows, const int32_t* __restrict__ long_first,
                                       int64_t n_long, int64_t n_seg, const float* __restrict__ seg_scratch, int d,
                                       float* __restrict__ out, float* __restrict__ wsum) {
  const int64_t r = blockIdx.x;
  if (r >= n_long) return;
  const int row = long_rows[r];
  const int s0 = long_first[r], s1 = long_first[r + 1];
  float W = 0.f;
  for (int s = s0; s < s1; ++s) W += seg_scratch[n_seg * d + s];
  const float inv = W > 0.f ? 1.0f / W : 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = 0.f;
    for (int s = s0; s < s1; ++s) acc += seg_scratch[(int64_t)s * d + k];
    out[(int64_t)row * d + k] = acc * inv;
  }
  if (threadIdx.x == 0 && wsum) wsum[row] = W;
}

template <int S, int U>
__global__ void __launch_bounds__(128)
gat_bwd_edges_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                     const float* __restrict__ h, int d, const float* __restrict__ s1, const float* __restrict__ s2,
                     float alpha, const float* __restrict__ edge_scale, const float* __restrict__ y,
                     const float* __restrict__ wsum, const float* __restrict__ dy, float* __restrict__ p_edge,
                     float* __restrict__ ds1, float* __restrict__ ds2, GatSegments sg) {
  const int lane = threadIdx.x & 31;
  const int64_t item = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (item >= n_rows + sg.n_seg) return;
  int64_t row;
  int b, e;
  const bool is_seg = item >= n_rows;
  if (!is_seg) {
    row = item;
    b = rowptr[row]; e = rowptr[row + 1];
    if (e - b > sg.thresh) return;                       // hub row: its segments accumulate ds1 atomically
  } else {
    row = sg.seg_row[item - n_rows];
    b = sg.seg_begin[item - n_rows]; e = sg.seg_end[item - n_rows];
  }
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
    for (int t = 0; t < cnt; t += U) {
      float x[U][S];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(0xffffffffu, my_col, (t + u) & 31);
        const bool live = t + u < cnt;
        const float* rp = h + (int64_t)c * d;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int k = lane + 32 * s;
          x[u][s] = (live && k < d) ? __ldg(rp + k) : 0.f;
        }
      }
      float part[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) acc = fmaf(g[s], x[u][s], acc);
        part[u] = acc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < U; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lane == t + u) my_dot = part[u];
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
  if (lane == 0) {
    if (is_seg) atomicAdd(ds1 + row, ds1_acc); else ds1[row] = ds1_acc;
  }
}

__global__ void permute_edges_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                                     float* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

template <int S>
static int gat_fwd_s(const int32_t* rowptr, const int32_t* col, int64_t n, const float* h, int d, const float* s1,
                     const float* s2, float alpha, const float* edge_scale, float* out, float* wsum, GatSegments sg,
                     float* seg_scratch, cudaStream_t s) {
  constexpr int U = S <= 4 ? 4 : 2;
  gat_fwd_kernel<S, U><<<(unsigned)ceil_div(n + sg.n_seg, 4), 128, 0, s>>>(rowptr, col, n, h, d, s1, s2, alpha,
                                                                            edge_scale, out, wsum, sg, seg_scratch);
  EG_LAUNCHED();
  return EG_OK;
}
template <int S>
static int gat_bwd_s(const int32_t* rowptr, const int32_t* col, int64_t n, const float* h, int d, const float* s1,
                     const float* s2, float alpha, const float* edge_scale, const float* y, const float* wsum,
                     const float* dy, float* p_edge, float* ds1, float* ds2, GatSegments sg, cudaStream_t s) {
  constexpr int U = S <= 4 ? 4 : 2;
  gat_bwd_edges_kernel<S, U><<<(unsigned)ceil_div(n + sg.n_seg, 4), 128, 0, s>>>(
      rowptr, col, n, h, d, s1, s2, alpha, edge_scale, y, wsum, dy, p_edge, ds1, ds2, sg);
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
               int long_row_threshold, const int32_t* seg_row, const int32_t* seg_begin, const int32_t* seg_end,
               int64_t n_seg, const int32_t* long_rows, const int32_t* long_first, int64_t n_long,
               float* seg_scratch, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || d <= 0 || n_seg < 0 || n_long < 0) return EG_ERR_INVALID;
  if (n_rows == 0) return EG_OK;
  if (!rowptr || !h || !s1 || !s2 || !out) return EG_ERR_INVALID;
  if (n_seg > 0 && (!seg_row || !seg_begin || !seg_end || !long_rows || !long_first || !seg_scratch || n_long == 0))
    return EG_ERR_INVALID;
  if (long_row_threshold <= 0 || n_seg == 0) long_row_threshold = 0x7fffffff;
  cudaStream_t s = as_stream(stream_);
  GatSegments sg{long_row_threshold, seg_row, seg_begin, seg_end, n_seg};
  int rc = [&]() -> int { EG_GAT_DISPATCH(gat_fwd_s, rowptr, col, n_rows, h, d, s1, s2, alpha, edge_scale, out, wsum,
                                          sg, seg_scratch, s); }();
  if (rc != EG_OK) return rc;
  if (n_seg > 0) {
    gat_long_finish_kernel<<<(unsigned)n_long, 128, 0, s>>>(long_rows, long_first, n_long, n_seg, seg_scratch, d, out,
                                                            wsum);
    EG_LAUNCHED();
  }
  return EG_OK;
}

int eg_gat_bwd_edges(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols, const float* h, int d,
                     const float* s1, const float* s2, float alpha, const float* edge_scale, const float* y,
                     const float* wsum, const float* dy, float* p_edge, float* ds1, float* ds2,
                     int long_row_threshold, const int32_t* seg_row, const int32_t* seg_begin,
                     const int32_t* seg_end, int64_t n_seg, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || d <= 0 || n_seg < 0) return EG_ERR_INVALID;
  if (n_rows == 0) return EG_OK;
  if (!rowptr || !h || !s1 || !s2 || !y || !wsum || !dy || !p_edge || !ds1 || !ds2) return EG_ERR_INVALID;
  if (n_seg > 0 && (!seg_row || !seg_begin || !seg_end)) return EG_ERR_INVALID;
  if (long_row_threshold <= 0 || n_seg == 0) long_row_threshold = 0x7fffffff;
  cudaStream_t s = as_stream(stream_);
  GatSegments sg{long_row_threshold, seg_row, seg_begin, seg_end, n_seg};
  EG_CUDA(cudaMemsetAsync(ds2, 0, sizeof(float) * (size_t)n_cols, s));
  if (n_seg > 0) EG_CUDA(cudaMemsetAsync(ds1, 0, sizeof(float) * (size_t)n_rows, s));
  EG_GAT_DISPATCH(gat_bwd_s, rowptr, col, n_rows, h, d, s1, s2, alpha, edge_scale, y, wsum, dy, p_edge, ds1, ds2, sg,
                  s);
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
