// Margin-based L1 ranking loss with hard negatives, fused gather + |·|₁ + hinge.
//
// Replaces the four [t·k, d] gathers + abs/sum/relu chain of models/models_ea.py:103-123 (EAModel.get_loss)
// and :185-204 (UEAModel.get_loss):
//     A_p   = ||out[left_p] - out[right_p]||_1
//     L     = sum_{p,q} relu(A_p + gamma - ||out[nl_pq] - out[nr_pq]||_1)
//           + sum_{p,q} relu(A_p + gamma - ||out[n2l_pq] - out[n2r_pq]||_1)          / (2 t k)
// One warp owns one anchor pair p and walks its 2k negative pairs; nothing of size [t·k, d] is ever
// materialised (the reference builds four 675 MB gathers at t·k = 562,500, d = 300).  HBM/L2-bound:
// 2·t·k·2·d·4 bytes of row reads.  Backward recomputes the distances, scatters sign(x_i - x_j) with
// float atomics, and keeps the anchor rows (which repeat k times) in registers.
#include "common.cuh"

namespace eg {

constexpr int kMlWarps = 8;
constexpr int kMlMaxVec = 4;     // float4 per lane -> d <= 512

__device__ __forceinline__ float l1_rows(const float* __restrict__ a, const float* __restrict__ b, int d, int lane) {
  float acc = 0.f;
  if ((d & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int k = lane; k < d / 4; k += 32) {
      float4 x = __ldg(a4 + k), y = __ldg(b4 + k);
      acc += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
    }
  } else {
    for (int k = lane; k < d; k += 32) acc += fabsf(__ldg(a + k) - __ldg(b + k));
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(kMlWarps * 32)
margin_loss_fwd_kernel(const float* __restrict__ out, int d, const int64_t* __restrict__ left,
                       const int64_t* __restrict__ right, const int64_t* __restrict__ nl,
                       const int64_t* __restrict__ nr, const int64_t* __restrict__ n2l,
                       const int64_t* __restrict__ n2r, int64_t t, int k, float gamma, double* __restrict__ loss) {
  __shared__ double red[kMlWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * kMlWarps + warp;
  double mine = 0.0;
  if (p < t) {
    const float A = l1_rows(out + left[p] * d, out + right[p] * d, d, lane);
    const float D = A + gamma;
    for (int q = 0; q < k; ++q) {
      const int64_t e = p * k + q;
      const float B1 = l1_rows(out + nl[e] * d, out + nr[e] * d, d, lane);
      const float B2 = l1_rows(out + n2l[e] * d, out + n2r[e] * d, d, lane);
      mine += (double)fmaxf(D - B1, 0.f) + (double)fmaxf(D - B2, 0.f);
    }
  }
  if (lane == 0) red[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kMlWarps; ++w) s += red[w];
    if (s != 0.0) atomicAdd(loss, s);
  }
}

// grad[row_i] += c * sign(x_i - x_j), grad[row_j] -= c * sign(x_i - x_j); rows equal to `keep_row`
// accumulate into the caller's registers instead of global atomics.
__device__ __forceinline__ void scatter_sign(const float* __restrict__ out, float* __restrict__ grad, int d, int lane,
                                             int64_t i, int64_t j, float c, int64_t keep_a, float* acc_a,
                                             int64_t keep_b, float* acc_b) {
  const float* xi = out + i * d;
  const float* xj = out + j * d;
  int slot = 0;
  for (int kk = lane; kk < d; kk += 32, ++slot) {
    const float df = __ldg(xi + kk) - __ldg(xj + kk);
    const float s = (df > 0.f) ? c : ((df < 0.f) ? -c : 0.f);
    if (s != 0.f) {
      if (i == keep_a) acc_a[slot] += s; else if (i == keep_b) acc_b[slot] += s; else atomicAdd(grad + i * d + kk, s);
      if (j == keep_a) acc_a[slot] -= s; else if (j == keep_b) acc_b[slot] -= s; else atomicAdd(grad + j * d + kk, -s);
    }
  }
}

// WIDE: rows wider than the per-lane register accumulators (d > 512) — the anchor rows go through the
// global atomics like every other row.
template <bool WIDE>
__global__ void __launch_bounds__(kMlWarps * 32)
margin_loss_bwd_kernel(const float* __restrict__ out, int d, const int64_t* __restrict__ left,
                       const int64_t* __restrict__ right, const int64_t* __restrict__ nl,
                       const int64_t* __restrict__ nr, const int64_t* __restrict__ n2l,
                       const int64_t* __restrict__ n2r, int64_t t, int k, float gamma, float scale,
                       float* __restrict__ grad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = (int64_t)blockIdx.x * kMlWarps + warp;
  if (p >= t) return;
  const int64_t lp = left[p], rp = right[p];
  const int64_t keep_l = WIDE ? -1 : lp, keep_r = WIDE ? -1 : rp;
  float acc_l[kMlMaxVec * 4], acc_r[kMlMaxVec * 4];       // this lane's columns of the two anchor rows (d <= 512)
#pragma unroll
  for (int s = 0; s < kMlMaxVec * 4; ++s) { acc_l[s] = 0.f; acc_r[s] = 0.f; }
  const float A = l1_rows(out + lp * d, out + rp * d, d, lane);
  const float D = A + gamma;
  int active = 0;
  for (int q = 0; q < k; ++q) {
    const int64_t e = p * k + q;
    const float B1 = l1_rows(out + nl[e] * d, out + nr[e] * d, d, lane);
    if (D - B1 > 0.f) { ++active; scatter_sign(out, grad, d, lane, nl[e], nr[e], -scale, keep_l, acc_l, keep_r, acc_r); }
    const float B2 = l1_rows(out + n2l[e] * d, out + n2r[e] * d, d, lane);
    if (D - B2 > 0.f) { ++active; scatter_sign(out, grad, d, lane, n2l[e], n2r[e], -scale, keep_l, acc_l, keep_r, acc_r); }
  }
  if (active) scatter_sign(out, grad, d, lane, lp, rp, scale * (float)active, keep_l, acc_l, keep_r, acc_r);
  if (WIDE) return;
  int slot = 0;
  for (int kk = lane; kk < d; kk += 32, ++slot) {
    if (acc_l[slot] != 0.f) atomicAdd(grad + lp * d + kk, acc_l[slot]);
    if (acc_r[slot] != 0.f) atomicAdd(grad + rp * d + kk, acc_r[slot]);
  }
}

}  // namespace eg

extern "C" {

int eg_margin_loss_fwd(const float* out, int64_t n, int d, const int64_t* left, const int64_t* right,
                       const int64_t* nl, const int64_t* nr, const int64_t* n2l, const int64_t* n2r, int64_t t, int k,
                       float gamma, double* loss_sum, eg_stream_t stream_) {
  using namespace eg;
  (void)n;
  if (t < 0 || k <= 0 || d <= 0) return EG_ERR_INVALID;
  if (!loss_sum) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  EG_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
  if (t == 0) return EG_OK;
  if (!out || !left || !right || !nl || !nr || !n2l || !n2r) return EG_ERR_INVALID;
  margin_loss_fwd_kernel<<<(unsigned)ceil_div(t, kMlWarps), kMlWarps * 32, 0, s>>>(out, d, left, right, nl, nr, n2l,
                                                                                  n2r, t, k, gamma, loss_sum);
  EG_LAUNCHED();
  return EG_OK;
}

int eg_margin_loss_bwd(const float* out, int64_t n, int d, const int64_t* left, const int64_t* right,
                       const int64_t* nl, const int64_t* nr, const int64_t* n2l, const int64_t* n2r, int64_t t, int k,
                       float gamma, float scale, float* grad, eg_stream_t stream_) {
  using namespace eg;
  if (t < 0 || k <= 0 || d <= 0 || n < 0) return EG_ERR_INVALID;
  if (t == 0) return EG_OK;
  if (!out || !grad || !left || !right || !nl || !nr || !n2l || !n2r) return EG_ERR_INVALID;
  if (d > 32 * kMlMaxVec * 4)
    margin_loss_bwd_kernel<true><<<(unsigned)ceil_div(t, kMlWarps), kMlWarps * 32, 0, as_stream(stream_)>>>(
        out, d, left, right, nl, nr, n2l, n2r, t, k, gamma, scale, grad);
  else
    margin_loss_bwd_kernel<false><<<(unsigned)ceil_div(t, kMlWarps), kMlWarps * 32, 0, as_stream(stream_)>>>(
        out, d, left, right, nl, nr, n2l, n2r, t, k, gamma, scale, grad);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
