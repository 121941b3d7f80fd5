// (b) FUSED log-domain Sinkhorn half-sweep: cost tiles are recomputed from the two
// embedding sets and reduced straight into the row log-sum-exp — the I×J cost of
// models/models_ea.py:218 (torch.cdist) and the Gibbs kernel of
// utils/ot_loss.py:41-47 are never written to memory.
//
// This file holds the SIMT (fp32 FMA) tile path, the operand preparation shared
// with the tcgen05 path (row norms, 3xTF32 hi/lo split) and the dispatch; the
// TMA + tcgen05 tile path lives in sinkhorn_tc.cu.
//
//   lse[i] = log sum_j exp(pot[j] - cost(A_i, B_j) * inv_reg)
//   cost:  L2        sqrt(max(|a|^2 + |b|^2 - 2 a·b, 0))      (torch.cdist p=2 formula)
//          SQEUCLID  max(|a|^2 + |b|^2 - 2 a·b, 0)            (cderivation.py:14-26, p=2)
//          COSINE    1 - a·b / (max(|a|,eps) max(|b|,eps))    (cderivation.py:44-61)
#include <math_constants.h>

#include "common.cuh"

namespace eg {

int lse_fused_tc(int cost, int64_t nA, int64_t nB, int d, const float* normA, const float* normB, float inv_reg,
                 const float* pot_in, const float* logw, float* pot_out, float* lse_out, const float* A_hi,
                 const float* A_lo, const float* B_hi, const float* B_lo, const float* A_raw, const float* B_raw,
                 void* ws, size_t ws_bytes, cudaStream_t s);
size_t lse_fused_tc_workspace(int64_t nA, int64_t nB, int d);
int plan_fused_tc(int cost, int64_t nA, int64_t nB, int d, const float* normA, const float* normB, float inv_reg,
                  const float* f, const float* g, double* loss, float* row_sum, const float* A_hi,
                  const float* A_lo, const float* B_hi, const float* B_lo, const float* A_raw, const float* B_raw,
                  cudaStream_t s);

int gemm_nt_tc(const float* A1_hi, const float* A1_lo, int k1p, const float* A2_hi, const float* A2_lo, int k2p,
               int64_t m, const float* B_hi, const float* B_lo, int64_t n, const float* bias, float* out1, int64_t ld1,
               int64_t n1, float* out2, int64_t ld2, int flush, cudaStream_t s);

constexpr int kFT = 64;    // tile edge
constexpr int kFK = 16;    // k-chunk
constexpr int kFPad = 68;  // [k][i] row stride in floats (272 B: 16-B aligned, skews banks)

// Close pairs (squared distance below a quarter of |a|^2 + |b|^2, i.e. cosine > 0.75) are re-evaluated from the
// rows themselves: the norm expansion cancels catastrophically exactly where the log-sum-exp is decided.
constexpr float kNearFracSimt = 0.25f;
__device__ __noinline__ float exact_pair_simt(const float* __restrict__ a, const float* __restrict__ b, int d, int want_dot) {
  float acc = 0.f;
  for (int k = 0; k < d; ++k) {
    const float x = __ldg(a + k), y = __ldg(b + k);
    acc = want_dot ? fmaf(x, y, acc) : fmaf(x - y, x - y, acc);
  }
  return acc;
}
__device__ __forceinline__ float cost_from_dot(int cost, float dot, float na, float nb, const float* a_row,
                                               const float* b_row, int d) {
  if (cost == EG_COST_COSINE) {
    float den = fmaxf(na, 1e-8f) * fmaxf(nb, 1e-8f);
    if (a_row && dot > (1.f - 2.f * kNearFracSimt) * den) dot = exact_pair_simt(a_row, b_row, d, 1);
    return 1.0f - dot / den;
  }
  float sq = fmaxf(fmaf(-2.0f, dot, na + nb), 0.0f);
  if (a_row && sq < kNearFracSimt * (na + nb)) sq = exact_pair_simt(a_row, b_row, d, 0);
  return cost == EG_COST_L2 ? sqrtf(sq) : sq;
}

struct OnlineLse {
  float m, s;
  __device__ __forceinline__ void init() { m = -CUDART_INF_F; s = 0.f; }
  __device__ __forceinline__ void push4(float z0, float z1, float z2, float z3) {
    float mx = fmaxf(fmaxf(z0, z1), fmaxf(z2, z3));
    if (mx > m) { s *= __expf(m - mx); m = mx; }
    if (m > -CUDART_INF_F) s += __expf(z0 - m) + __expf(z1 - m) + __expf(z2 - m) + __expf(z3 - m);
  }
  __device__ __forceinline__ void merge(float om, float os) {
    float mx = fmaxf(m, om);
    if (mx > -CUDART_INF_F) { s = s * __expf(m - mx) + os * __expf(om - mx); m = mx; }
  }
};

// MODE 0: partial (max, sum) per row over this CTA's column range -> part_m/part_s[split][row]
// MODE 1: plan statistics: loss += sum p*c, row_sum[i] += sum_j p, optional P
template <int MODE>
__global__ void __launch_bounds__(256, 2)
fused_simt_kernel(int cost, const float* __restrict__ A, int64_t nA, const float* __restrict__ B, int64_t nB,
                  int d, const float* __restrict__ normA, const float* __restrict__ normB, float inv_reg,
                  const float* __restrict__ potA /* f, MODE 1 */, const float* __restrict__ potB,
                  int64_t cols_per_split, float* __restrict__ part_m, float* __restrict__ part_s,
                  float* __restrict__ P, int64_t ldP, double* __restrict__ loss, float* __restrict__ row_sum) {
  __shared__ __align__(16) float As[kFK][kFPad];
  __shared__ __align__(16) float Bs[kFK][kFPad];
  __shared__ double red[8];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * kFT;
  const int64_t jbeg = (int64_t)blockIdx.y * cols_per_split;
  const int64_t jend = min(nB, jbeg + cols_per_split);
  float na[4], fa[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int64_t i = i0 + 4 * ty + a;
    na[a] = (i < nA) ? normA[i] : 0.f;
    fa[a] = (MODE == 1 && i < nA) ? potA[i] : 0.f;
  }
  OnlineLse lse[4];
  float rsum[4] = {0.f, 0.f, 0.f, 0.f};
  double my_loss = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) lse[a].init();

  for (int64_t j0 = jbeg; j0 < jend; j0 += kFT) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < d; k0 += kFK) {
      // stage 64 rows x 16 k of A and B: thread -> (row = tid/16 + 16 r, k = tid%16)
      const int kk = threadIdx.x & 15, rb = threadIdx.x >> 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int row = rb + 16 * r;
        float av = 0.f, bv = 0.f;
        if (k0 + kk < d) {
          if (i0 + row < nA) av = __ldg(A + (i0 + row) * d + k0 + kk);
          if (j0 + row < jend) bv = __ldg(B + (j0 + row) * d + k0 + kk);
        }
        As[kk][row] = av;
        Bs[kk][row] = bv;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kFK; ++k) {
        float4 av = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
        const float a_[4] = {av.x, av.y, av.z, av.w};
        const float b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a_[a], b_[b], acc[a][b]);
      }
      __syncthreads();
    }
    // epilogue on the 4x4 micro-tile
    float nb[4], gb[4];
    bool okb[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int64_t j = j0 + 4 * tx + b;
      okb[b] = j < jend;
      nb[b] = okb[b] ? normB[j] : 0.f;
      gb[b] = okb[b] ? potB[j] : 0.f;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float z[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t gi = i0 + 4 * ty + a, gj = j0 + 4 * tx + b;
        const bool live = okb[b] && gi < nA;
        float c = cost_from_dot(cost, acc[a][b], na[a], nb[b], live ? A + gi * d : nullptr, live ? B + gj * d : nullptr, d);
        if (MODE == 0) {
          z[b] = okb[b] ? fmaf(-c, inv_reg, gb[b]) : -CUDART_INF_F;
        } else {
          int64_t i = i0 + 4 * ty + a;
          float p = (okb[b] && i < nA) ? __expf(fa[a] + gb[b] - c * inv_reg) : 0.f;
          rsum[a] += p;
          my_loss += (double)p * (double)c;
          if (P && okb[b] && i < nA) P[i * ldP + j0 + 4 * tx + b] = p;
        }
      }
      if (MODE == 0) lse[a].push4(z[0], z[1], z[2], z[3]);
    }
  }
  // reduce over the 16 threads (tx) that share rows: lanes differ in the low 4 bits
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (MODE == 0) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, lse[a].m, o);
        float os = __shfl_xor_sync(0xffffffffu, lse[a].s, o);
        lse[a].merge(om, os);
      }
      int64_t i = i0 + 4 * ty + a;
      if (tx == 0 && i < nA) {
        part_m[(int64_t)blockIdx.y * nA + i] = lse[a].m;
        part_s[(int64_t)blockIdx.y * nA + i] = lse[a].s;
      }
    } else {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rsum[a] += __shfl_xor_sync(0xffffffffu, rsum[a], o);
      int64_t i = i0 + 4 * ty + a;
      if (row_sum && tx == 0 && i < nA) atomicAdd(&row_sum[i], rsum[a]);
    }
  }
  if (MODE == 1 && loss) {
    my_loss = warp_sum(my_loss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = my_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      atomicAdd(loss, t);
    }
  }
}

__global__ void lse_combine_kernel(const float* __restrict__ part_m, const float* __restrict__ part_s,
                                   int n_split, int64_t n, const float* __restrict__ logw,
                                   float* __restrict__ pot_out, float* __restrict__ lse_out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  OnlineLse acc;
  acc.init();
  for (int sidx = 0; sidx < n_split; ++sidx) acc.merge(part_m[(int64_t)sidx * n + i], part_s[(int64_t)sidx * n + i]);
  float l = acc.s > 0.f ? acc.m + logf(acc.s) : -CUDART_INF_F;
  if (lse_out) lse_out[i] = l;
  if (pot_out) pot_out[i] = (logw ? logw[i] : 0.f) - l;
}

__global__ void row_norms_kernel(const float* __restrict__ A, int64_t n, int d, int squared, float* __restrict__ out) {
  int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n) return;
  const float* row = A + w * d;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) { float v = row[k]; acc = fmaf(v, v, acc); }
  acc = warp_sum(acc);
  if (lane == 0) out[w] = squared ? acc : sqrtf(acc);
}

__global__ void split_tf32_kernel(const float* __restrict__ X, int64_t n, int d, int d_pad, float* __restrict__ hi,
                                  float* __restrict__ lo) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * d_pad) return;
  int64_t r = idx / d_pad;
  int k = (int)(idx - r * d_pad);
  float x = (k < d) ? X[r * d + k] : 0.f;
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
  float h = __uint_as_float(hb);
  float rem = x - h;  // exact in fp32
  uint32_t lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
  hi[idx] = h;
  lo[idx] = __uint_as_float(lb);
}

// Vector path (d % 4 == 0, 16-byte aligned rows): one float4 of the padded row per thread, streaming loads/stores.
__global__ void split_tf32_vec4_kernel(const float4* __restrict__ X, int64_t n, int d4, int dp4, float4* __restrict__ hi,
                                       float4* __restrict__ lo) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * dp4) return;
  const int64_t r = idx / dp4;
  const int k = (int)(idx - r * dp4);
  const float4 x = (k < d4) ? ld_stream_f4(X + r * d4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float xs[4] = {x.x, x.y, x.z, x.w};
  float h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(xs[e]));
    h[e] = __uint_as_float(hb);
    const float rem = xs[e] - h[e];   // exact in fp32
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
    l[e] = __uint_as_float(lb);
  }
  __stcs(hi + idx, make_float4(h[0], h[1], h[2], h[3]));
  __stcs(lo + idx, make_float4(l[0], l[1], l[2], l[3]));
}

static int pick_splits(int64_t nA, int64_t nB) {
  int64_t row_blocks = ceil_div(nA, kFT);
  int64_t want = ceil_div((int64_t)kNumSMs * 2, row_blocks);   // >= 2 CTAs per SM overall
  int64_t max_split = ceil_div(nB, kFT);
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

}  // namespace eg

extern "C" {

int eg_row_norms(const float* A, int64_t n, int d, int squared, float* out, eg_stream_t stream_) {
  using namespace eg;
  if (n < 0 || d <= 0) return EG_ERR_INVALID;
  if (n == 0) return EG_OK;
  if (!A || !out) return EG_ERR_INVALID;
  row_norms_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, as_stream(stream_)>>>(A, n, d, squared, out);
  EG_LAUNCHED();
  return EG_OK;
}

int eg_split_tf32(const float* X, int64_t n, int d, int d_pad, float* hi, float* lo, eg_stream_t stream_) {
  using namespace eg;
  if (n < 0 || d <= 0 || d_pad < d || d_pad % 8 != 0) return EG_ERR_INVALID;
  if (n == 0) return EG_OK;
  if (!X || !hi || !lo) return EG_ERR_INVALID;
  if (d % 4 == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0) {
    split_tf32_vec4_kernel<<<(unsigned)ceil_div(n * (d_pad / 4), 256), 256, 0, as_stream(stream_)>>>(
        reinterpret_cast<const float4*>(X), n, d / 4, d_pad / 4, reinterpret_cast<float4*>(hi),
        reinterpret_cast<float4*>(lo));
  } else {
    split_tf32_kernel<<<(unsigned)ceil_div(n * d_pad, 256), 256, 0, as_stream(stream_)>>>(X, n, d, d_pad, hi, lo);
  }
  EG_LAUNCHED();
  return EG_OK;
}

size_t eg_lse_fused_workspace_bytes(int algo, int64_t nA, int64_t nB, int d) {
  if (nA <= 0 || nB <= 0) return 0;
  if (algo == EG_ALGO_TCGEN05) return eg::lse_fused_tc_workspace(nA, nB, d);
  int splits = eg::pick_splits(nA, nB);
  return 2 * eg::align_up(sizeof(float) * (size_t)splits * (size_t)nA);
}

int eg_lse_fused(int algo, int cost, const float* A, int64_t nA, const float* B, int64_t nB, int d,
                 const float* normA, const float* normB, float inv_reg, const float* pot_in, const float* logw,
                 float* pot_out, float* lse_out, const float* A_hi, const float* A_lo, const float* B_hi,
                 const float* B_lo, void* ws, size_t ws_bytes, eg_stream_t stream_) {
  using namespace eg;
  if (nA < 0 || nB <= 0 || d <= 0 || cost < 0 || cost > 2) return EG_ERR_INVALID;
  if (nA == 0) return EG_OK;
  if (!normA || !normB || !pot_in || (!pot_out && !lse_out) || !ws) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (algo == EG_ALGO_TCGEN05) {
    if (!A_hi || !A_lo || !B_hi || !B_lo) return EG_ERR_INVALID;
    return lse_fused_tc(cost, nA, nB, d, normA, normB, inv_reg, pot_in, logw, pot_out, lse_out, A_hi, A_lo, B_hi,
                        B_lo, A, B, ws, ws_bytes, s);
  }
  if (algo != EG_ALGO_SIMT || !A || !B) return EG_ERR_INVALID;
  int splits = pick_splits(nA, nB);
  size_t half = align_up(sizeof(float) * (size_t)splits * (size_t)nA);
  if (ws_bytes < 2 * half) return EG_ERR_WORKSPACE;
  float* part_m = reinterpret_cast<float*>(ws);
  float* part_s = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + half);
  int64_t cols_per_split = ceil_div(ceil_div(nB, splits), kFT) * kFT;
  splits = (int)ceil_div(nB, cols_per_split);
  int64_t gx = ceil_div(nA, kFT);
  if (gx > 0x7fffffff) return EG_ERR_UNSUPPORTED;
  dim3 grid((unsigned)gx, (unsigned)splits);
  fused_simt_kernel<0><<<grid, 256, 0, s>>>(cost, A, nA, B, nB, d, normA, normB, inv_reg, nullptr, pot_in,
                                            cols_per_split, part_m, part_s, nullptr, 0, nullptr, nullptr);
  EG_LAUNCHED();
  lse_combine_kernel<<<(unsigned)ceil_div(nA, 256), 256, 0, s>>>(part_m, part_s, splits, nA, logw, pot_out, lse_out);
  EG_LAUNCHED();
  return EG_OK;
}

int eg_gemm_nt_3xtf32(const float* A1_hi, const float* A1_lo, int k1_pad, const float* A2_hi, const float* A2_lo,
                      int k2_pad, int64_t m, const float* B_hi, const float* B_lo, int64_t n, const float* bias,
                      float* out1, int64_t ld1, int64_t n1, float* out2, int64_t ld2, eg_stream_t stream_) {
  using namespace eg;
  if (m < 0 || n <= 0) return EG_ERR_INVALID;
  if (m == 0) return EG_OK;
  if (!A1_hi || !A1_lo || !B_hi || !B_lo || !out1) return EG_ERR_INVALID;
  if (k2_pad > 0 && (!A2_hi || !A2_lo)) return EG_ERR_INVALID;
  return gemm_nt_tc(A1_hi, A1_lo, k1_pad, A2_hi, A2_lo, k2_pad, m, B_hi, B_lo, n, bias, out1, ld1, n1, out2, ld2, 0,
                    as_stream(stream_));
}

int eg_gemm_nt_3xtf32_chained(const float* A1_hi, const float* A1_lo, int k1_pad, const float* A2_hi,
                              const float* A2_lo, int k2_pad, int64_t m, const float* B_hi, const float* B_lo,
                              int64_t n, const float* bias, float* out1, int64_t ld1, int64_t n1, float* out2,
                              int64_t ld2, eg_stream_t stream_) {
  using namespace eg;
  if (m < 0 || n <= 0) return EG_ERR_INVALID;
  if (m == 0) return EG_OK;
  if (!A1_hi || !A1_lo || !B_hi || !B_lo || !out1) return EG_ERR_INVALID;
  if (k2_pad > 0 && (!A2_hi || !A2_lo)) return EG_ERR_INVALID;
  return gemm_nt_tc(A1_hi, A1_lo, k1_pad, A2_hi, A2_lo, k2_pad, m, B_hi, B_lo, n, bias, out1, ld1, n1, out2, ld2, 1,
                    as_stream(stream_));
}

int eg_plan_fused(int algo, int cost, const float* A, int64_t nA, const float* B, int64_t nB, int d,
                  const float* normA, const float* normB, float inv_reg, const float* f, const float* g, float* P,
                  int64_t ldP, double* loss, float* row_sum, const float* A_hi, const float* A_lo,
                  const float* B_hi, const float* B_lo, eg_stream_t stream_) {
  using namespace eg;
  if (nA < 0 || nB <= 0 || d <= 0 || cost < 0 || cost > 2) return EG_ERR_INVALID;
  if (P && ldP < nB) return EG_ERR_INVALID;
  if (nA == 0) return EG_OK;
  if (!normA || !normB || !f || !g) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (loss) EG_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), s));
  if (row_sum) EG_CUDA(cudaMemsetAsync(row_sum, 0, sizeof(float) * (size_t)nA, s));
  if (algo == EG_ALGO_TCGEN05) {
    if (P) return EG_ERR_UNSUPPORTED;   // the plan itself is only ever written by the SIMT path (small sizes)
    if (!A_hi || !A_lo || !B_hi || !B_lo) return EG_ERR_INVALID;
    return plan_fused_tc(cost, nA, nB, d, normA, normB, inv_reg, f, g, loss, row_sum, A_hi, A_lo, B_hi, B_lo, A, B, s);
  }
  if (algo != EG_ALGO_SIMT || !A || !B) return EG_ERR_INVALID;
  int splits = pick_splits(nA, nB);
  int64_t cols_per_split = ceil_div(ceil_div(nB, splits), kFT) * kFT;
  splits = (int)ceil_div(nB, cols_per_split);
  dim3 grid((unsigned)ceil_div(nA, kFT), (unsigned)splits);
  fused_simt_kernel<1><<<grid, 256, 0, s>>>(cost, A, nA, B, nB, d, normA, normB, inv_reg, f, g, cols_per_split,
                                            nullptr, nullptr, P, ldP, loss, row_sum);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
