// (b) Backward of the fused OT loss: d<P, C(X, Y)>/dX with the plan held fixed, cost never materialised.
//
// The reference differentiates  sum(newT * M)  with the Sinkhorn plan computed on M.detach()
// (models/models_ea.py:218-224: the gradient flows through the cost matrix only).  For the plan
// P_ij = exp(f_i + g_j - c_ij / reg) that the fused solver leaves as potentials this is
//     dX_i = sum_j P_ij dc_ij/dx_i
//       L2       : c = |x - y|        dc/dx = (x - y) / c            W = P / c     dX_i = (sum_j W_ij) x_i - sum_j W_ij y_j
//       sqeuclid : c = |x - y|^2      dc/dx = 2 (x - y)              W = 2 P       (same form)
//       cosine   : c = 1 - x.y/(|x||y|)                              W = P / (|x||y|)
//                                     dX_i = -sum_j W_ij y_j + (sum_j W_ij x_i.y_j) x_i / |x_i|^2
// i.e. a second contraction G = W · Y ([I, J] x [J, d]) whose left operand is produced tile by tile from the first
// one (the cost tile), exactly like the P·V product of an attention kernel.  Both contractions run on the CUDA
// cores here (64 x 64 tiles; the tcgen05 kernels cover the solve, which dominates: one backward = 1.5 sweeps of
// work).  dY is the same call with the roles of X and Y (and of f and g) exchanged.
// Column ranges are split over grid.y; partial G / s / t go to a [splits, ...] scratch and are summed in split
// order (deterministic).
#include <math_constants.h>

#include "common.cuh"

namespace eg {

constexpr int kGT = 64;      // tile edge
constexpr int kGK = 16;      // k-chunk of the cost contraction / j-chunk of the W·B contraction
constexpr int kGPad = 68;

// NG = column groups of 64 features (d <= 64 * NG)
template <int NG>
__global__ void __launch_bounds__(256, (NG > 2) ? 1 : 2)
plan_grad_simt_kernel(int cost, const float* __restrict__ A, int64_t nA, const float* __restrict__ B, int64_t nB,
                      int d, const float* __restrict__ normA, const float* __restrict__ normB, float inv_reg,
                      const float* __restrict__ potA, const float* __restrict__ potB, int64_t cols_per_split,
                      float* __restrict__ G_part /* [splits, nA, d] */, float* __restrict__ s_part /* [splits, nA] */,
                      float* __restrict__ t_part /* [splits, nA] (cosine) */) {
  __shared__ __align__(16) float As[kGK][kGPad];
  __shared__ __align__(16) float Bs[kGK][kGPad];
  __shared__ __align__(16) float Ws[kGT][kGPad];            // [j][row]: W tile, transposed for the second product
  __shared__ __align__(16) float B2[kGK][64 * NG + 4];      // [j in chunk][feature]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * kGT;
  const int64_t jbeg = (int64_t)blockIdx.y * cols_per_split;
  const int64_t jend = min(nB, jbeg + cols_per_split);
  float na[4], fa[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t i = i0 + 4 * ty + a;
    na[a] = (i < nA) ? normA[i] : 0.f;
    fa[a] = (i < nA) ? potA[i] : -CUDART_INF_F;
  }
  float4 gacc[4][NG];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int g = 0; g < NG; ++g) gacc[a][g] = make_float4(0.f, 0.f, 0.f, 0.f);
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, tsum[4] = {0.f, 0.f, 0.f, 0.f};

  for (int64_t j0 = jbeg; j0 < jend; j0 += kGT) {
    // ---- first contraction: dot products of the 64 x 64 tile ------------------------------------------------
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < d; k0 += kGK) {
      const int kk = threadIdx.x & 15, rb = threadIdx.x >> 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = rb + 16 * r;
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
      for (int k = 0; k < kGK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
        const float a_[4] = {av.x, av.y, av.z, av.w};
        const float b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a_[a], b_[b], acc[a][b]);
      }
      __syncthreads();
    }
    // ---- W tile ------------------------------------------------------------------------------------------------
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = j0 + 4 * tx + b;
      const bool okb = j < jend;
      const float nb = okb ? normB[j] : 0.f;
      const float gb = okb ? potB[j] : -CUDART_INF_F;
      float w4[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int64_t i = i0 + 4 * ty + a;
        const bool live = okb && i < nA;
        float w = 0.f;
        if (live) {
          if (cost == EG_COST_COSINE) {
            const float den = fmaxf(na[a], 1e-8f) * fmaxf(nb, 1e-8f);
            const float c = 1.0f - acc[a][b] / den;
            const float p = __expf(fa[a] + gb - c * inv_reg);
            w = p / den;
            tsum[a] = fmaf(w, acc[a][b], tsum[a]);
          } else {
            float sq = fmaxf(fmaf(-2.0f, acc[a][b], na[a] + nb), 0.0f);
            if (sq < 0.25f * (na[a] + nb)) {            // close pair: the norm expansion cancels — exact from the rows
              const float* ar = A + i * d;
              const float* br = B + j * d;
              float e = 0.f;
              for (int k = 0; k < d; ++k) { const float df = __ldg(ar + k) - __ldg(br + k); e = fmaf(df, df, e); }
              sq = e;
            }
            if (cost == EG_COST_L2) {
              const float c = sqrtf(sq);
              const float p = __expf(fa[a] + gb - c * inv_reg);
              w = (c > 0.f) ? p / c : 0.f;               // subgradient 0 at coincident points (as torch.cdist)
            } else {
              w = 2.0f * __expf(fa[a] + gb - sq * inv_reg);
            }
          }
        }
        ssum[a] += w;
        w4[a] = w;
      }
      *reinterpret_cast<float4*>(&Ws[4 * tx + b][4 * ty]) = make_float4(w4[0], w4[1], w4[2], w4[3]);
    }
    // ---- second contraction: G[rows, :] += W[rows, j] * B[j, :] over the tile's 64 columns, 16 at a time ----------
    for (int jc = 0; jc < kGT; jc += kGK) {
      __syncthreads();                                      // Ws complete (first chunk) / B2 free (later chunks)
      for (int e = threadIdx.x; e < kGK * 16 * NG; e += 256) {
        const int jr = e / (16 * NG), c4 = e % (16 * NG);
        const int64_t j = j0 + jc + jr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < jend) {
          const float* src = B + j * d + 4 * c4;
          if (4 * c4 + 3 < d && (d & 3) == 0) v = __ldg(reinterpret_cast<const float4*>(src));
          else {
            if (4 * c4 + 0 < d) v.x = __ldg(src + 0);
            if (4 * c4 + 1 < d) v.y = __ldg(src + 1);
            if (4 * c4 + 2 < d) v.z = __ldg(src + 2);
            if (4 * c4 + 3 < d) v.w = __ldg(src + 3);
          }
        }
        *reinterpret_cast<float4*>(&B2[jr][4 * c4]) = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int jr = 0; jr < kGK; ++jr) {
        const float4 wv = *reinterpret_cast<const float4*>(&Ws[jc + jr][4 * ty]);
        const float w_[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float4 bv = *reinterpret_cast<const float4*>(&B2[jr][64 * g + 4 * tx]);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            gacc[a][g].x = fmaf(w_[a], bv.x, gacc[a][g].x); gacc[a][g].y = fmaf(w_[a], bv.y, gacc[a][g].y);
            gacc[a][g].z = fmaf(w_[a], bv.z, gacc[a][g].z); gacc[a][g].w = fmaf(w_[a], bv.w, gacc[a][g].w);
          }
        }
      }
    }
    __syncthreads();                                        // Ws / B2 reusable by the next tile
  }
  // ---- partial outputs of this column split -----------------------------------------------------------------------
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t i = i0 + 4 * ty + a;
    float s = ssum[a], t = tsum[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
    if (i >= nA) continue;
    if (tx == 0) {
      s_part[(int64_t)blockIdx.y * nA + i] = s;
      if (t_part) t_part[(int64_t)blockIdx.y * nA + i] = t;
    }
    float* dst = G_part + ((int64_t)blockIdx.y * nA + i) * d;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int c = 64 * g + 4 * tx;
      if (c + 0 < d) dst[c + 0] = gacc[a][g].x;
      if (c + 1 < d) dst[c + 1] = gacc[a][g].y;
      if (c + 2 < d) dst[c + 2] = gacc[a][g].z;
      if (c + 3 < d) dst[c + 3] = gacc[a][g].w;
    }
  }
}

// dA_i = scale * ( coef_i * a_i - sum_splits G ),  coef_i = sum_j W_ij  (L2 / sqeuclid)  or  t_i / |a_i|^2 (cosine)
__global__ void plan_grad_finish_kernel(int cost, const float* __restrict__ A, int64_t nA, int d,
                                        const float* __restrict__ normA, const float* __restrict__ G_part,
                                        const float* __restrict__ s_part, const float* __restrict__ t_part,
                                        int splits, float scale, float* __restrict__ dA) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= nA * d) return;
  const int64_t i = idx / d;
  float g = 0.f, coef = 0.f;
  for (int sp = 0; sp < splits; ++sp) {
    g += G_part[(int64_t)sp * nA * d + idx];
    coef += (cost == EG_COST_COSINE) ? t_part[(int64_t)sp * nA + i] : s_part[(int64_t)sp * nA + i];
  }
  if (cost == EG_COST_COSINE) { const float n = fmaxf(normA[i], 1e-8f); coef /= n * n; }
  dA[idx] = scale * (coef * A[idx] - g);
}

static int grad_splits(int64_t nA, int64_t nB) {
  const int64_t row_blocks = ceil_div(nA, (int64_t)kGT);
  int64_t want = ceil_div((int64_t)kNumSMs * 2, row_blocks);
  want = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, ceil_div(nB, (int64_t)kGT)), 16));
  return (int)want;
}

}  // namespace eg

extern "C" {

size_t eg_plan_grad_fused_workspace_bytes(int64_t nA, int64_t nB, int d) {
  if (nA <= 0 || nB <= 0 || d <= 0) return 0;
  const int sp = eg::grad_splits(nA, nB);
  return eg::align_up(sizeof(float) * (size_t)sp * (size_t)nA * (size_t)d) + 2 * eg::align_up(sizeof(float) * (size_t)sp * (size_t)nA);
}

int eg_plan_grad_fused(int cost, const float* A, int64_t nA, const float* B, int64_t nB, int d, const float* normA,
                       const float* normB, float inv_reg, const float* f, const float* g, float scale, void* ws,
                       size_t ws_bytes, float* dA, eg_stream_t stream_) {
  using namespace eg;
  if (nA < 0 || nB < 0 || d <= 0 || cost < 0 || cost > 2) return EG_ERR_INVALID;
  if (d > 64 * 5) return EG_ERR_UNSUPPORTED;
  if (nA == 0) return EG_OK;
  if (!A || !normA || !f || !dA || !ws) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (nB == 0) { EG_CUDA(cudaMemsetAsync(dA, 0, sizeof(float) * (size_t)nA * (size_t)d, s)); return EG_OK; }
  if (!B || !normB || !g) return EG_ERR_INVALID;
  if (ws_bytes < eg_plan_grad_fused_workspace_bytes(nA, nB, d)) return EG_ERR_WORKSPACE;
  const int sp = grad_splits(nA, nB);
  const int64_t cols_per = ceil_div(ceil_div(nB, (int64_t)sp), (int64_t)kGT) * kGT;
  const int splits = (int)ceil_div(nB, cols_per);
  float* G_part = reinterpret_cast<float*>(ws);
  float* s_part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up(sizeof(float) * (size_t)sp * (size_t)nA * (size_t)d));
  float* t_part = reinterpret_cast<float*>(reinterpret_cast<char*>(s_part) + align_up(sizeof(float) * (size_t)sp * (size_t)nA));
  const int64_t gx = ceil_div(nA, (int64_t)kGT);
  if (gx > 0x7fffffff) return EG_ERR_UNSUPPORTED;
  dim3 grid((unsigned)gx, (unsigned)splits);
  if (d <= 128)
    plan_grad_simt_kernel<2><<<grid, 256, 0, s>>>(cost, A, nA, B, nB, d, normA, normB, inv_reg, f, g, cols_per, G_part,
                                                 s_part, t_part);
  else
    plan_grad_simt_kernel<5><<<grid, 256, 0, s>>>(cost, A, nA, B, nB, d, normA, normB, inv_reg, f, g, cols_per, G_part,
                                                 s_part, t_part);
  EG_LAUNCHED();
  plan_grad_finish_kernel<<<(unsigned)ceil_div(nA * (int64_t)d, (int64_t)256), 256, 0, s>>>(cost, A, nA, d, normA, G_part,
                                                                                          s_part, t_part, splits, scale, dA);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
