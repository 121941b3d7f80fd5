// (c) Alignment evaluation: exact fp64 L1 (cityblock) distances and the
// rank / arg-min / top-k reductions the reference does with SciPy + NumPy.
//
// Replaces scipy.spatial.distance.cdist(..., 'cityblock') + argsort loops at
// utils/eval_utils.py:74-89, models/models_ea.py:24-27,149-153.  SciPy widens the
// fp32 embeddings to fp64 and sums |l_k - r_k| for k = 0..d-1 in order; doing
// exactly that on the fp64 pipe (DADD only — there is no multiply, so no FMA
// contraction can change a bit) makes every distance bit-equal to the reference
// and every rank / top-k index exact by construction.  The distance tile kernel
// is FP64-issue bound (64 DADD/clk/SM); the reducers stream the fp64 matrix once
// (16 B per pair against 600 DADDs per pair at d=300, i.e. noise).
#include <math_constants.h>

#include "common.cuh"

namespace eg {

constexpr int kTile = 64;     // 64 x 64 distances per CTA
constexpr int kKC = 32;       // k-chunk staged in shared memory
constexpr int kPad = 66;      // row stride (doubles) of the [k][i] tiles: 528 B keeps 16-B alignment

__global__ void __launch_bounds__(256, 2)
l1_tile_kernel(const float* __restrict__ L, int64_t nL, const float* __restrict__ R, int64_t nR, int d,
               double* __restrict__ D, int64_t ldD) {
  __shared__ __align__(16) double Ls[kKC][kPad];
  __shared__ __align__(16) double Rs[kKC][kPad];
  const int tx = threadIdx.x & 15;   // columns j0 + 2*tx + {0,1} and j0 + 32 + 2*tx + {0,1}: 16-byte lane stride -> conflict-free LDS.128
  const int ty = threadIdx.x >> 4;   // row group:    i = i0 + 4*ty .. +3
  const int64_t i0 = (int64_t)blockIdx.y * kTile;
  const int64_t j0 = (int64_t)blockIdx.x * kTile;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  // staging: thread t owns element (row = t / 32 + 8*r, k = t % 32) of both tiles — coalesced along k.
  // The NEXT chunk's global loads are issued before the current chunk's DADDs so their latency is hidden.
  const int kk = threadIdx.x & 31;
  const int rbase = threadIdx.x >> 5;
  float lreg[kTile / 8], rreg[kTile / 8];
  auto fetch = [&](int k0) {
    const int kc = min(kKC, d - k0);
#pragma unroll
    for (int r = 0; r < kTile / 8; ++r) {
      const int row = rbase + 8 * r;
      lreg[r] = (kk < kc && i0 + row < nL) ? __ldg(L + (i0 + row) * d + k0 + kk) : 0.f;
      rreg[r] = (kk < kc && j0 + row < nR) ? __ldg(R + (j0 + row) * d + k0 + kk) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < d; k0 += kKC) {
    const int kc = min(kKC, d - k0);
#pragma unroll
    for (int r = 0; r < kTile / 8; ++r) {
      Ls[kk][rbase + 8 * r] = (double)lreg[r];
      Rs[kk][rbase + 8 * r] = (double)rreg[r];
    }
    __syncthreads();
    if (k0 + kKC < d) fetch(k0 + kKC);
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
      const double2 l01 = *reinterpret_cast<const double2*>(&Ls[k][4 * ty]);
      const double2 l23 = *reinterpret_cast<const double2*>(&Ls[k][4 * ty + 2]);
      const double2 r01 = *reinterpret_cast<const double2*>(&Rs[k][2 * tx]);
      const double2 r23 = *reinterpret_cast<const double2*>(&Rs[k][32 + 2 * tx]);
      const double lv[4] = {l01.x, l01.y, l23.x, l23.y};
      const double rv[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = __dadd_rn(acc[a][b], fabs(__dsub_rn(lv[a], rv[b])));
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int64_t i = i0 + 4 * ty + a;
    if (i >= nL) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t j = j0 + 32 * h + 2 * tx;
      double* dst = D + i * ldD + j;
      if (j + 1 < nR && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(acc[a][2 * h], acc[a][2 * h + 1]);
      } else {
        if (j < nR) dst[0] = acc[a][2 * h];
        if (j + 1 < nR) dst[1] = acc[a][2 * h + 1];
      }
    }
  }
}

__global__ void l1_paired_kernel(const float* __restrict__ L, const float* __restrict__ R, int64_t n, int d,
                                 double* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* l = L + i * d;
  const float* r = R + i * d;
  double acc = 0.0;
  for (int k = 0; k < d; ++k) acc = __dadd_rn(acc, fabs(__dsub_rn((double)l[k], (double)r[k])));
  out[i] = acc;
}

// Tile of 32 rows x 256 columns; thread = column.  Row counts via ballot, column counts local.
__global__ void __launch_bounds__(256)
rank_kernel(const double* __restrict__ D, int64_t ldD, int64_t row0, int64_t n_rows, int64_t n_cols,
            const double* __restrict__ diag, int32_t* __restrict__ rank_row, int32_t* __restrict__ rank_col) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t ib = (int64_t)blockIdx.y * 32;
  const int lane = threadIdx.x & 31;
  const bool col_ok = j < n_cols;
  const double dj = col_ok ? diag[j] : 0.0;
  int col_cnt = 0;
  const int rows = (int)min((int64_t)32, n_rows - ib);
  for (int r = 0; r < rows; ++r) {
    const int64_t gi = row0 + ib + r;           // global row id == index of its true match
    const double di = diag[gi];
    bool row_hit = false;
    if (col_ok) {
      const double v = D[(ib + r) * ldD + j];
      row_hit = (v < di) || (v == di && j < gi);
      col_cnt += (v < dj) || (v == dj && gi < j);
    }
    unsigned m = __ballot_sync(0xffffffffu, row_hit);
    if (lane == 0 && m) atomicAdd(&rank_row[gi], __popc(m));
  }
  if (col_ok && col_cnt) atomicAdd(&rank_col[j], col_cnt);
}

// One warp per row: (min, lowest arg).
__global__ void row_argmin_kernel(const double* __restrict__ D, int64_t ldD, int64_t n_rows, int64_t n_cols,
                                  double* __restrict__ row_min, int64_t* __restrict__ row_arg) {
  int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const double* row = D + w * ldD;
  double best = CUDART_INF;
  int64_t arg = 0x7fffffffffffffffll;
  for (int64_t j = lane; j < n_cols; j += 32) {
    double v = row[j];
    if (v < best) { best = v; arg = j; }   // ascending j per lane: strict < keeps the lowest index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, o);
    int64_t oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) { row_min[w] = best; row_arg[w] = (n_cols > 0) ? arg : -1; }
}

// Thread per column over the rows of this block; merged into the running column result.
__global__ void col_argmin_merge_kernel(const double* __restrict__ D, int64_t ldD, int64_t row0, int64_t n_rows,
                                        int64_t n_cols, double* __restrict__ col_min,
                                        int64_t* __restrict__ col_arg) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n_cols) return;
  double best = col_min[j];
  int64_t arg = col_arg[j];
  for (int64_t i = 0; i < n_rows; ++i) {
    double v = D[i * ldD + j];
    if (v < best || (v == best && arg < 0)) { best = v; arg = row0 + i; }
  }
  col_min[j] = best;
  col_arg[j] = arg;
}

// ---- per-row top-k by (value, index) ------------------------------------------
__device__ __forceinline__ unsigned long long order_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

constexpr int kTopkThreads = 256;
constexpr int kTopkMax = 2048;

__global__ void __launch_bounds__(kTopkThreads)
topk_rows_kernel(const double* __restrict__ D, int64_t ldD, int64_t n_cols, int skip, int k,
                 int64_t* __restrict__ out_idx) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need;
  __shared__ unsigned long long keys[kTopkMax];
  __shared__ int idxs[kTopkMax];
  __shared__ int s_less, s_eq_base;
  __shared__ int warp_cnt[kTopkThreads / 32];

  const double* row = D + (int64_t)blockIdx.x * ldD;
  const int tid = threadIdx.x;
  int want = skip + k;                       // number of smallest elements to extract
  if (want > n_cols) want = (int)n_cols;
  // ---- radix select: find the key T of the want-th smallest element --------
  if (tid == 0) { s_prefix = 0ull; s_need = want; }
  __syncthreads();
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    hist[tid] = 0;                            // kTopkThreads == 256
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    for (int64_t j = tid; j < n_cols; j += kTopkThreads) {
      unsigned long long key = order_key(row[j]);
      if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need;
      unsigned acc = 0;
      int digit = 255;
      for (int b = 0; b < 256; ++b) {
        if (acc + hist[b] >= (unsigned)need) { digit = b; break; }
        acc += hist[b];
      }
      s_need = need - (int)acc;               // how many to take inside this digit bucket
      s_prefix = prefix | ((unsigned long long)digit << shift);
    }
    __syncthreads();
  }
  const unsigned long long T = s_prefix;
  const int need_eq = s_need;                 // elements equal to T to take, lowest indices first
  const int n_less = want - need_eq;
  if (tid == 0) { s_less = 0; s_eq_base = 0; }
  for (int i = tid; i < kTopkMax; i += kTopkThreads) { keys[i] = ~0ull; idxs[i] = 0x7fffffff; }
  __syncthreads();
  // ---- collect: all keys < T (any order), first need_eq keys == T in index order
  for (int64_t base = 0; base < n_cols; base += kTopkThreads) {
    int64_t j = base + tid;
    unsigned long long key = ~0ull;
    bool is_less = false, is_eq = false;
    if (j < n_cols) {
      key = order_key(row[j]);
      is_less = key < T;
      is_eq = key == T;
    }
    if (is_less) {
      int slot = atomicAdd(&s_less, 1);
      keys[slot] = key;
      idxs[slot] = (int)j;
    }
    unsigned m = __ballot_sync(0xffffffffu, is_eq);
    int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int before = s_eq_base;
    for (int w2 = 0; w2 < wid; ++w2) before += warp_cnt[w2];
    if (is_eq) {
      int pos = before + __popc(m & ((1u << lane) - 1u));
      if (pos < need_eq) { keys[n_less + pos] = key; idxs[n_less + pos] = (int)j; }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w2 = 0; w2 < kTopkThreads / 32; ++w2) tot += warp_cnt[w2];
      s_eq_base += tot;
    }
    __syncthreads();
  }
  // ---- bitonic sort of kTopkMax (key, idx) pairs --------------------------------
  int n_sort = 1;
  while (n_sort < want) n_sort <<= 1;
  if (n_sort < 2) n_sort = 2;
  for (int size = 2; size <= n_sort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < n_sort / 2; i += kTopkThreads) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        unsigned long long ka = keys[lo], kb = keys[hi];
        int ia = idxs[lo], ib = idxs[hi];
        bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
        if (a_gt_b == up) { keys[lo] = kb; keys[hi] = ka; idxs[lo] = ib; idxs[hi] = ia; }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += kTopkThreads) {
    int src = skip + i;
    out_idx[(int64_t)blockIdx.x * k + i] = (src < want) ? (int64_t)idxs[src] : -1;
  }
}

}  // namespace eg

extern "C" {

int eg_l1_matrix(const float* L, int64_t nL, const float* R, int64_t nR, int d, double* D, int64_t ldD,
                 eg_stream_t stream_) {
  using namespace eg;
  if (nL < 0 || nR < 0 || d <= 0 || ldD < nR) return EG_ERR_INVALID;
  if (nL == 0 || nR == 0) return EG_OK;
  if (!L || !R || !D) return EG_ERR_INVALID;
  int64_t gy = ceil_div(nL, kTile);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;  // callers block rows (<= 4M rows per call)
  dim3 grid((unsigned)ceil_div(nR, kTile), (unsigned)gy);
  l1_tile_kernel<<<grid, 256, 0, as_stream(stream_)>>>(L, nL, R, nR, d, D, ldD);
  EG_LAUNCHED();
  return EG_OK;
}

int eg_l1_paired(const float* L, const float* R, int64_t n, int d, double* diag, eg_stream_t stream_) {
  using namespace eg;
  if (n < 0 || d <= 0) return EG_ERR_INVALID;
  if (n == 0) return EG_OK;
  if (!L || !R || !diag) return EG_ERR_INVALID;
  l1_paired_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, as_stream(stream_)>>>(L, R, n, d, diag);
  EG_LAUNCHED();
  return EG_OK;
}

int eg_rank_accumulate(const double* D, int64_t ldD, int64_t row0, int64_t n_rows, int64_t n_cols,
                       const double* diag, int32_t* rank_row, int32_t* rank_col, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || row0 < 0 || ldD < n_cols) return EG_ERR_INVALID;
  if (n_rows == 0 || n_cols == 0) return EG_OK;
  if (!D || !diag || !rank_row || !rank_col) return EG_ERR_INVALID;
  int64_t gy = ceil_div(n_rows, 32);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;
  cudaStream_t s = as_stream(stream_);
  EG_CUDA(cudaMemsetAsync(rank_row + row0, 0, sizeof(int32_t) * (size_t)n_rows, s));
  dim3 grid((unsigned)ceil_div(n_cols, 256), (unsigned)gy);
  rank_kernel<<<grid, 256, 0, s>>>(D, ldD, row0, n_rows, n_cols, diag, rank_row, rank_col);
  EG_LAUNCHED();
  return EG_OK;
}

int eg_argmin_accumulate(const double* D, int64_t ldD, int64_t row0, int64_t n_rows, int64_t n_cols,
                         double* row_min, int64_t* row_arg, double* col_min, int64_t* col_arg,
                         eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || row0 < 0 || ldD < n_cols) return EG_ERR_INVALID;
  if (n_rows == 0 || n_cols == 0) return EG_OK;
  if (!D) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  if (row_min && row_arg) {
    row_argmin_kernel<<<(unsigned)ceil_div(n_rows * 32, 256), 256, 0, s>>>(D, ldD, n_rows, n_cols,
                                                                            row_min + row0, row_arg + row0);
    EG_LAUNCHED();
  }
  if (col_min && col_arg) {
    col_argmin_merge_kernel<<<(unsigned)ceil_div(n_cols, 128), 128, 0, s>>>(D, ldD, row0, n_rows, n_cols,
                                                                             col_min, col_arg);
    EG_LAUNCHED();
  }
  return EG_OK;
}

int eg_topk_rows(const double* D, int64_t ldD, int64_t n_rows, int64_t n_cols, int skip, int k,
                 int64_t* out_idx, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols < 0 || skip < 0 || k <= 0 || ldD < n_cols) return EG_ERR_INVALID;
  if (skip + k > kTopkMax) return EG_ERR_UNSUPPORTED;
  if (n_cols >= (1ll << 31)) return EG_ERR_UNSUPPORTED;
  if (n_rows == 0) return EG_OK;
  if (!D || !out_idx) return EG_ERR_INVALID;
  topk_rows_kernel<<<(unsigned)n_rows, kTopkThreads, 0, as_stream(stream_)>>>(D, ldD, n_cols, skip, k, out_idx);
  EG_LAUNCHED();
  return EG_OK;
}

}  // extern "C"
