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

#include <algorithm>

#include "common.cuh"

namespace eg {
int g_tune_l1_filter = 1;   // eg_debug_set(13, 0): exact fp64 evaluation of every pair (no fp32 candidate filter)


constexpr int kTile = 64;     // 64 x 64 distances per CTA
constexpr int kKC = 32;       // k-chunk staged in shared memory
constexpr int kPad = 66;      // row stride (doubles) of the [k][i] tiles: 528 B keeps 16-B alignment

// acc[a][b] = sum_k |L[i0 + 4*ty + a, k] - R[j0 + col(b), k]| for one 64 x 64 tile, k ascending (SciPy's order).
// Thread layout: tx = tid & 15 owns columns j0 + 2*tx + {0,1} and j0 + 32 + 2*tx + {0,1} (16-byte lane stride ->
// conflict-free LDS.128), ty = tid >> 4 owns rows i0 + 4*ty .. +3.  Ends with a __syncthreads().
__device__ __forceinline__ int l1_col_of(int tx, int b) { return 32 * (b >> 1) + 2 * tx + (b & 1); }

__device__ __forceinline__ void l1_tile_accumulate(const float* __restrict__ L, int64_t nL,
                                                   const float* __restrict__ R, int64_t nR, int d, int64_t i0,
                                                   int64_t j0, double (*Ls)[kPad], double (*Rs)[kPad],
                                                   double (&acc)[4][4]) {
  const int tx = threadIdx.x & 15;
  const int ty = threadIdx.x >> 4;
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
}

// Same tile with an 8 x 4 register block per thread and 128 threads: thread (ty = tid >> 4, tx = tid & 15) owns
// rows i0 + 8*ty .. +7 and the same four columns as above.  An LDS.128 costs four shared-memory wavefronts per
// warp whether or not lanes share addresses, so the 4 x 4 block (4 LDS.128 per 32 DADD) keeps the shared-memory
// pipe ~83 % busy at full FP64 rate (profiles/README.md); 8 x 4 needs 6 LDS.128 per 64 DADD.
constexpr int kKC8 = 16;     // k-chunk of the 8 x 4 variant (prefetch registers: 2 x 8 per thread)

__device__ __forceinline__ void l1_tile_accumulate8(const float* __restrict__ L, int64_t nL,
                                                    const float* __restrict__ R, int64_t nR, int d, int64_t i0,
                                                    int64_t j0, double (*Ls)[kPad], double (*Rs)[kPad],
                                                    double (&acc)[8][4]) {
  const int tx = threadIdx.x & 15;
  const int ty = threadIdx.x >> 4;
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  // staging: thread t owns element (row = t / 16 + 8*r, k = t % 16) of both tiles
  const int kk = threadIdx.x & 15;
  const int rbase = threadIdx.x >> 4;
  float lreg[kTile / 8], rreg[kTile / 8];
  auto fetch = [&](int k0) {
    const int kc = min(kKC8, d - k0);
#pragma unroll
    for (int r = 0; r < kTile / 8; ++r) {
      const int row = rbase + 8 * r;
      lreg[r] = (kk < kc && i0 + row < nL) ? __ldg(L + (i0 + row) * d + k0 + kk) : 0.f;
      rreg[r] = (kk < kc && j0 + row < nR) ? __ldg(R + (j0 + row) * d + k0 + kk) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < d; k0 += kKC8) {
    const int kc = min(kKC8, d - k0);
#pragma unroll
    for (int r = 0; r < kTile / 8; ++r) {
      Ls[kk][rbase + 8 * r] = (double)lreg[r];
      Rs[kk][rbase + 8 * r] = (double)rreg[r];
    }
    __syncthreads();
    if (k0 + kKC8 < d) fetch(k0 + kKC8);
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
      double lv[8], rv[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const double2 l2 = *reinterpret_cast<const double2*>(&Ls[k][8 * ty + 2 * h]);
        lv[2 * h] = l2.x; lv[2 * h + 1] = l2.y;
      }
      const double2 r01 = *reinterpret_cast<const double2*>(&Rs[k][2 * tx]);
      const double2 r23 = *reinterpret_cast<const double2*>(&Rs[k][32 + 2 * tx]);
      rv[0] = r01.x; rv[1] = r01.y; rv[2] = r23.x; rv[3] = r23.y;
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = __dadd_rn(acc[a][b], fabs(__dsub_rn(lv[a], rv[b])));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128, 3)
l1_tile_kernel(const float* __restrict__ L, int64_t nL, const float* __restrict__ R, int64_t nR, int d,
               double* __restrict__ D, int64_t ldD) {
  __shared__ __align__(16) double Ls[kKC8][kPad];
  __shared__ __align__(16) double Rs[kKC8][kPad];
  const int tx = threadIdx.x & 15;
  const int ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * kTile;
  const int64_t j0 = (int64_t)blockIdx.x * kTile;
  double acc[8][4];
  l1_tile_accumulate8(L, nL, R, nR, d, i0, j0, Ls, Rs, acc);
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    int64_t i = i0 + 8 * ty + a;
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

// ---- streamed variants: the distance matrix is never stored ---------------------------------------
// A CTA owns a 64-row strip and walks `tiles_per_cta` consecutive column tiles.

// Rank counts of eg_rank_accumulate, taken straight from the accumulators.
__global__ void __launch_bounds__(128, 3)
l1_rank_fused_kernel(const float* __restrict__ L, int64_t nL, int64_t row0, const float* __restrict__ R, int64_t nR,
                     int d, int tiles_per_cta, const double* __restrict__ diag, int32_t* __restrict__ rank_row,
                     int32_t* __restrict__ rank_col, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;       // conditional launch (fallback of the filtered path)
  __shared__ __align__(16) double Ls[kKC8][kPad];
  __shared__ __align__(16) double Rs[kKC8][kPad];
  __shared__ int col_cnt_s[2][kTile];
  const int tx = threadIdx.x & 15;
  const int ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * kTile;
  const int64_t n_col_tiles = (nR + kTile - 1) / kTile;
  const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t_end = min(n_col_tiles, t_begin + tiles_per_cta);
  double di[8];
  int64_t gi[8];
  int row_cnt[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t i = i0 + 8 * ty + a;
    gi[a] = row0 + i;                              // global row id == index of its true match
    di[a] = (i < nL) ? diag[gi[a]] : 0.0;
    row_cnt[a] = 0;
  }
  (&col_cnt_s[0][0])[threadIdx.x] = 0;             // 128 threads == 2 * kTile entries
  __syncthreads();
  int buf = 0;
  for (int64_t t = t_begin; t < t_end; ++t, buf ^= 1) {
    const int64_t j0 = t * kTile;
    double acc[8][4];
    l1_tile_accumulate8(L, nL, R, nR, d, i0, j0, Ls, Rs, acc);
    int col_cnt[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = j0 + l1_col_of(tx, b);
      const bool col_ok = j < nR;
      const double dj = col_ok ? __ldg(diag + j) : 0.0;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const bool ok = col_ok && (i0 + 8 * ty + a < nL);
        const double v = acc[a][b];
        row_cnt[a] += ok && ((v < di[a]) || (v == di[a] && j < gi[a]));
        col_cnt[b] += ok && ((v < dj) || (v == dj && gi[a] < j));
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int c = col_cnt[b] + __shfl_xor_sync(0xffffffffu, col_cnt[b], 16);   // the warp's two row groups
      if ((threadIdx.x & 16) == 0 && c) atomicAdd(&col_cnt_s[buf][l1_col_of(tx, b)], c);
    }
    __syncthreads();
    if (threadIdx.x < kTile) {
      const int c = col_cnt_s[buf][threadIdx.x];
      if (c) { atomicAdd(&rank_col[j0 + threadIdx.x], c); col_cnt_s[buf][threadIdx.x] = 0; }
    }
    // no second barrier: the next tile adds into the other buffer, and l1_tile_accumulate8's own barriers
    // order this buffer's reset before its reuse two tiles later
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    int c = row_cnt[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (tx == 0 && c && i0 + 8 * ty + a < nL) atomicAdd(&rank_row[gi[a]], c);
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

// ---- streamed per-row top-k ------------------------------------------------------------------------
// A CTA keeps, for each of its 64 rows, the `want` smallest (key, index) pairs seen so far in shared memory
// (unsorted, with the position of the current worst entry tracked).  After every tile each thread tests its 16
// distances against its rows' worst entries; survivors (rare once the lists have warmed up: ~want*ln(n/want)
// per row in total) go into a small per-row queue and are folded in by the row's quad of threads: replace the
// worst entry, rescan for the new worst (4 lanes x want/4 entries + two shuffles).  The order is the total
// order (value, index), so the result does not depend on arrival order and equals the stable argsort the
// reference takes.  Lists are rank-sorted once at the end.
constexpr int kRowQueue = 16;

struct TopkSmem {
  double (*Ls)[kPad];
  double (*Rs)[kPad];
  unsigned long long* list_key;   // [64][wp]
  unsigned long long* q_key;      // [64][kRowQueue]
  unsigned long long* worst_key;  // [64]
  int* list_idx;                  // [64][wp]
  int* q_idx;                     // [64][kRowQueue]
  int* worst_idx;                 // [64]
  int* worst_pos;                 // [64]
  int* q_cnt;                     // [64]
};

__host__ __device__ inline size_t topk_smem_bytes(int wp) {
  return 2 * sizeof(double) * kKC * kPad + (size_t)kTile * (wp + kRowQueue + 1) * 12 + (size_t)kTile * 8;
}

__device__ __forceinline__ TopkSmem carve_topk(unsigned char* base, int wp) {
  TopkSmem m;
  m.Ls = reinterpret_cast<double(*)[kPad]>(base);
  m.Rs = reinterpret_cast<double(*)[kPad]>(base + sizeof(double) * kKC * kPad);
  unsigned char* p = base + 2 * sizeof(double) * kKC * kPad;
  m.list_key = reinterpret_cast<unsigned long long*>(p);  p += (size_t)kTile * wp * 8;
  m.q_key = reinterpret_cast<unsigned long long*>(p);     p += (size_t)kTile * kRowQueue * 8;
  m.worst_key = reinterpret_cast<unsigned long long*>(p); p += (size_t)kTile * 8;
  m.list_idx = reinterpret_cast<int*>(p);                 p += (size_t)kTile * wp * 4;
  m.q_idx = reinterpret_cast<int*>(p);                    p += (size_t)kTile * kRowQueue * 4;
  m.worst_idx = reinterpret_cast<int*>(p);                p += (size_t)kTile * 4;
  m.worst_pos = reinterpret_cast<int*>(p);                p += (size_t)kTile * 4;
  m.q_cnt = reinterpret_cast<int*>(p);
  return m;
}

__device__ __forceinline__ bool pair_less(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(256)
l1_topk_stream_kernel(const float* __restrict__ L, int64_t nL, const float* __restrict__ R, int64_t nR, int d,
                      int tiles_per_cta, int want, int wp, unsigned long long* __restrict__ part_key,
                      int* __restrict__ part_idx) {
  extern __shared__ __align__(16) unsigned char topk_smem_raw[];
  const TopkSmem m = carve_topk(topk_smem_raw, wp);
  const int tid = threadIdx.x;
  const int tx = tid & 15;
  const int ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * kTile;
  const int64_t n_col_tiles = (nR + kTile - 1) / kTile;
  const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t_end = min(n_col_tiles, t_begin + tiles_per_cta);
  for (int i = tid; i < kTile * wp; i += 256) { m.list_key[i] = ~0ull; m.list_idx[i] = 0x7fffffff; }
  if (tid < kTile) {
    m.worst_key[tid] = ~0ull; m.worst_idx[tid] = 0x7fffffff; m.worst_pos[tid] = 0; m.q_cnt[tid] = 0;
  }
  __syncthreads();
  const int qrow = tid >> 2, qlane = tid & 3;          // drain: a quad of lanes per row
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t j0 = t * kTile;
    double acc[4][4];
    l1_tile_accumulate(L, nL, R, nR, d, i0, j0, m.Ls, m.Rs, acc);
    // bit a*4+b set <=> acc[a][b] still beats the row's current worst entry
    auto beats_worst = [&]() {
      unsigned pending = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int row = 4 * ty + a;
        if (i0 + row >= nL) continue;
        const unsigned long long tk = m.worst_key[row];
        const int ti = m.worst_idx[row];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int64_t j = j0 + l1_col_of(tx, b);
          if (j < nR && pair_less(order_key(acc[a][b]), (int)j, tk, ti)) pending |= 1u << (a * 4 + b);
        }
      }
      return pending;
    };
    unsigned pending = beats_worst();
    while (__syncthreads_or(pending != 0)) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (pending & (1u << (a * 4 + b))) {
            const int row = 4 * ty + a;
            const int slot = atomicAdd(&m.q_cnt[row], 1);
            if (slot < kRowQueue) {                    // queued: the drain below settles it for good
              m.q_key[row * kRowQueue + slot] = order_key(acc[a][b]);
              m.q_idx[row * kRowQueue + slot] = (int)(j0 + l1_col_of(tx, b));
              pending &= ~(1u << (a * 4 + b));
            }
          }
      __syncthreads();
      {
        const unsigned quad = 0xFu << (tid & 28);               // the four lanes of this row (tid & 31 & ~3)
        const int n = min(m.q_cnt[qrow], kRowQueue);
        unsigned long long* lk = m.list_key + qrow * wp;
        int* li = m.list_idx + qrow * wp;
        unsigned long long wk = m.worst_key[qrow];
        int wi = m.worst_idx[qrow], wpos = m.worst_pos[qrow];
        for (int e = 0; e < n; ++e) {
          const unsigned long long key = m.q_key[qrow * kRowQueue + e];
          const int idx = m.q_idx[qrow * kRowQueue + e];
          if (!pair_less(key, idx, wk, wi)) continue;            // uniform across the quad
          __syncwarp(quad);
          if (qlane == 0) { lk[wpos] = key; li[wpos] = idx; }
          __syncwarp(quad);
          // new worst: maximum of the list under (key, idx, position)
          unsigned long long bk = 0ull;
          int bi = -1, bp = -1;
          for (int p2 = qlane; p2 < want; p2 += 4) {
            const unsigned long long k2 = lk[p2];
            const int i2 = li[p2];
            if (bp < 0 || pair_less(bk, bi, k2, i2) || (bk == k2 && bi == i2)) { bk = k2; bi = i2; bp = p2; }
          }
#pragma unroll
          for (int o = 1; o < 4; o <<= 1) {
            const unsigned long long ok2 = __shfl_xor_sync(quad, bk, o);
            const int oi = __shfl_xor_sync(quad, bi, o);
            const int op = __shfl_xor_sync(quad, bp, o);
            const bool take = op >= 0 && (bp < 0 || pair_less(bk, bi, ok2, oi) || (bk == ok2 && bi == oi && op > bp));
            if (take) { bk = ok2; bi = oi; bp = op; }
          }
          wk = bk; wi = bi; wpos = bp;
        }
        __syncwarp(quad);
        if (qlane == 0) {
          m.worst_key[qrow] = wk; m.worst_idx[qrow] = wi; m.worst_pos[qrow] = wpos; m.q_cnt[qrow] = 0;
        }
      }
      __syncthreads();
      if (pending) pending &= beats_worst();       // left out for lack of room: retry unless the bar has moved past them
    }
  }
  __syncthreads();
  // rank sort of each row's list by (key, idx, position) and write-out in order
  for (int i = tid; i < kTile * want; i += 256) {
    const int row = i / want, pos = i - row * want;
    if (i0 + row >= nL) continue;
    const unsigned long long* lk = m.list_key + row * wp;
    const int* li = m.list_idx + row * wp;
    const unsigned long long key = lk[pos];
    const int idx = li[pos];
    int rank = 0;
    for (int p2 = 0; p2 < want; ++p2) {
      const unsigned long long k2 = lk[p2];
      const int i2 = li[p2];
      rank += pair_less(k2, i2, key, idx) || (k2 == key && i2 == idx && p2 < pos);
    }
    const int64_t o = ((int64_t)blockIdx.x * nL + i0 + row) * want + rank;
    part_key[o] = key;
    part_idx[o] = idx;
  }
}

// Merge the per-segment sorted lists of one row (thread = row) and emit entries [skip, skip + k).
__global__ void topk_merge_kernel(const unsigned long long* __restrict__ part_key, const int* __restrict__ part_idx,
                                  int64_t nL, int n_seg, int want, int skip, int k, int64_t n_cols,
                                  int64_t* __restrict__ out_idx) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= nL) return;
  int head[64];
  for (int s = 0; s < n_seg; ++s) head[s] = 0;
  const int avail = (int)min((int64_t)want, n_cols);
  for (int pos = 0; pos < skip + k; ++pos) {
    int64_t pick = -1;
    if (pos < avail) {
      int best = -1;
      unsigned long long bk = ~0ull;
      int bi = 0x7fffffff;
      for (int s = 0; s < n_seg; ++s) {
        if (head[s] >= want) continue;
        const int64_t o = ((int64_t)s * nL + row) * want + head[s];
        const unsigned long long key = part_key[o];
        const int idx = part_idx[o];
        if (idx != 0x7fffffff && (best < 0 || pair_less(key, idx, bk, bi))) { best = s; bk = key; bi = idx; }
      }
      if (best >= 0) { ++head[best]; pick = bi; }
    }
    if (pos >= skip) out_idx[row * k + (pos - skip)] = pick;
  }
}

static int topk_segments(int64_t nL, int64_t nR, int* tiles_per_cta) {
  const int64_t row_tiles = ceil_div(nL, kTile), col_tiles = ceil_div(nR, kTile);
  int64_t n_seg = ceil_div((int64_t)(4 * 148), row_tiles);          // aim for >= 4 CTAs per SM in flight overall
  n_seg = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_seg, col_tiles), 64));
  const int64_t per = ceil_div(col_tiles, n_seg);
  *tiles_per_cta = (int)per;
  return (int)ceil_div(col_tiles, per);
}

// rank_row[row0 .. row0 + nL) must be zero on entry when run_if is given (the filtered path leaves it untouched
// on overflow and the caller zeroes it up front); the unconditional call zeroes it itself.
int l1_rank_exact_launch(const float* L, int64_t nL, int64_t row0, const float* R, int64_t nR, int d,
                         const double* diag, int32_t* rank_row, int32_t* rank_col, const int* run_if, cudaStream_t s) {
  const int64_t gy = ceil_div(nL, kTile), col_tiles = ceil_div(nR, kTile);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;
  // column segments per strip: enough CTAs for ~8 per SM overall, at most one per column tile
  int64_t n_seg = std::max<int64_t>(1, std::min<int64_t>(col_tiles, ceil_div((int64_t)(8 * 148), gy)));
  const int per = (int)ceil_div(col_tiles, n_seg);
  n_seg = ceil_div(col_tiles, (int64_t)per);
  if (run_if == nullptr) EG_CUDA(cudaMemsetAsync(rank_row + row0, 0, sizeof(int32_t) * (size_t)nL, s));
  dim3 grid((unsigned)n_seg, (unsigned)gy);
  l1_rank_fused_kernel<<<grid, 128, 0, s>>>(L, nL, row0, R, nR, d, per, diag, rank_row, rank_col, run_if);
  EG_LAUNCHED();
  return EG_OK;
}

}  // namespace eg

extern "C" {

int eg_l1_rank_fused(const float* L, int64_t nL, int64_t row0, const float* R, int64_t nR, int d,
                     const double* diag, int32_t* rank_row, int32_t* rank_col, eg_stream_t stream_) {
  using namespace eg;
  if (nL < 0 || nR < 0 || row0 < 0 || d <= 0 || row0 + nL > nR) return EG_ERR_INVALID;
  if (nL == 0 || nR == 0) return EG_OK;
  if (!L || !R || !diag || !rank_row || !rank_col) return EG_ERR_INVALID;
  return l1_rank_exact_launch(L, nL, row0, R, nR, d, diag, rank_row, rank_col, nullptr, as_stream(stream_));
}

size_t eg_l1_topk_fused_workspace_bytes(int64_t nL, int64_t nR, int skip, int k) {
  using namespace eg;
  if (nL <= 0 || nR <= 0 || skip < 0 || k <= 0) return 0;
  int per = 0;
  const int n_seg = topk_segments(nL, nR, &per);
  return align_up((size_t)n_seg * (size_t)nL * (size_t)(skip + k) * 8) +
         align_up((size_t)n_seg * (size_t)nL * (size_t)(skip + k) * 4);
}

int eg_l1_topk_fused(const float* L, int64_t nL, const float* R, int64_t nR, int d, int skip, int k, void* ws,
                     size_t ws_bytes, int64_t* out_idx, eg_stream_t stream_) {
  using namespace eg;
  if (nL < 0 || nR < 0 || d <= 0 || skip < 0 || k <= 0) return EG_ERR_INVALID;
  const int want = skip + k;
  if (want > 128 || nR >= (1ll << 31)) return EG_ERR_UNSUPPORTED;
  if (nL == 0) return EG_OK;
  if (!L || !R || !out_idx || !ws) return EG_ERR_INVALID;
  if (ws_bytes < eg_l1_topk_fused_workspace_bytes(nL, nR, skip, k)) return EG_ERR_WORKSPACE;
  const int64_t gy = ceil_div(nL, kTile);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;
  int per = 0;
  const int n_seg = topk_segments(nL, nR, &per);
  unsigned long long* part_key = reinterpret_cast<unsigned long long*>(ws);
  int* part_idx = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) +
                                         align_up((size_t)n_seg * (size_t)nL * (size_t)want * 8));
  const int wp = want | 1;                                   // odd row stride: conflict-free list columns
  const size_t smem = topk_smem_bytes(wp);
  cudaStream_t s = as_stream(stream_);
  EG_CUDA(cudaFuncSetAttribute(l1_topk_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)n_seg, (unsigned)gy);
  l1_topk_stream_kernel<<<grid, 256, smem, s>>>(L, nL, R, nR, d, per, want, wp, part_key, part_idx);
  EG_LAUNCHED();
  topk_merge_kernel<<<(unsigned)ceil_div(nL, 128), 128, 0, s>>>(part_key, part_idx, nL, n_seg, want, skip, k, nR,
                                                                out_idx);
  EG_LAUNCHED();
  return EG_OK;
}


int eg_l1_matrix(const float* L, int64_t nL, const float* R, int64_t nR, int d, double* D, int64_t ldD,
                 eg_stream_t stream_) {
  using namespace eg;
  if (nL < 0 || nR < 0 || d <= 0 || ldD < nR) return EG_ERR_INVALID;
  if (nL == 0 || nR == 0) return EG_OK;
  if (!L || !R || !D) return EG_ERR_INVALID;
  int64_t gy = ceil_div(nL, kTile);
  if (gy > 65535) return EG_ERR_UNSUPPORTED;  // callers block rows (<= 4M rows per call)
  dim3 grid((unsigned)ceil_div(nR, kTile), (unsigned)gy);
  l1_tile_kernel<<<grid, 128, 0, as_stream(stream_)>>>(L, nL, R, nR, d, D, ldD);
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
