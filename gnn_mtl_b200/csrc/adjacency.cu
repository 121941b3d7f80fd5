// (A) Triples -> degree-normalised CSR on the device, and CSR transpose.
//
// Restates utils/data_utils.py:296-336 (+ :51-57 value cast) of the reference as
// integer/key work on the GPU:
//   degree[v] = 1 + #(non-self-loop triples touching v as head) + #(… as tail)   (:298-305)
//   edge set  = {(h,t),(t,h) : h != t} ∪ {(v,v) : v seen in any triple}          (:308-320)
//   value     = (1 / sqrt(deg_i)) / sqrt(deg_j) in IEEE fp64, rounded once to fp32 (:334, :55)
// Sorting/unique of the 64-bit (row<<32|col) keys uses CUB (one-time preprocessing,
// not a hot op); everything numeric is explicit so the result is bit-exact.
#include <cub/cub.cuh>

#include "common.cuh"

namespace eg {

constexpr unsigned long long kInvalidKey = ~0ull;

__global__ void degree_count_kernel(const int64_t* __restrict__ heads, const int64_t* __restrict__ tails,
                                    int64_t n_triples, int* __restrict__ count, int* __restrict__ seen) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_triples) return;
  int h = (int)heads[i], t = (int)tails[i];
  seen[h] = 1;
  seen[t] = 1;
  if (h != t) {
    atomicAdd(&count[h], 1);
    atomicAdd(&count[t], 1);
  }
}

__global__ void make_keys_kernel(const int64_t* __restrict__ heads, const int64_t* __restrict__ tails,
                                 int64_t n_triples, int64_t n_ent, const int* __restrict__ seen,
                                 unsigned long long* __restrict__ keys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_triples) {
    unsigned long long h = (unsigned long long)heads[i], t = (unsigned long long)tails[i];
    bool self = (h == t);
    keys[2 * i] = self ? kInvalidKey : ((h << 32) | t);
    keys[2 * i + 1] = self ? kInvalidKey : ((t << 32) | h);
  } else if (i < n_triples + n_ent) {
    unsigned long long v = (unsigned long long)(i - n_triples);
    keys[2 * n_triples + (int64_t)v] = seen[v] ? ((v << 32) | v) : kInvalidKey;
  }
}

// Unique keys -> col (int64 + int32) and the normalised value.
__global__ void emit_entries_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ n_unique,
                                    const int* __restrict__ count, const int* __restrict__ seen,
                                    int64_t* __restrict__ col64, int32_t* __restrict__ col32,
                                    float* __restrict__ val, int64_t* __restrict__ nnz_out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t n = *n_unique;
  if (n > 0 && keys[n - 1] == kInvalidKey) --n;  // the invalid key, if present, sorts last
  if (i == 0) *nnz_out = n;
  if (i >= n) return;
  unsigned long long k = keys[i];
  int r = (int)(k >> 32), c = (int)(k & 0xffffffffull);
  double dr = (double)(seen[r] + count[r]);
  double dc = (double)(seen[c] + count[c]);
  // (1 / sqrt(dr)) / sqrt(dc), each op correctly rounded (matches CPython float arithmetic)
  double v = __ddiv_rn(__ddiv_rn(1.0, __dsqrt_rn(dr)), __dsqrt_rn(dc));
  col64[i] = c;
  col32[i] = c;
  val[i] = __double2float_rn(v);
}

// rowptr[r] = first index whose key >= (r << 32); keys sorted ascending.
__global__ void rowptr_from_keys_kernel(const unsigned long long* __restrict__ keys,
                                        const int64_t* __restrict__ nnz_ptr, int64_t n_rows,
                                        int64_t* __restrict__ crow64, int32_t* __restrict__ rowptr32) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  int64_t nnz = *nnz_ptr;
  unsigned long long target = ((unsigned long long)r) << 32;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  if (crow64) crow64[r] = lo;
  if (rowptr32) rowptr32[r] = (int32_t)lo;
}

struct AdjWorkspace {
  int* count;
  int* seen;
  unsigned long long* keys_a;
  unsigned long long* keys_b;
  int* n_unique;
  int64_t* nnz_dev;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static AdjWorkspace carve_adj(void* ws, int64_t n_triples, int64_t n_ent) {
  AdjWorkspace w{};
  int64_t n_keys = 2 * n_triples + n_ent;
  size_t sort_bytes = 0, uniq_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (unsigned long long*)nullptr,
                                 (unsigned long long*)nullptr, n_keys);
  cub::DeviceSelect::Unique(nullptr, uniq_bytes, (unsigned long long*)nullptr,
                            (unsigned long long*)nullptr, (int*)nullptr, n_keys);
  w.cub_bytes = sort_bytes > uniq_bytes ? sort_bytes : uniq_bytes;
  char* p = reinterpret_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes); return q; };
  w.count = (int*)take(sizeof(int) * (size_t)n_ent);
  w.seen = (int*)take(sizeof(int) * (size_t)n_ent);
  w.keys_a = (unsigned long long*)take(8 * (size_t)n_keys);
  w.keys_b = (unsigned long long*)take(8 * (size_t)n_keys);
  w.n_unique = (int*)take(sizeof(int));
  w.nnz_dev = (int64_t*)take(sizeof(int64_t));
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

// ---- transpose -------------------------------------------------------------

__global__ void transpose_keys_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                      int64_t n_rows, unsigned long long* __restrict__ keys,
                                      int32_t* __restrict__ pos) {
  // one warp per row
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  int b = rowptr[warp], e = rowptr[warp + 1];
  for (int i = b + lane; i < e; i += 32) {
    keys[i] = (((unsigned long long)(unsigned)col[i]) << 32) | (unsigned long long)warp;
    pos[i] = i;
  }
}

__global__ void transpose_emit_kernel(const unsigned long long* __restrict__ keys,
                                      const int32_t* __restrict__ pos, const float* __restrict__ val,
                                      int64_t nnz, int32_t* __restrict__ col_t, float* __restrict__ val_t,
                                      int32_t* __restrict__ perm_t) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  col_t[i] = (int32_t)(keys[i] & 0xffffffffull);
  val_t[i] = val[pos[i]];
  if (perm_t) perm_t[i] = pos[i];
}

struct TrWorkspace {
  unsigned long long *keys_a, *keys_b;
  int32_t *pos_a, *pos_b;
  int64_t* nnz_dev;
  void* cub_temp;
  size_t cub_bytes, total;
};

static TrWorkspace carve_tr(void* ws, int64_t nnz) {
  TrWorkspace w{};
  cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, (unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, nnz);
  char* p = reinterpret_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = p ? p + off : nullptr; off += align_up(bytes); return q; };
  size_t n = (size_t)(nnz > 0 ? nnz : 1);
  w.keys_a = (unsigned long long*)take(8 * n);
  w.keys_b = (unsigned long long*)take(8 * n);
  w.pos_a = (int32_t*)take(4 * n);
  w.pos_b = (int32_t*)take(4 * n);
  w.nnz_dev = (int64_t*)take(8);
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace eg

extern "C" {

size_t eg_adj_workspace_bytes(int64_t n_triples, int64_t n_ent) {
  if (n_triples < 0 || n_ent <= 0) return 0;
  return eg::carve_adj(nullptr, n_triples, n_ent).total;
}

int eg_adj_build(const int64_t* heads, const int64_t* tails, int64_t n_triples, int64_t n_ent, void* ws,
                 size_t ws_bytes, int64_t* crow, int64_t* col, float* val, int32_t* rowptr32,
                 int32_t* col32, int64_t* h_nnz, eg_stream_t stream_) {
  using namespace eg;
  if (n_triples < 0 || n_ent <= 0 || n_ent >= (1ll << 31) || !ws || !crow || !col || !val || !rowptr32 ||
      !col32 || !h_nnz || (n_triples > 0 && (!heads || !tails)))
    return EG_ERR_INVALID;
  AdjWorkspace w = carve_adj(ws, n_triples, n_ent);
  if (ws_bytes < w.total) return EG_ERR_WORKSPACE;
  cudaStream_t s = as_stream(stream_);
  const int T = 256;
  int64_t n_keys = 2 * n_triples + n_ent;
  EG_CUDA(cudaMemsetAsync(w.count, 0, sizeof(int) * (size_t)n_ent, s));
  EG_CUDA(cudaMemsetAsync(w.seen, 0, sizeof(int) * (size_t)n_ent, s));
  if (n_triples > 0) {
    degree_count_kernel<<<(unsigned)ceil_div(n_triples, T), T, 0, s>>>(heads, tails, n_triples, w.count, w.seen);
    EG_LAUNCHED();
  }
  make_keys_kernel<<<(unsigned)ceil_div(n_triples + n_ent, T), T, 0, s>>>(heads, tails, n_triples, n_ent,
                                                                         w.seen, w.keys_a);
  EG_LAUNCHED();
  int end_bit = 64;
  size_t tmp = w.cub_bytes;
  EG_CUDA(cub::DeviceRadixSort::SortKeys(w.cub_temp, tmp, w.keys_a, w.keys_b, n_keys, 0, end_bit, s));
  g_launches.fetch_add(1);
  tmp = w.cub_bytes;
  EG_CUDA(cub::DeviceSelect::Unique(w.cub_temp, tmp, w.keys_b, w.keys_a, w.n_unique, n_keys, s));
  g_launches.fetch_add(1);
  emit_entries_kernel<<<(unsigned)ceil_div(n_keys, T), T, 0, s>>>(w.keys_a, w.n_unique, w.count, w.seen, col,
                                                                   col32, val, w.nnz_dev);
  EG_LAUNCHED();
  rowptr_from_keys_kernel<<<(unsigned)ceil_div(n_ent + 1, T), T, 0, s>>>(w.keys_a, w.nnz_dev, n_ent, crow,
                                                                         rowptr32);
  EG_LAUNCHED();
  EG_CUDA(cudaMemcpyAsync(h_nnz, w.nnz_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  EG_CUDA(cudaStreamSynchronize(s));
  return EG_OK;
}

size_t eg_csr_transpose_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols) {
  (void)n_rows; (void)n_cols;
  if (nnz < 0) return 0;
  return eg::carve_tr(nullptr, nnz).total;
}

int eg_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* rowptr, const int32_t* col,
                     const float* val, void* ws, size_t ws_bytes, int32_t* rowptr_t, int32_t* col_t,
                     float* val_t, int32_t* perm_t, eg_stream_t stream_) {
  using namespace eg;
  if (n_rows < 0 || n_cols <= 0 || nnz < 0 || !rowptr || !rowptr_t || !ws) return EG_ERR_INVALID;
  if (nnz > 0 && (!col || !val || !col_t || !val_t)) return EG_ERR_INVALID;
  TrWorkspace w = carve_tr(ws, nnz);
  if (ws_bytes < w.total) return EG_ERR_WORKSPACE;
  cudaStream_t s = as_stream(stream_);
  const int T = 256;
  int64_t nnz_host = nnz;
  EG_CUDA(cudaMemcpyAsync(w.nnz_dev, &nnz_host, 8, cudaMemcpyHostToDevice, s));
  const unsigned long long* sorted = w.keys_a;
  if (nnz > 0) {
    transpose_keys_kernel<<<(unsigned)ceil_div(n_rows * 32, T), T, 0, s>>>(rowptr, col, n_rows, w.keys_a, w.pos_a);
    EG_LAUNCHED();
    size_t tmp = w.cub_bytes;
    EG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tmp, w.keys_a, w.keys_b, w.pos_a, w.pos_b, nnz, 0, 64, s));
    g_launches.fetch_add(1);
    sorted = w.keys_b;
    transpose_emit_kernel<<<(unsigned)ceil_div(nnz, T), T, 0, s>>>(w.keys_b, w.pos_b, val, nnz, col_t, val_t, perm_t);
    EG_LAUNCHED();
  }
  rowptr_from_keys_kernel<<<(unsigned)ceil_div(n_cols + 1, T), T, 0, s>>>(sorted, w.nnz_dev, n_cols, nullptr, rowptr_t);
  EG_LAUNCHED();
  // the host copy of nnz must outlive the async H2D above
  EG_CUDA(cudaStreamSynchronize(s));
  return EG_OK;
}

}  // extern "C"
