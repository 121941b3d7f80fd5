/*
 * eagraft — C ABI of the B200 (sm_100a) entity-alignment hot path.
 *
 * One shared library (gnn_mtl_b200/csrc/libeagraft.so), plain C linkage, raw
 * device pointers + sizes + a cudaStream_t passed as void*.  No PyTorch types.
 * The reference (HestiaSky/GNN-MTL) has no FFI of its own — its operator API is
 * Python classes/functions — so each entry point below names the reference
 * call site (file:line, relative to the reference checkout) whose work it takes
 * over.  The Python host code in gnn_mtl_b200/ keeps the reference's names and
 * signatures and calls these through ctypes.
 *
 * Conventions
 *   - return 0 (EG_OK) or a negative eg_status; never throws, never exits;
 *   - all pointers are DEVICE pointers unless the name starts with h_;
 *   - work is enqueued on `stream` and is asynchronous unless stated;
 *   - no hidden allocation: scratch comes from the caller (`ws`, `ws_bytes`),
 *     sized by the matching *_workspace_bytes();
 *   - pointers are borrowed for the duration of the enqueued work only;
 *   - row-major, contiguous unless an ld* argument says otherwise.
 */
#ifndef EAGRAFT_H_
#define EAGRAFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* eg_stream_t; /* cudaStream_t */

enum eg_status {
  EG_OK = 0,
  EG_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unsupported d) */
  EG_ERR_CUDA = -2,        /* a CUDA runtime call failed; see eg_last_cuda_error() */
  EG_ERR_WORKSPACE = -3,   /* ws_bytes smaller than *_workspace_bytes() */
  EG_ERR_UNSUPPORTED = -4, /* valid request this build cannot serve (e.g. k too large) */
  EG_ERR_NO_DEVICE = -5    /* no sm_100 device visible */
};

enum eg_act { EG_ACT_IDENTITY = 0, EG_ACT_RELU = 1 };
enum eg_cost { EG_COST_L2 = 0, EG_COST_SQEUCLID = 1, EG_COST_COSINE = 2 };
enum eg_algo { EG_ALGO_SIMT = 0, EG_ALGO_TCGEN05 = 1 };

/* ---- library ---------------------------------------------------------------- */
int eg_version(void);
const char* eg_strerror(int status);
int eg_last_cuda_error(void);          /* cudaError_t of the last failure on this thread */
int eg_device_check(void);             /* EG_OK iff current device is compute capability 10.x */
int64_t eg_launch_count(void);         /* kernels launched by this library since reset */
void eg_launch_count_reset(void);

/* Diagnostics / kernel-variant selection for tests and tools.  NOT part of the data path: production callers
 * never call it and every knob defaults to the shipping kernel.  The knobs are process-global plain ints, so —
 * unlike every other entry point, which may be called concurrently from several host threads on different
 * streams — this call must not race with calls that are inside the library.
 *   set (returns EG_OK):  0/1 SpMM rows-in-flight (0 = chosen by row width) / warps per CTA, 2 SpMM L2 hints, 3 persistent Sinkhorn on/off,
 *     4 resident rows on/off, 5 on-chip fp32 Sinkhorn on/off, 6 persistent SpMM CTAs per SM, 7 scaling-domain
 *     continuation on/off, 10 fold threshold (|log2| x 1000), 11 force the log-domain redo, 12 2-D tiled scaling
 *     kernel on/off (off: row-block kernel), 13 fp32 candidate filter of the L1 rank kernels on/off, 14 SpMM feature
 *     slab width in float4 (0 / 32 / 64), 16 SpMM neighbour rows through cp.async.bulk + shared memory on/off,
 *     17 persistent SpMM (knob 6) hands out rows through an atomic counter on/off,
 *     18 eg_gemm_nt_3xtf32_raw on CTA pairs (tcgen05.mma.cta_group::2, 256-row tiles; default on) / single-SM kernel,
 *     19 the fused Sinkhorn / plan / split-operand NT GEMM kernels (tcgen05) on CTA pairs (default on) / single CTAs
 *        (EG_TC_PAIR=0|1 in the environment overrides knob 19 for whole-process measurements)
 *   query (value ignored): 8 scaling-domain solves redone in the log domain so far, 9 fold steps so far
 * Queries 8/9 read device counters and synchronise the device. */
int eg_debug_set(int key, int value);

/* ---- (A) adjacency: triples -> degree-normalised CSR -------------------------
 * Replaces utils/data_utils.py:296-336 (get_matrix, get_sparse_tensor) and the
 * fp64->fp32 cast of sparse_mx_to_torch_sparse_tensor (:51-57).  Output equals
 * that tensor's .coalesce().to_sparse_csr() bit for bit.
 * capacity = 2*n_triples + n_ent entries for col/val outputs.
 * SYNCHRONOUS: waits for the stream so *h_nnz is valid on return.
 */
size_t eg_adj_workspace_bytes(int64_t n_triples, int64_t n_ent);
int eg_adj_build(const int64_t* heads, const int64_t* tails, int64_t n_triples, int64_t n_ent,
                 void* ws, size_t ws_bytes,
                 int64_t* crow /* [n_ent+1] */, int64_t* col /* [capacity] */, float* val /* [capacity] */,
                 int32_t* rowptr32 /* [n_ent+1] */, int32_t* col32 /* [capacity] */,
                 int64_t* h_nnz /* host */, eg_stream_t stream);

/* CSR -> CSR of the transpose (for dH = A^T dS; layers/layers.py:35,64 autograd).
 * Entries of each output row are ordered by ascending source row.  perm_t (nullable, [nnz]) receives,
 * for every transposed entry, its position in the source CSR (to carry per-edge values across). */
size_t eg_csr_transpose_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols);
int eg_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz,
                     const int32_t* rowptr, const int32_t* col, const float* val,
                     void* ws, size_t ws_bytes,
                     int32_t* rowptr_t /* [n_cols+1] */, int32_t* col_t /* [nnz] */, float* val_t /* [nnz] */,
                     int32_t* perm_t /* [nnz], nullable */, eg_stream_t stream);

/* ---- (a) message-passing SpMM with fused epilogue ----------------------------
 * Replaces torch.spmm(adj, hidden) + act + highway blend, layers/layers.py:35-38
 * and :64-76:   S = A·H;  a = act(S);
 *   gate_pre == NULL : out = a
 *   gate_pre != NULL : t = sigmoid(gate_pre); out = t*a + (1-t)*x_res
 * act_out (nullable) receives a = act(S) for the backward pass.
 * Rows longer than `long_row_threshold` are split into segments listed in
 * seg_* (built by the host mirror once per adjacency; n_seg may be 0); their
 * partial sums go through `seg_scratch` [n_seg, d] and are finished by a second
 * small launch over long_rows[n_long] / long_first[n_long+1].
 * H, out, gate_pre, x_res, act_out: [n_rows or n_cols, d] fp32, 16-byte aligned
 * when d % 4 == 0.
 */
int eg_spmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_rows,
            const float* H, int d, int act,
            const float* gate_pre, const float* x_res,
            float* out, float* act_out,
            int long_row_threshold,
            const int32_t* seg_row, const int32_t* seg_begin, const int32_t* seg_end, int64_t n_seg,
            const int32_t* long_rows, const int32_t* long_first, int64_t n_long,
            float* seg_scratch, eg_stream_t stream);

/* Element-wise backward of the fused epilogue (autograd of layers/layers.py:67-76):
 *   dS      = dout * t * act'(S)           (t = 1 when gate_pre == NULL)
 *   d_gate  = dout * (a - x_res) * t*(1-t) (nullable)
 *   d_xres  = dout * (1 - t)               (nullable)
 * `a` is act_out saved by eg_spmm (relu mask = a > 0). */
int eg_epilogue_bwd(const float* dout, const float* a, const float* gate_pre, const float* x_res,
                    int64_t n_elem, int act, float* dS, float* d_gate, float* d_xres, eg_stream_t stream);

/* ---- (c) alignment evaluation: exact fp64 L1 distances + ranks/top-k/argmin ---
 * Replaces scipy.spatial.distance.cdist(..., 'cityblock') at
 * utils/eval_utils.py:74 and models/models_ea.py:24,149: fp32 inputs widened to
 * fp64, |l-r| summed over k = 0..d-1 in that order, so the result is bit-equal
 * to SciPy's.  D is [nL, ldD] fp64 (ldD >= nR). */
int eg_l1_matrix(const float* L, int64_t nL, const float* R, int64_t nR, int d,
                 double* D, int64_t ldD, eg_stream_t stream);
/* diag[i] = sum_k |L[i,k] - R[i,k]|, same summation order as eg_l1_matrix. */
int eg_l1_paired(const float* L, const float* R, int64_t n, int d, double* diag, eg_stream_t stream);
/* Rank of the true match (utils/eval_utils.py:77-89, stable tie order): for the
 * row block [row0, row0+n_rows) of a score matrix D (fp64, [n_rows, ldD], n_cols
 * columns): rank_row[row0+i] = #{j : D[i,j] < diag[row0+i]} + #{j < row0+i : ==},
 * and rank_col[j] += #{i : D[i,j] < diag[j]} + #{row0+i < j : ==}.
 * rank_col must be zeroed by the caller before the first block. */
int eg_rank_accumulate(const double* D, int64_t ldD, int64_t row0, int64_t n_rows, int64_t n_cols,
                       const double* diag, int32_t* rank_row, int32_t* rank_col, eg_stream_t stream);
/* Row / column arg-min with lowest-index ties (utils/eval_utils.py:167,
 * models/models_ea.py:151-153).  Column results are merged into col_min/col_arg
 * across successive row blocks (caller initialises col_min=+inf, col_arg=-1). */
int eg_argmin_accumulate(const double* D, int64_t ldD, int64_t row0, int64_t n_rows, int64_t n_cols,
                         double* row_min, int64_t* row_arg, double* col_min, int64_t* col_arg,
                         eg_stream_t stream);
/* Per row the indices ranked [skip, skip+k) by (value, index) ascending
 * (models/models_ea.py:26-27: argsort()[1:k+1]).  skip+k <= 2048. */
int eg_topk_rows(const double* D, int64_t ldD, int64_t n_rows, int64_t n_cols,
                 int skip, int k, int64_t* out_idx /* [n_rows, k] */, eg_stream_t stream);

/* Streamed variants of the three calls above: the fp64 distance matrix is never stored (16 B per pair of
 * HBM traffic avoided; what BASELINE config 5's 1M x 1M evaluation needs).  Bit-identical results.
 * eg_l1_rank_fused == eg_l1_matrix(L, R) + eg_rank_accumulate(row0, ...): rank_row[row0 .. row0+nL) is
 * overwritten, rank_col accumulated.
 * eg_l1_topk_fused == eg_l1_matrix + eg_topk_rows for skip + k <= 128. */
int eg_l1_rank_fused(const float* L, int64_t nL, int64_t row0, const float* R, int64_t nR, int d,
                     const double* diag, int32_t* rank_row, int32_t* rank_col, eg_stream_t stream);
/* eg_l1_rank_fused with the distance tiles on the FP32 pipe as a CANDIDATE FILTER (utils/eval_utils.py:74-89):
 * fl32 sums decide every comparison whose outcome is certain under the rounding bound (d + 2) * 2^-24; the pairs
 * inside the band are re-evaluated exactly in fp64 (SciPy's summation order) by a second kernel, so ranks are
 * identical to eg_l1_rank_fused.  If the candidate queue overflows (tie-heavy data) the exact kernel runs
 * instead — decided on the device; the call never synchronises.  nR < 2^31. */
size_t eg_l1_rank_filtered_workspace_bytes(int64_t nL, int64_t nR);
int eg_l1_rank_filtered(const float* L, int64_t nL, int64_t row0, const float* R, int64_t nR, int d,
                        const double* diag, int32_t* rank_row, int32_t* rank_col, void* ws, size_t ws_bytes,
                        eg_stream_t stream);
size_t eg_l1_topk_fused_workspace_bytes(int64_t nL, int64_t nR, int skip, int k);
int eg_l1_topk_fused(const float* L, int64_t nL, const float* R, int64_t nR, int d, int skip, int k,
                     void* ws, size_t ws_bytes, int64_t* out_idx, eg_stream_t stream);

/* ---- (b) Sinkhorn, log domain --------------------------------------------------
 * One half-sweep ("pass") on a MATERIALISED cost (utils/ot_loss.py:53-55 and
 * SinkhornOT/sinkhorn_loss.py:197-201 in log form):
 *     lse[i]     = log sum_j exp(pot_in[j] - M[i,j] * inv_reg)
 *     pot_out[i] = logw[i] - lse[i]
 * for every row i of M [n_rows, ld].  The column half-sweep is the same call on
 * the transposed cost (eg_transpose).  dtype: 0 = fp32, 1 = fp64 for M, pot_*,
 * logw, lse alike.  lse_out nullable. */
int eg_lse_dense(int dtype, const void* M, int64_t n_rows, int64_t n_cols, int64_t ld, double inv_reg,
                 const void* pot_in, const void* logw, void* pot_out, void* lse_out, eg_stream_t stream);
int eg_transpose(int dtype, const void* src, int64_t n_rows, int64_t n_cols, int64_t ld_src,
                 void* dst, int64_t ld_dst, eg_stream_t stream);
/* Plan and primal cost (utils/ot_loss.py:75-76): P[i,j] = exp(f[i] + g[j] - M[i,j]*inv_reg)
 * (P nullable), *loss = sum P∘M (fp64), row_sum / col_sum (nullable, dtype-typed)
 * receive the marginals.  loss/row_sum/col_sum are zeroed by the call itself. */
int eg_plan_dense(int dtype, const void* M, int64_t n_rows, int64_t n_cols, int64_t ld, double inv_reg,
                  const void* f, const void* g, void* P, int64_t ldP, double* loss,
                  void* row_sum, void* col_sum, eg_stream_t stream);
/* Whole solver of utils/ot_loss.py:26-76 on the device (materialised cost):
 * u0 = 1/I, v0 = 1/J, column then row update per sweep, marginal-error check on
 * sweeps 0,10,20,…, stop at err <= stop_thr or max_iter.  log_u/log_v receive the final log-scalings.
 * Mt is scratch for the transposed cost [n_cols, n_rows]; only the streaming path (shapes the one-launch
 * kernels cannot take) needs it: pass NULL to skip the allocation — the call then returns
 * EG_ERR_WORKSPACE, before doing any work, when that path is required.
 * h_sweeps / h_err are host outputs.  Synchronisation: with stop_thr >= 0 the call waits for the solve (the
 * sweep count and error are results); with stop_thr < 0 ("run every sweep") on the one-launch fp32 path
 * nothing is read back — the call is ASYNCHRONOUS, *h_sweeps = max_iter and *h_err = -1. */
int eg_sinkhorn_dense(int dtype, const void* M, int64_t n_rows, int64_t n_cols, double reg,
                      const void* a, const void* b, int max_iter, double stop_thr,
                      void* Mt, void* log_u, void* log_v, void* ws, size_t ws_bytes,
                      int* h_sweeps, double* h_err, eg_stream_t stream);
size_t eg_sinkhorn_dense_workspace_bytes(int dtype, int64_t n_rows, int64_t n_cols);
/* Measurement aid: `iters` repetitions of the per-sweep EXCHANGE of the one-launch fp32 solver with no mat-vec
 * work (column partials through global memory + one grid barrier + partial reads, row partials through
 * distributed shared memory + one cluster barrier), same launch shape as the solve of an n_rows x n_cols problem;
 * ws as for eg_sinkhorn_dense(dtype 0).  Timed by the caller: the latency floor bench.py reports next to the
 * measured sweep time.  EG_ERR_UNSUPPORTED when the tile kernel does not take this shape. */
int eg_sinkhorn_sync_floor(int64_t n_rows, int64_t n_cols, int iters, void* ws, size_t ws_bytes, eg_stream_t stream);
/* Measurement aid: sustained issue rate of one instruction class on the current device, in lane-operations per second
 * (kind 0 fp32 FMA, 1 MUFU ex2.approx, 2 fp64 add, 3 fp32 add with |.| — the pattern of the L1 candidate filter).
 * These are the measured denominators of the SIMT-bound rooflines (SURVEY.md 8d quotes nominal ones; BASELINE.md
 * section 3 asks for measured).  SYNCHRONOUS (times its own launches with events). */
int eg_issue_peak(int kind, int iters, double* h_lane_ops_per_s, void* scratch /* device, >= 4 bytes */, eg_stream_t stream);

/* FUSED half-sweep: the cost tile is recomputed from the two embedding sets and
 * reduced straight into the row log-sum-exp; M is never written
 * (models/models_ea.py:218 + utils/ot_loss.py:53-55 fused).
 *     lse[i] = log sum_j exp(pot_in[j] - cost(A_i, B_j) * inv_reg),  pot_out = logw - lse
 * A [nA, d], B [nB, d] fp32.  Call with (A,B) = (X,Y) for the row update and
 * (Y,X) for the column update.  normA/normB: squared row norms (L2, SQEUCLID) or
 * row norms (COSINE), from eg_row_norms.  algo: EG_ALGO_SIMT (fp32 FMA tiles) or
 * EG_ALGO_TCGEN05 (TMA-fed tcgen05 3xTF32 tiles; needs the split operands from
 * eg_split_tf32 passed as A_hi/A_lo/B_hi/B_lo, else NULL). */
int eg_row_norms(const float* A, int64_t n, int d, int squared, float* out, eg_stream_t stream);
int eg_lse_fused(int algo, int cost, const float* A, int64_t nA, const float* B, int64_t nB, int d,
                 const float* normA, const float* normB, float inv_reg,
                 const float* pot_in, const float* logw, float* pot_out, float* lse_out,
                 const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo,
                 void* ws, size_t ws_bytes, eg_stream_t stream);
size_t eg_lse_fused_workspace_bytes(int algo, int64_t nA, int64_t nB, int d);
/* hi = tf32(x) (round-to-nearest, low 13 mantissa bits zero), lo = tf32(x - hi);
 * rows are zero-padded from d to d_pad columns (d_pad % 8 == 0). */
int eg_split_tf32(const float* X, int64_t n, int d, int d_pad, float* hi, float* lo, eg_stream_t stream);
/* Fused plan statistics: *loss = sum_ij P_ij * cost_ij with P = exp(f_i + g_j - cost*inv_reg);
 * row_sum[i] (nullable) = sum_j P_ij (both zeroed by the call).  P (nullable) [nA, ldP]
 * is written only on request and only by EG_ALGO_SIMT; EG_ALGO_TCGEN05 takes the
 * split operands like eg_lse_fused. */
int eg_plan_fused(int algo, int cost, const float* A, int64_t nA, const float* B, int64_t nB, int d,
                  const float* normA, const float* normB, float inv_reg,
                  const float* f, const float* g, float* P, int64_t ldP, double* loss, float* row_sum,
                  const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo,
                  eg_stream_t stream);

/* Backward of the fused OT loss with the plan held fixed (models/models_ea.py:218-224: sinkhorn() is called on
 * M.detach(), the gradient flows through the cost only):
 *     dA[i, :] = scale * sum_j P_ij * d cost(A_i, B_j) / d A_i ,   P_ij = exp(f[i] + g[j] - cost_ij * inv_reg)
 * streamed over 64 x 64 tiles — neither the cost nor the plan is stored.  normA / normB as for eg_lse_fused
 * (squared norms for the L2 costs, norms for cosine).  d <= 320.  dB is the same call with (A, f) and (B, g)
 * exchanged.  Deterministic (partials summed in a fixed order). */
size_t eg_plan_grad_fused_workspace_bytes(int64_t nA, int64_t nB, int d);
int eg_plan_grad_fused(int cost, const float* A, int64_t nA, const float* B, int64_t nB, int d,
                       const float* normA, const float* normB, float inv_reg, const float* f, const float* g,
                       float scale, void* ws, size_t ws_bytes, float* dA, eg_stream_t stream);

/* ---- dense products of the layers on the same tcgen05 3xTF32 tiles ----------------
 * C[m, n] = [A1 | A2][m, k1+k2] · B[n, k1+k2]ᵀ + bias[n]   (fp32 accuracy, fp32 in/out)
 * Serves  x·Wᵀ + b  and  x·G + c  of layers/layers.py:61,69 in one launch
 * (B = [W ; Gᵀ], columns [0,n1) -> out1 = hidden, [n1,n) -> out2 = gate_pre) and their
 * input gradient  dx = [dH | d_gate] · [Wᵀ | G]ᵀ.  All operands are eg_split_tf32
 * outputs whose K extents (k1_pad, k2_pad, and B's k1_pad+k2_pad) are multiples of 16;
 * A2 may be NULL with k2_pad = 0; out2 may be NULL when n1 == n; n, n1, ld1, ld2 % 4 == 0. */
int eg_gemm_nt_3xtf32(const float* A1_hi, const float* A1_lo, int k1_pad,
                      const float* A2_hi, const float* A2_lo, int k2_pad, int64_t m,
                      const float* B_hi, const float* B_lo, int64_t n, const float* bias,
                      float* out1, int64_t ld1, int64_t n1, float* out2, int64_t ld2, eg_stream_t stream);
/* Same product with the A operands RAW (fp32): their hi/lo split is done inside the kernel (converter warps between
 * the TMA ring and the MMAs write the pair into tensor memory, from where tcgen05.mma reads its A operand), so the
 * streamed operand costs 8 KB instead of 16 KB per k-block, its ring is ten stages deep instead of four, and no
 * eg_split_tf32 pass over A is needed.  A1 [m, k1] and A2 [m, k2] row-major with row strides lda1 / lda2 (multiples
 * of 4, no K padding needed); B_hi / B_lo [n, ldb]: eg_split_tf32 of the weights with the two K parts back to back,
 * each zero-padded to a multiple of 16 columns; addend (nullable, requires n1 == n): [m, ld_add] added to the result.
 * Bit-identical to eg_split_tf32 + eg_gemm_nt_3xtf32. */
int eg_gemm_nt_3xtf32_raw(const float* A1, int64_t lda1, int k1, const float* A2, int64_t lda2, int k2, int64_t m,
                          const float* B_hi, const float* B_lo, int64_t ldb, int64_t n, const float* bias,
                          const float* addend, int64_t ld_add,
                          float* out1, int64_t ld1, int64_t n1, float* out2, int64_t ld2, eg_stream_t stream);
/* Same product with SHORT ACCUMULATION CHAINS (a fresh tensor-memory accumulator every 6 MMAs, folded into fp32
 * registers): ~3e-7 relative instead of ~2e-6, i.e. the accuracy of an fp32 SIMT product.  For the products
 * that feed a ReLU through the aggregation (layers/layers.py:32-38, 61-64), where a 2e-6 error flips the sign of
 * too many near-zero pre-activations for gradient parity. */
int eg_gemm_nt_3xtf32_chained(const float* A1_hi, const float* A1_lo, int k1_pad,
                              const float* A2_hi, const float* A2_lo, int k2_pad, int64_t m,
                              const float* B_hi, const float* B_lo, int64_t n, const float* bias,
                              float* out1, int64_t ld1, int64_t n1, float* out2, int64_t ld2, eg_stream_t stream);


/* Weight gradient of those products: C[m, n] = sum_k A[k, m] * B[k, n]  (dW = dH^T x; K = #entities).
 * A, B are the SAME row-major hi/lo split arrays ([K, lda], [K, ldb], lda/ldb % 4 == 0, 16-byte aligned) that
 * eg_gemm_nt_3xtf32 consumes — the tensor cores read them as MN-major operands, no transposed copy.  Split-K
 * over the grid; ws holds the per-split partial tiles (eg_gemm_tn_3xtf32_workspace_bytes), summed in split
 * order (deterministic). */
size_t eg_gemm_tn_3xtf32_workspace_bytes(int64_t K, int m, int n);
int eg_gemm_tn_3xtf32(const float* A_hi, const float* A_lo, int m, int lda,
                      const float* B_hi, const float* B_lo, int n, int ldb, int64_t K,
                      void* ws, size_t ws_bytes, float* out, int64_t ldo, eg_stream_t stream);

/* ---- margin-based L1 ranking loss with hard negatives -------------
 * Replaces the four [t*k, d] gathers + abs/sum/relu of models/models_ea.py:103-123,185-204:
 *   *loss_sum = sum_{p<t, q<k} relu(A_p + gamma - |out[nl]-out[nr]|_1) + relu(A_p + gamma - |out[n2l]-out[n2r]|_1),
 *   A_p = |out[left_p] - out[right_p]|_1     (the caller divides by 2*t*k).  Index arrays are int64,
 * negatives are laid out [t, k] row-major.  eg_margin_loss_bwd ADDS scale * d(loss_sum)/d(out) into
 * grad [n, d] (caller zeroes it); the sub-gradient of |.| at 0 is 0, as torch.abs.  d <= 512 for bwd. */
int eg_margin_loss_fwd(const float* out, int64_t n, int d, const int64_t* left, const int64_t* right,
                       const int64_t* nl, const int64_t* nr, const int64_t* n2l, const int64_t* n2r,
                       int64_t t, int k, float gamma, double* loss_sum, eg_stream_t stream);
int eg_margin_loss_bwd(const float* out, int64_t n, int d, const int64_t* left, const int64_t* right,
                       const int64_t* nl, const int64_t* nr, const int64_t* n2l, const int64_t* n2r,
                       int64_t t, int k, float gamma, float scale, float* grad, eg_stream_t stream);

/* ---- GAT edge-softmax aggregation (layers/att_layers.py:29-61) ----------
 *   w_ij = exp(-leakyrelu_alpha(s1_i + s2_j)) over the stored (i, j) of A (values of A unused),
 *   W_i = sum_j w_ij (-> wsum, nullable),  out_i = (sum_j m_ij w_ij h_j) / W_i   (pre-activation)
 * m = edge_scale (nullable, [nnz] in CSR order): the edge-dropout mask/scale the reference applies AFTER the
 * row sum (:50).
 * eg_gat_bwd_edges: from dy = dL/d out it writes p_edge[e] = m_ij w_ij / W_i (CSR edge order), ds1 [n_rows] and
 * ds2 [n_cols] (zeroed by the call, accumulated atomically); dh = P^T dy is then an ordinary eg_spmm over
 * CSR(A^T) whose values are p_edge permuted with eg_permute_edges(p_edge, perm_t).  d <= 512. */
int eg_gat_fwd(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* h, int d,
               const float* s1, const float* s2, float alpha, const float* edge_scale,
               float* out, float* wsum,
               /* hub-row segmentation, same meaning as eg_spmm's; seg_scratch holds n_seg * (d + 1) floats */
               int long_row_threshold, const int32_t* seg_row, const int32_t* seg_begin,
               const int32_t* seg_end, int64_t n_seg, const int32_t* long_rows, const int32_t* long_first,
               int64_t n_long, float* seg_scratch, eg_stream_t stream);
int eg_gat_bwd_edges(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols,
                     const float* h, int d, const float* s1, const float* s2, float alpha,
                     const float* edge_scale, const float* y, const float* wsum, const float* dy,
                     float* p_edge, float* ds1, float* ds2,
                     int long_row_threshold, const int32_t* seg_row, const int32_t* seg_begin,
                     const int32_t* seg_end, int64_t n_seg, eg_stream_t stream);
int eg_permute_edges(const float* src, const int32_t* perm, int64_t n, float* dst, eg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EAGRAFT_H_ */
