// Shared pieces of the persistent Sinkhorn solvers (sinkhorn_dense.cu, sinkhorn_tile2d.cu).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace eg {

struct PersistState {
  unsigned int barrier;      // monotonically increasing arrival counter
  int sweeps;                // sweeps completed (host output)
  int final_buf;             // which log-v buffer holds the accepted iterate
  double err;                // last marginal error evaluated
  double err2[128];          // one accumulator per check (sweeps 0,10,...)
  unsigned long long t_phase[8];   // ns spent by CTA 0 in the phases of a sweep (diagnostic)
  int fallback;              // scaling-domain kernels: a sum left the fp32 range -> the solve is redone in the log domain
  int absorb_req;            // row-block scaling kernel: sweep (+1) at which every CTA folds u, v into its kernel entries
  int absorbs;               // how many times that happened (diagnostic)
  int flag_code;             // tile kernel: 0x7fffffff - (sweep at which `fallback` was first raised), 0 = never
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Diagnostic counters live on the device (the solve does not synchronise with the host); these read them back.
int sinkhorn_tile2d_absorbs_read();

// sinkhorn_tile2d.cu: scaling-domain continuation with the kernel matrix tiled over (cluster, CTA-in-cluster).
// Returns EG_OK and sets *launched when the shape / device allow it; *launched = false -> caller uses another path.
int sinkhorn_tile2d_launch(const float* M, int64_t I, int64_t J, int64_t ld, double inv_reg, const float* a,
                           const float* b, float* log_u, float* log_v, const PersistState* warm, int start_iter,
                           int max_iter, double stop_thr, void* part, size_t part_bytes, PersistState* st,
                           float absorb_log2, int force_fallback, cudaStream_t s, bool* launched);

// Communication skeleton of one tile-kernel sweep (no mat-vec work), `iters` times: the latency floor of the design.
int sinkhorn_tile2d_sync_floor_launch(int64_t I, int64_t J, int iters, void* part, size_t part_bytes,
                                      PersistState* st, cudaStream_t s, bool* launched);

}  // namespace eg
