// Issue-rate probes: the roofline denominators of the SIMT-bound kernels (L1 evaluation on the FP32 / FP64 pipes,
// the exponentials of the log-domain Sinkhorn on the MUFU), measured on the device the library runs on instead of
// taken from a data sheet.  Measurement aid only (tools/measure_peaks.py, bench.py's roofline blocks).
#include <cuda_runtime.h>

#include "common.cuh"

namespace eg {

// KIND 0: fp32 FMA (8 independent chains per thread)   -> lane-ops = FMAs
// KIND 1: MUFU ex2.approx (8 independent chains)        -> lane-ops = ex2
// KIND 2: fp64 add (8 independent chains)               -> lane-ops = DADDs
// KIND 3: fp32 add + abs (the |a-b| + acc pattern of the L1 filter: FADD, then FADD with |.| modifier)
template <int KIND>
__global__ void __launch_bounds__(256) issue_probe_kernel(float* sink, int iters, float seed) {
  float x[8];
  double d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { x[k] = seed + (float)(threadIdx.x + k) * 1e-3f; d[k] = (double)x[k]; }
  const float a = 1e-9f * seed;
  const double db = 1e-9 * (double)seed;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (KIND == 0) x[k] = fmaf(x[k], a, x[k]);     // two distinct source registers: no bank conflict on the operand fetch
        else if (KIND == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
        else if (KIND == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(db));
        else x[k] += fabsf(x[k] - 0.999f);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k] + (float)d[k];
  if (s == 123.456f) sink[0] = s;      // never true in practice: keeps the chains alive
}

}  // namespace eg

using namespace eg;

extern "C" int eg_issue_peak(int kind, int iters, double* h_lane_ops_per_s, void* scratch /* >= 4 B */,
                             eg_stream_t stream_) {
  if (kind < 0 || kind > 3 || iters <= 0 || !h_lane_ops_per_s || !scratch) return EG_ERR_INVALID;
  cudaStream_t s = as_stream(stream_);
  const int ctas = kNumSMs * 8, threads = 256;          // 2048 threads per SM: every scheduler has warps to pick from
  cudaEvent_t e0, e1;
  EG_CUDA(cudaEventCreate(&e0));
  EG_CUDA(cudaEventCreate(&e1));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {                   // first repetition is the warm-up
    EG_CUDA(cudaEventRecord(e0, s));
    float* sink = reinterpret_cast<float*>(scratch);
    switch (kind) {
      case 0: issue_probe_kernel<0><<<ctas, threads, 0, s>>>(sink, iters, 1.0f); break;
      case 1: issue_probe_kernel<1><<<ctas, threads, 0, s>>>(sink, iters, 1.0f); break;
      case 2: issue_probe_kernel<2><<<ctas, threads, 0, s>>>(sink, iters, 1.0f); break;
      default: issue_probe_kernel<3><<<ctas, threads, 0, s>>>(sink, iters, 1.0f); break;
    }
    EG_LAUNCHED();
    EG_CUDA(cudaEventRecord(e1, s));
    EG_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    EG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double per_thread = (double)iters * 64.0 * (kind == 3 ? 2.0 : 1.0);
  *h_lane_ops_per_s = per_thread * (double)ctas * threads / ((double)best_ms * 1e-3);
  return EG_OK;
}
