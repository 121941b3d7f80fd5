// Shared helpers for the eagraft kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "eagraft.h"

namespace eg {

extern std::atomic<int64_t> g_launches;
extern thread_local int t_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  t_last_cuda_error = static_cast<int>(e);
  return EG_ERR_CUDA;
}

#define EG_CUDA(expr)                                   \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return ::eg::cuda_fail(_e);  \
  } while (0)

// Count + check a kernel launch.
#define EG_LAUNCHED()                                         \
  do {                                                        \
    ::eg::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return ::eg::cuda_fail(_e);        \
  } while (0)

inline cudaStream_t as_stream(eg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// Function attributes are per device: remember, per kernel call site and device, whether the opt-in dynamic
// shared-memory size has been set (ops.py serves several devices in one process).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done[4] = {};     // 256 devices
  bool test(int dev) const { return dev >= 0 && dev < 256 && ((done[dev >> 6].load(std::memory_order_acquire) >> (dev & 63)) & 1ull); }
  void set(int dev) { if (dev >= 0 && dev < 256) done[dev >> 6].fetch_or(1ull << (dev & 63), std::memory_order_release); }
};
#define EG_SET_SMEM_ONCE(once, ...)                                   \
  do {                                                                \
    int _dev = -1;                                                    \
    EG_CUDA(cudaGetDevice(&_dev));                                    \
    if (!(once).test(_dev)) {                                         \
      __VA_ARGS__;                                                    \
      (once).set(_dev);                                               \
    }                                                                 \
  } while (0)

// 128-bit streaming loads/stores: bypass L1 allocation for data touched once.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ int ld_stream_i32(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

}  // namespace eg
