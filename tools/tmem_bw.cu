// Micro-benchmark: how fast can 16 warps of one CTA stream a register-shaped working set out of tensor memory
// (tcgen05.ld 32x32b) compared with shared memory (LDS.128)?  Decides whether the slow rows of the tile Sinkhorn
// kernel should live in TMEM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: TMEM x16 loads, 1: SMEM LDS.128, 2: both interleaved (TMEM 1 : SMEM 2)
__global__ void __launch_bounds__(512, 1) bw_kernel(float* out, int iters, long long* cycles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float4* sm = reinterpret_cast<float4*>(smem_raw);
  for (int i = tid; i < 8192; i += 512) sm[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_slot;
  // this warp's TMEM window: lanes 32*(warp%4).., columns 128*(warp/4)..+127
  const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
  // fill my window with something
  for (int c = 0; c < 128; c += 16) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 ::"r"(taddr + c), "r"(__float_as_uint(1.0f + lane)) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      // 84 columns ~ 6 x16 loads -> here 6 loads of 16 columns (96 columns)
#pragma unroll
      for (int c = 0; c < 96; c += 48) {
        uint32_t r[3][16];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(r[j][0]), "=r"(r[j][1]), "=r"(r[j][2]), "=r"(r[j][3]), "=r"(r[j][4]), "=r"(r[j][5]), "=r"(r[j][6]), "=r"(r[j][7]),
                         "=r"(r[j][8]), "=r"(r[j][9]), "=r"(r[j][10]), "=r"(r[j][11]), "=r"(r[j][12]), "=r"(r[j][13]), "=r"(r[j][14]), "=r"(r[j][15])
                       : "r"(taddr + c + 16 * j));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int k = 0; k < 16; ++k) acc += __uint_as_float(r[j][k]);
      }
    }
    if (MODE == 1 || MODE == 2) {
      // same bytes from shared memory: 24 (MODE 1) or 48 (MODE 2) LDS.128 per thread
#pragma unroll
      for (int c = 0; c < (MODE == 1 ? 24 : 48); ++c) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(smem_u32(&sm[(warp * 8 + (c & 7)) * 32 + lane + ((c >> 3) & 1) * 4096])));
        acc += v.x + v.y + v.z + v.w;
      }
    }
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  out[blockIdx.x * 512 + tid] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) { cudaFuncSetAttribute(bw_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16); bw_kernel<0><<<148, 512, 8192 * 16>>>(out, iters, cyc); }
      if (mode == 1) { cudaFuncSetAttribute(bw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16); bw_kernel<1><<<148, 512, 8192 * 16>>>(out, iters, cyc); }
      if (mode == 2) { cudaFuncSetAttribute(bw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16); bw_kernel<2><<<148, 512, 8192 * 16>>>(out, iters, cyc); }
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double tmem_bytes = (mode == 1) ? 0 : 512.0 * 96 * 4, smem_bytes = (mode == 0) ? 0 : 512.0 * (mode == 1 ? 24 : 48) * 16;
      printf("mode %d rep %d: %s, %.1f cycles/iter, TMEM %.1f B/clk/SM, SMEM %.1f B/clk/SM\n", mode, rep, cudaGetErrorString(e),
             (double)h / iters, tmem_bytes * iters / h, smem_bytes * iters / h);
    }
  }
  return 0;
}
