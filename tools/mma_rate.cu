// Micro-benchmark: how long does one tcgen05.mma.kind::tf32 of shape 128 (or 256 per CTA pair) x N x 8 take when nothing
// else is in the way?  One CTA (or CTA pair) per SM issues `iters` MMAs back to back into one accumulator; operands are
// whatever shared / tensor memory holds (zeros).  Second half: the same MMAs spread over 2 or 3 accumulators used in turn.  Variants: A from shared memory (SS) or tensor memory (TS), cta_group 1 / 2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {   // K-major, 64-byte swizzle, 8-row groups of 512 B
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int PAIR, int TS, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  uint32_t crank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_slot;
  if (threadIdx.x == 0 && crank == 0) {
    const uint32_t idesc = make_idesc(PAIR ? 256 : 128, N);
    const uint64_t adesc = make_desc(smem_u32(smem));             // 128 rows x 64 B = 8 KB
    const uint64_t bdesc = make_desc(smem_u32(smem + 16384));     // up to 256 rows x 64 B
    const uint32_t a_tmem = tbase + 384u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {                                          // 6 MMAs per trip, as one k-block of the GEMMs
        const uint64_t koff = (uint64_t)((k & 1) * 2);
        const uint32_t d = tbase + (uint32_t)((k % NACC) * N);                 // NACC accumulators, used in turn
        if (PAIR) {
          if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                               "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5,%5,%5,%5,%5,%5,%5,%5}, p;\n\t}"
                               ::"r"(d), "r"(a_tmem + (uint32_t)(8 * (k & 1))), "l"(bdesc + koff), "r"(idesc), "r"(1u), "r"(0u) : "memory");
          else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5,%5,%5,%5,%5,%5,%5,%5}, p;\n\t}"
                            ::"r"(d), "l"(adesc + koff), "l"(bdesc + koff), "r"(idesc), "r"(1u), "r"(0u) : "memory");
        } else {
          if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                               ::"r"(d), "r"(a_tmem + (uint32_t)(8 * (k & 1))), "l"(bdesc + koff), "r"(idesc), "r"(1u) : "memory");
          else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(d), "l"(adesc + koff), "l"(bdesc + koff), "r"(idesc), "r"(1u) : "memory");
        }
      }
    }
    if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                           ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
  }
}

template <int PAIR, int TS, int NACC>
static void run(int N, long long* cyc) {
  const int iters = 1000, ksteps = 6, nacc = NACC, smem = 64 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<PAIR, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<PAIR, TS, NACC>, N, iters, cyc);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  %s\n", cudaGetErrorString(e)); return; }
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    best = h / (double)(iters * ksteps) < best ? h / (double)(iters * ksteps) : best;
  }
  printf("cta_group::%d %s N=%3d, %d accumulator(s) in turn: %.1f cycles per MMA (128 rows x N x 8 per SM) = %.0f MAC/clk/SM\n", PAIR ? 2 : 1, TS ? "TS" : "SS", N, nacc, best,
         128.0 * N * 8 / best);
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int N : {64, 96, 128, 160, 192, 256}) {
    run<0, 0, 1>(N, cyc); run<0, 1, 1>(N, cyc);
    if (N % 32 == 0) { run<1, 0, 1>(N, cyc); run<1, 1, 1>(N, cyc); }
  }
  // the same with the MMAs spread over 2 or 3 accumulators used in turn: is the floor above a dependency or an issue cost?
  for (int N : {64, 96, 128, 160, 192}) {
    run<0, 0, 2>(N, cyc); run<0, 1, 2>(N, cyc);
    if (N % 32 == 0) run<1, 1, 2>(N, cyc);
  }
  for (int N : {64, 96, 128}) {
    run<0, 0, 3>(N, cyc); run<0, 1, 3>(N, cyc);
    if (N % 32 == 0) run<1, 1, 3>(N, cyc);
  }
  return 0;
}
