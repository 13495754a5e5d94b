// Microbenchmark: tcgen05.mma.cta_group::2 (kind::f16, bf16, M = 256 over a CTA pair, K = 16) cycles per MMA for N = 64 /
// 128 / 256 in the SS form (each CTA supplies its 128 A rows, 4 KB per K step, and its half of B) and the TS form (A from
// each CTA's TMEM).  The leader CTA's elected thread issues ITERS x 8 MMAs back to back, commits to a barrier in both
// CTAs, both wait.  Companion of umma_rate.cu (cta_group::1).  B200: nominal N/2 cycles everywhere except SS N = 64 (43).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I glue_factory_colon_b200/csrc tools/micro/umma_rate_pair.cu -o umma_rate_pair -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "lg_tc_common.cuh"

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TS is a compile-time switch: predicated-off MMAs left in the issue stream by a runtime switch are not free
template <int TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k(int N, int iters, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_rank();
  // A: 128 rows x 128 K (two swizzle atoms of 16 KB); B half: N/2 rows x 128 K (two atoms of N/2 * 128 bytes)
  for (int i = threadIdx.x; i < (32768 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    __syncwarp();
    t0 = clock64();
    if (crank == 0) {
      const uint32_t idesc = tc::idesc_bf16(256, N, 0);
      const uint64_t dA = tc::smem_desc_sw128(tc::smem_u32(smem), 0, 1024);
      const uint64_t dB = tc::smem_desc_sw128(tc::smem_u32(smem + 32768), 0, 1024);
      if (tc::elect_one()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a = dA + (uint64_t)((ks >> 2) * (16384 >> 4) + (ks & 3) * 2);
            const uint64_t b = dB + (uint64_t)((ks >> 2) * (((N / 2) * 128) >> 4) + (ks & 3) * 2);
            const uint32_t acc = (it | ks) != 0;
            if constexpr (TS)  // A from TMEM (each CTA's own 128 rows at columns 192..), N <= 128
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "setp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                  "r"(tmem + 192 + ks * 8), "l"(b), "r"(idesc), "r"(acc)
                  : "memory");
            else
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "setp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                  "l"(a), "l"(b), "r"(idesc), "r"(acc)
                  : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         tc::smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
      }
      __syncwarp();
    }
    tc::mbar_wait(&bar, 0);
    t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
  }
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc::fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  const int iters = 2000;
  const int smem = 163840 - 2048;  // one CTA per SM
  for (int cfg = 0; cfg < 5; ++cfg) {
    const int N = cfg == 0 ? 128 : cfg == 1 ? 256 : cfg == 2 ? 64 : cfg == 3 ? 64 : 128;
    const int ts = cfg >= 3;
    const int grid = 148;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) {
      if (ts) k<1><<<grid, 128, smem, 0>>>(N, iters, cyc);
      else k<0><<<grid, 128, smem, 0>>>(N, iters, cyc);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    int n = 0;
    for (int i = 0; i < grid; i += 2) { avg += (double)h[i]; ++n; }  // leaders
    const double per = avg / n / (iters * 8.0);
    printf("%s cta_group::2 M=256 N=%3d K=16: %6.1f cycles per MMA (ideal at 8192 FLOP/clk/SM: %5.1f), per CTA: A %d B + B %d B from shared memory\n",
           ts ? "TS" : "SS", N, per, N / 2.0, ts ? 0 : 4096, (N / 2) * 32);
  }
  return 0;
}
