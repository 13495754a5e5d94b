// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM, 4 and 8 reader warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glue_factory_colon_b200/csrc/lg_tc_common.cuh"
__global__ void __launch_bounds__(256, 1) k(uint32_t* out, long long* cyc, int iters, int cols_per_iter) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols_per_iter; c += 32) {
      uint32_t r[32];
      tc::tmem_ld32(base + ((c + (warp >> 2) * 128) & 511), r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(slot, 512); }
}
int main() {
  uint32_t* d; long long* c; cudaMalloc(&d, 148 * 256 * 4); cudaMalloc(&c, 148 * 8);
  for (int threads : {128, 256}) {
    const int iters = 2000, cols = 128;
    k<<<148, threads>>>(d, c, 10, cols);
    k<<<148, threads>>>(d, c, iters, cols);
    long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)threads * cols * 4 * iters;  // per SM
    printf("%d reader threads: %lld cycles, %.1f B/clk/SM, %.1f cycles per ld32(+wait) per warp  err=%d\n", threads, h[0],
           bytes / h[0], (double)h[0] / (iters * cols / 32), (int)cudaGetLastError());
  }
  return 0;
}
