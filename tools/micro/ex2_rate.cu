// Microbenchmark: MUFU ex2 throughput for f32 vs packed f16x2 / bf16x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003800u + threadIdx.x + i;  // two small halves / a small float
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  k<MODE><<<148 * 8, 512>>>(d, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 512>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = 148.0 * 8 * 512 * iters * 8;  // thread-level MUFU ops
  printf("%s: %.3f ms  %.2f thread-ops/clk/SM (at 1.965 GHz)  err=%d\n", name, ms, instr / (ms * 1e-3) / 148 / 1.965e9, (int)cudaGetLastError());
  cudaFree(d);
}
int main() { run<0>("ex2.f32   "); run<1>("ex2.f16x2 "); run<2>("ex2.bf16x2"); return 0; }
