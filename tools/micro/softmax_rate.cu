// Microbenchmark: per-element cost of the flash-attention softmax inner loop on sm_100a,
// as a function of resident warps per SM.  cycles/element/SM (lower is better).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float v[64];
  for (int i = 0; i < 64; ++i) v[i] = seed * (threadIdx.x + i) * 1e-3f;
  float m = 0.5f, acc = 0.f; uint32_t x = 0;
  for (int it = 0; it < iters; ++it) {
    float mx[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
    for (int i = 4; i < 64; ++i) mx[i & 3] = fmaxf(mx[i & 3], v[i]);
    m = fmaxf(m, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) - 8.f);
    float rs[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float p0, p1;
      if (MODE == 0) { p0 = ex2(v[2 * i] - m); p1 = ex2(v[2 * i + 1] - m); }
      if (MODE == 1) { p0 = (v[2 * i] - m); p0 *= p0; p1 = (v[2 * i + 1] - m); p1 *= p1; }
      if (MODE == 2) { p0 = ex2(v[2 * i] - m); p1 = ex2(v[2 * i + 1] - m); }
      if (MODE != 2) rs[i & 3] += p0 + p1;
      const uint32_t pk = pack(p0, p1);
      x ^= pk;
      v[2 * i] = p0 * 0.25f - 3.f; v[2 * i + 1] = p1 * 0.25f - 2.f;  // feed back so nothing is hoisted
    }
    acc += (rs[0] + rs[1]) + (rs[2] + rs[3]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(x & 0x7fffff);
}
template <int MODE> void run(const char* n, int threads) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(d, 10, 1.f);
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double elems = (double)threads * 64 * iters;  // per SM
  printf("%s threads/SM=%4d: %.3f ms -> %.2f ns per 1024 elements/SM; elements/clk/SM @1.9GHz = %.1f\n", n, threads, ms,
         ms * 1e6 / (elems / 1024), elems / (ms * 1e-3 * 1.9e9));
  cudaFree(d);
}
int main() {
  for (int t : {128, 256, 512, 1024}) { run<0>("ex2+sum+pack   ", t); run<1>("mul+sum+pack   ", t); run<2>("ex2+pack(nosum)", t); }
  return 0;
}
