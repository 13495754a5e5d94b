// Microbenchmark: per-element cost of the flash-attention softmax inner loop on sm_100a,
// as a function of resident warps per SM.  cycles/element/SM (lower is better).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
// 2^x for x <= 0 on the FMA/ALU pipes: round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax
// (max rel. err 1.0e-4), exponent patched in with integer ops.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05500889f, 0.24221097f);
  p = fmaf(p, f, 0.69328294f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ unsigned long long g_clk[4];
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  unsigned long long c0 = clock64(), g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
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
      if (MODE == 3) { p0 = ex2(v[2 * i] - m); p1 = (i & 1) ? ex2_poly(v[2 * i + 1] - m) : ex2(v[2 * i + 1] - m); }  // 25 % poly
      if (MODE == 4) { p0 = ex2(v[2 * i] - m); p1 = ex2_poly(v[2 * i + 1] - m); }                                    // 50 % poly
      if (MODE == 5) { p0 = ex2(v[2 * i] - m); p1 = (i % 3 == 0) ? ex2_poly(v[2 * i + 1] - m) : ex2(v[2 * i + 1] - m); }  // 17 %
      if (MODE == 6 || MODE == 7) {  // packed fp32x2 (FADD2 / FFMA2) for the subtract, the sum and the polynomial
        const float2 x2 = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), make_float2(-m, -m));
        if (MODE == 7 && (i & 1)) {
          // both lanes by polynomial: 2 of every 4 elements... (i odd) -> 25 % of pairs fully poly = 25 % elements
          float2 xc = make_float2(fmaxf(x2.x, -126.f), fmaxf(x2.y, -126.f));
          const float2 t = __fadd2_rn(xc, make_float2(12582912.f, 12582912.f));
          const float2 f = __fadd2_rn(xc, __fmul2_rn(__fadd2_rn(t, make_float2(-12582912.f, -12582912.f)), make_float2(-1.f, -1.f)));
          float2 p = __ffma2_rn(f, make_float2(0.05500889f, 0.05500889f), make_float2(0.24221097f, 0.24221097f));
          p = __ffma2_rn(p, f, make_float2(0.69328294f, 0.69328294f));
          p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
          p0 = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
          p1 = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
        } else { p0 = ex2(x2.x); p1 = ex2(x2.y); }
        const float2 r2 = __fadd2_rn(make_float2(rs[(i & 1) * 2], rs[(i & 1) * 2 + 1]), make_float2(p0, p1));
        rs[(i & 1) * 2] = r2.x; rs[(i & 1) * 2 + 1] = r2.y;
      }
      if (MODE != 2 && MODE != 6 && MODE != 7) rs[i & 3] += p0 + p1;
      const uint32_t pk = pack(p0, p1);
      x ^= pk;
      v[2 * i] = p0 * 0.25f - 3.f; v[2 * i + 1] = p1 * 0.25f - 2.f;  // feed back so nothing is hoisted
    }
    acc += (rs[0] + rs[1]) + (rs[2] + rs[3]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(x & 0x7fffff);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    g_clk[0] = clock64() - c0; g_clk[1] = g1 - g0;
  }
}
template <int MODE> void run(const char* n, int threads) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(d, 10, 1.f);
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long hc[4]; cudaMemcpyFromSymbol(hc, g_clk, sizeof(hc));
  double elems = (double)threads * 64 * iters;  // per SM
  printf("%s threads/SM=%4d: %.3f ms, SM clock %.0f MHz -> elements/clk/SM = %.2f\n", n, threads, ms,
         hc[0] * 1e3 / (double)hc[1], elems / (double)hc[0]);
  cudaFree(d);
}
int main() {
  for (int t : {256, 512}) { run<0>("ex2+sum+pack   ", t); run<3>("25% poly       ", t); run<6>("f32x2 no poly  ", t); run<7>("f32x2 25% poly ", t); }
  return 0;
}
