// Microbenchmark: issue rate of the instructions the attention softmax is made of, per SM sub-partition
// (warp-instructions per clock per SMSP), 16 warps per SM, 8 independent dependency chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, float seed) {
  float a[8], b = seed * 1.0001f, c = seed * 0.5f;
  unsigned long long d[8];
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i; d[i] = (unsigned long long)__float_as_uint(a[i]) << 32 | __float_as_uint(b); u[i] = i + (uint32_t)seed; }
  const unsigned long long bb = (unsigned long long)__float_as_uint(b) << 32 | __float_as_uint(c);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (OP == 0) {
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        REP8(X)
#undef X
      } else if (OP == 1) {
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        REP8(X)
#undef X
      } else if (OP == 2) {
#define X(i) asm volatile("fma.rn.f32 %0, %0, 0f3F800100, %1;" : "+f"(a[i]) : "f"(c));
        REP8(X)
#undef X
      } else if (OP == 3) {
#define X(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        REP8(X)
#undef X
      } else if (OP == 4) {
#define X(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        REP8(X)
#undef X
      } else if (OP == 5) {
#define X(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(bb));
        REP8(X)
#undef X
      } else if (OP == 6) {
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(bb));
        REP8(X)
#undef X
      } else if (OP == 7) {
#define X(i) asm volatile("mad.lo.u32 %0, %0, 8388608, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        REP8(X)
#undef X
      } else if (OP == 8) {
#define X(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        REP8(X)
#undef X
      } else if (OP == 9) {
#define X(i) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        REP8(X)
#undef X
      } else if (OP == 10) {  // FADD + MUFU 1:1
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
        REP8(X)
#undef X
      } else if (OP == 11) {  // FADD + FMNMX 1:1
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 7]) : "f"(c));
        REP8(X)
#undef X
      } else if (OP == 12) {  // 3 FADD : 1 MUFU
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 2) & 7]) : "f"(c)); \
             asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 5) & 7]) : "f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
        REP8(X)
#undef X
      } else if (OP == 13) {  // FADD + IADD 1:1
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        REP8(X)
#undef X
      } else if (OP == 14) {  // FFMA + FADD 1:1
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 4) & 7]) : "f"(b), "f"(c));
        REP8(X)
#undef X
      } else if (OP == 15) {  // max3
#define X(i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        REP8(X)
#undef X
      } else if (OP == 16) {  // mul
#define X(i) asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        REP8(X)
#undef X
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)(d[i] >> 32)) + __uint_as_float((uint32_t)d[i]) + __uint_as_float(u[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int per_x) {
  float* d; long long* c; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&c, 148 * 8);
  const int iters = 2000;
  k<OP><<<148, 512>>>(d, c, 10, 1.f);
  k<OP><<<148, 512>>>(d, c, iters, 1.f);
  long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  // 16 warps per SM = 4 per SMSP; each executes iters * 64 * per_x instructions
  const double instr_per_smsp = 4.0 * iters * 64 * per_x;
  printf("%-28s %.3f warp-instr/clk/SMSP  (%.2f clk per warp-instr)  err=%d\n", name, instr_per_smsp / h[0], h[0] / instr_per_smsp, (int)cudaGetLastError());
  cudaFree(d); cudaFree(c);
}
int main() {
  run<0>("FADD r,r", 1);
  run<16>("FMUL r,r", 1);
  run<1>("FFMA r,r,r", 1);
  run<2>("FFMA r,imm,r", 1);
  run<3>("FMNMX r,r", 1);
  run<15>("FMNMX3", 1);
  run<4>("MUFU.EX2", 1);
  run<5>("FADD2 (f32x2)", 1);
  run<6>("FFMA2 (f32x2)", 1);
  run<7>("IMAD (x*2^23+y)", 1);
  run<8>("IADD", 1);
  run<9>("F2FP bf16x2", 1);
  run<10>("FADD+MUFU 1:1", 2);
  run<12>("3 FADD + MUFU", 4);
  run<11>("FADD+FMNMX 1:1", 2);
  run<13>("FADD+IADD 1:1", 2);
  run<14>("FADD+FFMA 1:1", 2);
  return 0;
}
