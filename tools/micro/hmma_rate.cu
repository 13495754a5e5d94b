// Microbenchmark: legacy warp-level MMA (mma.sync -> HMMA) issue rate on sm_100a: tf32 m16n8k8, bf16 / f16 m16n8k16, fp32 accumulate.
// 8 independent accumulator tiles per warp, WARPS warps per SM sub-partition; reports cycles per MMA per SMSP and TFLOP/s at the run's clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cyc) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      if (MODE == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      if (MODE == 2)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_smsp, double flop_per_mma) {
  float* d; long long* c; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&c, 148 * 8);
  const int threads = warps_per_smsp * 4 * 32, iters = 20000;
  k<MODE><<<148, threads>>>(d, 100, c);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(d, iters, c);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  const double mma_per_smsp = (double)iters * 8 * warps_per_smsp;
  const double total_flop = mma_per_smsp * 4 * 148 * flop_per_mma;
  printf("%-22s %d warps/SMSP: %6.2f cycles per MMA per SMSP, %7.1f TFLOP/s  (%.3f ms, err %d)\n", name, warps_per_smsp,
         (double)h[0] / mma_per_smsp, total_flop / (ms * 1e-3) / 1e12, ms, (int)cudaGetLastError());
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int w : {1, 2, 4}) {
    run<0>("tf32 m16n8k8", w, 2.0 * 16 * 8 * 8);
    run<1>("bf16 m16n8k16", w, 2.0 * 16 * 8 * 16);
    run<2>("f16  m16n8k16", w, 2.0 * 16 * 8 * 16);
  }
  return 0;
}
