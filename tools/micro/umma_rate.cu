// Microbenchmark: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate) cycles per MMA on sm_100a by shape and operand
// source: SS (A and B from shared memory, 128-byte swizzle, K-major) against TS (A from TMEM), N = 64 / 128 / 256 at
// M = 128, one or two CTAs per SM.  One elected thread issues ITERS x 8 MMAs (8 K-steps of 16 over a 128-wide K tile, all
// into one accumulator), commits once, waits; cycles = clock64 around issue + completion.  Operands are zero-filled
// shared memory / TMEM (timing does not depend on the values).  Answers: does an N = 64 MMA cost half an N = 128 one, and
// are SS MMAs bound by the shared-memory operand fetch ((M + N) x 32 bytes per MMA against 128 B/clk)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I glue_factory_colon_b200/csrc tools/micro/umma_rate.cu -o umma_rate -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "lg_tc_common.cuh"

__global__ void __launch_bounds__(128) k(int N, int ts, int iters, int bmn, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A: 128 rows x 128 K (two 64-wide swizzle atoms, 16 KB each); B: up to 256 rows x 128 K
  for (int i = threadIdx.x; i < (32768 + 65536) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&slot, 256);  // 2 CTAs per SM fit
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = tc::idesc_bf16(128, N, bmn);
    const uint64_t dA = tc::smem_desc_sw128(tc::smem_u32(smem), 0, 1024);
    // bmn: B is MN-major (row = K index, 64 N-elements per 128-byte row; N = 64 only): K step = 16 rows = 2048 B
    const uint64_t dB = tc::smem_desc_sw128(tc::smem_u32(smem + 32768), bmn ? 8192 : 0, 1024);
    // accumulator at column 0 (N <= 256 columns would need 256: use N <= 128 for 2 CTAs/SM + TS), A planes at column 192
    const uint32_t tD = tmem, tA = tmem + 192;
    __syncwarp();
    const long long t0 = clock64();
    if (tc::elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a = dA + (uint64_t)((ks >> 2) * (16384 >> 4) + (ks & 3) * 2);
          const uint64_t b = bmn ? dB + (uint64_t)(ks * (2048 >> 4)) : dB + (uint64_t)((ks >> 2) * ((N * 128) >> 4) + (ks & 3) * 2);
          if (ts == 2) {  // A slice copied smem -> TMEM by tcgen05.cp (128 rows x 32 bytes = one K step), then the TS MMA
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tA + ks * 8), "l"(a) : "memory");
            tc::umma_ts(tD, tA + ks * 8, b, idesc, (it | ks) != 0);
          } else if (ts) tc::umma_ts(tD, tA + ks * 8, b, idesc, (it | ks) != 0);
          else tc::umma_ss(tD, a, b, idesc, (it | ks) != 0);
        }
      }
      tc::umma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 256);
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  const int iters = 2000;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int ts = 0; ts <= 1; ++ts)
      for (int N : {64, 128, 192}) {
        if (ts == 0 && N == 192) N = 256;
        if (ctas == 2 && N > 128) continue;      // 256 TMEM columns per CTA: accumulator + A planes
        if (ts == 1 && N > 128) continue;
        // dynamic smem: 96 KB (1 CTA/SM forced by asking for > half) or 96 KB with 2 CTAs/SM
        const int smem = 32768 + 65536;
        const int grid = 148 * ctas;
        if (ctas == 1) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, ctas == 1 ? smem + 65536 - 1024 : smem, 0>>>(N, ts, iters, 0, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[1024];
        cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += (double)h[i];
        avg /= grid;
        const double per = avg / (iters * 8.0);
        const double ideal = 128.0 * N * 16 * 2 / 8192.0;
        printf("%s M=128 N=%3d K=16, %d CTA/SM: %6.1f cycles per MMA per CTA (ideal at 8192 FLOP/clk/SM: %5.1f%s), operand bytes from smem per MMA %5d\n",
               ts ? "TS" : "SS", N, ctas, per, ideal, ctas == 2 ? " x2 when both CTAs issue" : "", (ts ? 0 : 128 * 32) + N * 32);
      }
  for (int N : {64, 128}) {  // tcgen05.cp of the A slice + TS MMA: does the copy overlap the previous MMA?
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 32768 + 65536 + 65536 - 1024, 0>>>(N, 2, iters, 0, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error (cp)\n"); return 1; }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    printf("cp+TS M=128 N=%3d K=16, 1 CTA/SM: %6.1f cycles per (tcgen05.cp 128x256b + MMA) (ideal %5.1f)\n", N, avg / 148 / (iters * 8.0), N / 2.0);
  }
  for (int ts = 0; ts <= 1; ++ts) {  // MN-major B (the P.V / dS.K form), N = 64, one CTA per SM
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 32768 + 65536 + 65536 - 1024, 0>>>(64, ts, iters, 1, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    printf("%s M=128 N= 64 K=16, B MN-major, 1 CTA/SM: %6.1f cycles per MMA (ideal 32.0)\n", ts ? "TS" : "SS", avg / 148 / (iters * 8.0));
  }
  return 0;
}
