// Microbenchmark: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate) cycles per MMA on sm_100a by shape and operand
// source: SS (A and B from shared memory, 128-byte swizzle) against TS (A from TMEM) and tcgen05.cp + TS, N = 64 / 128 / 256
// at M = 128, one or two CTAs per SM, K-major and MN-major B.  One elected thread issues ITERS x 8 MMAs (8 K-steps of 16
// over a 128-wide K tile, all into one accumulator), commits once, waits; cycles = clock64 around issue + completion.
// Operands are zero-filled shared memory / whatever TMEM holds (timing does not depend on the values).
// B200: every form runs at the nominal N/2 cycles except SS N = 64 (48: 6 KB of shared-memory operands per MMA at 128 B/clk).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I glue_factory_colon_b200/csrc tools/micro/umma_rate.cu -o umma_rate -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "lg_tc_common.cuh"

// MODE (compile time: a runtime switch leaves predicated-off UTCHMMA / UTCCP instructions in the issue stream, and those
// are not free -- the first version of this benchmark measured them): 0 = SS, 1 = TS, 2 = tcgen05.cp + TS
template <int MODE>
__global__ void __launch_bounds__(128) k(int N, int iters, int bmn, long long* cyc) {
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
          if constexpr (MODE == 2) {  // A slice copied smem -> TMEM by tcgen05.cp (128 rows x 32 bytes = one K step), then the TS MMA
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tA + ks * 8), "l"(a) : "memory");
            tc::umma_ts(tD, tA + ks * 8, b, idesc, (it | ks) != 0);
          } else if constexpr (MODE == 1) tc::umma_ts(tD, tA + ks * 8, b, idesc, (it | ks) != 0);
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

template <int MODE>
static double run(int N, int bmn, int ctas, int iters, long long* cyc) {
  const int smem = 32768 + 65536;
  const int grid = 148 * ctas;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<grid, 128, ctas == 1 ? smem + 65536 - 1024 : smem, 0>>>(N, iters, bmn, cyc);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  long long h[1024];
  cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += (double)h[i];
  return avg / grid / (iters * 8.0);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  const int iters = 2000;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    for (int N : {64, 128, 256}) {
      if (ctas == 2 && N > 128) continue;  // 256 TMEM columns per CTA
      printf("SS    M=128 N=%3d K=16, %d CTA/SM: %6.1f cycles per MMA per CTA (ideal %5.1f%s)\n", N, ctas, run<0>(N, 0, ctas, iters, cyc), N / 2.0,
             ctas == 2 ? " x2: both CTAs issue" : "");
    }
    for (int N : {64, 128}) printf("TS    M=128 N=%3d K=16, %d CTA/SM: %6.1f cycles per MMA per CTA (ideal %5.1f%s)\n", N, ctas, run<1>(N, 0, ctas, iters, cyc), N / 2.0,
                                   ctas == 2 ? " x2: both CTAs issue" : "");
  }
  for (int N : {64, 128}) printf("cp+TS M=128 N=%3d K=16, 1 CTA/SM: %6.1f cycles per (tcgen05.cp 128x256b + MMA) (ideal %5.1f)\n", N, run<2>(N, 0, 1, iters, cyc), N / 2.0);
  printf("SS    M=128 N= 64 K=16, B MN-major, 1 CTA/SM: %6.1f cycles per MMA (ideal 32.0)\n", run<0>(64, 1, 1, iters, cyc));
  printf("TS    M=128 N= 64 K=16, B MN-major, 1 CTA/SM: %6.1f cycles per MMA (ideal 32.0)\n", run<1>(64, 1, 1, iters, cyc));
  return 0;
}
