// Microbenchmark: which (lane, column) of TMEM does each (thread, register) of the 16x256b load / 16x128b store
// shapes touch?  TMEM is filled through the 32x32b shape (thread = lane, register = column) with lane * 1000 + column.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glue_factory_colon_b200/csrc/lg_tc_common.cuh"
__global__ void __launch_bounds__(128, 1) k(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tc::tmem_alloc(&slot, 128);
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
  for (int c = 0; c < 4; ++c) {
    for (int i = 0; i < 32; ++i) v[i] = (uint32_t)((warp * 32 + lane) * 1000 + c * 32 + i);
    tc::tmem_st32(base + c * 32, v);
  }
  tc::tmem_st_wait();
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  // 16x256b.x2: 16 lanes x 16 columns -> 8 registers per thread
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(base + ((uint32_t)16 << 16) + 32));  // lanes 16..31 of the quarter, columns 32..47
  tc::tmem_ld_wait();
  for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 16 + i] = r[i];
  // 16x128b.x2 store: 16 lanes x 8 columns -> 4 registers per thread; write thread/register ids, read back with 32x32b
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  uint32_t w[4];
  for (int i = 0; i < 4; ++i) w[i] = 900000u + lane * 10 + i;
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + 64), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  tc::tmem_st_wait();
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  uint32_t q[32];
  tc::tmem_ld32(base + 64, q);
  tc::tmem_ld_wait();
  for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 16 + 8 + i] = q[i];
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(slot, 128); }
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 16 * 4);
  k<<<1, 128>>>(d);
  static uint32_t h[128 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("err=%d\n16x256b.x2 load at (lane 16 of quarter, column 32): thread -> 8 registers as lane*1000+column\n", (int)cudaGetLastError());
  for (int t = 0; t < 32; ++t) { printf("t%2d:", t); for (int i = 0; i < 8; ++i) printf(" %6u", h[t * 16 + i]); printf("\n"); }
  printf("16x128b.x2 store at column 64 (lanes 0..15), read back 32x32b: lane -> columns 64..71 as 900000+thread*10+reg\n");
  for (int t = 0; t < 16; ++t) { printf("lane%2d:", t); for (int i = 0; i < 8; ++i) printf(" %6u", h[t * 16 + 8 + i]); printf("\n"); }
  return 0;
}
