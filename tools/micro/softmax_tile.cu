// Microbenchmark: the flash-attention softmax step exactly as the kernel runs it, minus the MMAs and mbarriers:
// tcgen05.ld 64 score columns -> row max (+ exchange with the partner warp) -> 2^x -> row sum -> bf16 pack ->
// tcgen05.st 32 P columns.  8 softmax warps per CTA, 2 CTAs per SM.  Reports cycles per 128x128 tile per SM
// (the tensor pipe needs 512 for QK^T + PV at d = 64; MUFU alone needs 1024 at 16 ex2/clk/SM).
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../glue_factory_colon_b200/csrc/lg_tc_common.cuh"

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05500889f, 0.24221097f);
  p = fmaf(p, f, 0.69328294f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// both lanes by polynomial, packed fp32x2 arithmetic
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.05500889f, 0.05500889f), make_float2(0.24221097f, 0.24221097f));
  p = __ffma2_rn(p, f, make_float2(0.69328294f, 0.69328294f));
  p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// MODE 0: kernel as of r1 (scalar, 25 % poly)   1: scalar no poly
//      2: f32x2 sub/sum, no poly                3..6: f32x2, NP of every 8 pairs fully by polynomial (NP = MODE-2 -> 12.5 .. 50 %)
//      7: f32x2 sub/sum + scalar 25 % poly (one lane of every other pair)
template <int MODE>
__global__ void __launch_bounds__(320, 2) k(float* out, long long* cyc, int iters, long long* prog) {
  __shared__ uint32_t slot;
  __shared__ float xch[2][2][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tc::tmem_alloc(&slot, 256);
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const int quarter = warp & 3, half = warp >> 2, r = quarter * 32 + lane;
  const uint32_t lane_base = slot + ((uint32_t)(quarter * 32) << 16);
  {  // defined TMEM contents: scores in [-4, 4)
    uint32_t z[32];
    for (int c = 0; c < 4; ++c) {
      for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(((threadIdx.x * 37 + (c * 32 + i) * 11) % 64) * 0.125f - 4.f);
      if (half == (c >> 1)) tc::tmem_st32(lane_base + c * 32, z);
    }
    tc::tmem_st_wait();
  }
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  float m_ref = -INFINITY, l_part = 0.f;
  long long t0 = clock64();
  unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  for (int j = 0; j < iters; ++j) {
    if (threadIdx.x == 0 && (j & 63) == 0) prog[blockIdx.x * 64 + (j >> 6)] = (long long)clock64();
    uint32_t sv[64];
    tc::tmem_ld32(lane_base + half * 64, sv);
    tc::tmem_ld32(lane_base + half * 64 + 32, sv + 32);
    tc::tmem_ld_wait();
    float mxs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) mxs[i] = __uint_as_float(sv[i]);
#pragma unroll
    for (int i = 4; i < 64; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sv[i]));
    float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
    xch[j & 1][half][r] = mx;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
    mx = fmaxf(mx, xch[j & 1][half ^ 1][r]);
    float m_new = m_ref;
    if (mx > m_ref + 8.f) m_new = mx;
    const float alpha = ex2(m_ref - m_new);
    uint32_t pk[32];
    float tile_sum;
    if (MODE <= 1) {
      float rsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float p0 = __uint_as_float(sv[2 * i]) - m_new, p1 = __uint_as_float(sv[2 * i + 1]) - m_new;
        p0 = ex2(p0);
        p1 = (MODE == 0 && i % 2 == 0) ? ex2_poly(p1) : ex2(p1);
        rsum[i & 3] += p0 + p1;
        pk[i] = tc::pack_bf16(p0, p1);
      }
      tile_sum = (rsum[0] + rsum[1]) + (rsum[2] + rsum[3]);
    } else {
      float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 nm = make_float2(-m_new, -m_new);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float2 x = __fadd2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), nm);
        float2 p;
        constexpr int NP = MODE - 2;
        if (MODE >= 3 && MODE <= 6 && ((i * NP) % 8) < NP) {
          p = ex2_poly2(x);
        } else if (MODE == 7 && i % 2 == 0) {
          p = make_float2(ex2(x.x), ex2_poly(x.y));
        } else {
          p = make_float2(ex2(x.x), ex2(x.y));
        }
        acc[i & 1] = __fadd2_rn(acc[i & 1], p);
        pk[i] = tc::pack_bf16(p.x, p.y);
      }
      tile_sum = (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y);
    }
    l_part = l_part * alpha + tile_sum;
    m_ref = m_new;
    tc::tmem_st32(lane_base + 128 + half * 32, pk);
    tc::tmem_st_wait();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l_part + m_ref;
  if (threadIdx.x == 0) {
    prog[blockIdx.x * 64 + 63] = t1;
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cyc[blockIdx.x * 4] = t1 - t0; cyc[blockIdx.x * 4 + 1] = (long long)g0; cyc[blockIdx.x * 4 + 2] = (long long)g1; cyc[blockIdx.x * 4 + 3] = smid;
  }
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(slot, 256); }
}

template <int MODE> void run(const char* name) {
  float* d; long long *c, *pr; cudaMalloc(&d, 296 * 256 * 4); cudaMalloc(&c, 296 * 8 * 4); cudaMalloc(&pr, 296 * 64 * 8);
  const int iters = 63 * 64;  // stamps every 64 tiles: 63 + the end stamp
  k<MODE><<<296, 256>>>(d, c, 10, pr);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<296, 256>>>(d, c, iters, pr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long h[296 * 4], hp[296 * 64]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hp, pr, sizeof(hp), cudaMemcpyDeviceToHost);
  float chk; cudaMemcpy(&chk, d + 5, 4, cudaMemcpyDeviceToHost);
  // steady state: for every SM, tiles finished by both resident CTAs inside the window in which both were running
  // (clock64 is per SM, so stamps of co-resident CTAs are comparable)
  int partner[256]; for (int i = 0; i < 256; ++i) partner[i] = -1;
  double tiles = 0, cycles = 0; long long cmin = h[0], cmax = h[0];
  auto done_at = [&](int cta, long long t) {  // tiles finished by `cta` at clock t (linear inside a 64-tile block)
    const long long* p = hp + cta * 64;
    if (t <= p[0]) return 0.0;
    for (int b = 0; b < 63; ++b) if (t < p[b + 1]) return 64.0 * (b + (double)(t - p[b]) / (double)(p[b + 1] - p[b]));
    return 64.0 * 63;
  };
  for (int i = 0; i < 296; ++i) {
    cmin = std::min(cmin, h[4 * i]); cmax = std::max(cmax, h[4 * i]);
    const int sm = (int)h[4 * i + 3] & 255;
    if (partner[sm] < 0) { partner[sm] = i; continue; }
    const int a = partner[sm], b = i;
    const long long w0 = std::max(hp[a * 64], hp[b * 64]), w1 = std::min(hp[a * 64 + 63], hp[b * 64 + 63]);
    tiles += done_at(a, w1) - done_at(a, w0) + done_at(b, w1) - done_at(b, w0);
    cycles += (double)(w1 - w0);
  }
  printf("%-30s %7.1f cyc/tile/SM steady (%.2f el/clk/SM) | cta cycles %lld..%lld, %.3f ms, chk %.4g err=%d\n", name, cycles / tiles,
         16384.0 * tiles / cycles, cmin, cmax, ms, chk, (int)cudaGetLastError());
  cudaFree(d); cudaFree(c); cudaFree(pr);
}
int main() {
  run<1>("scalar, no poly");
  run<0>("scalar, 25% poly (r1 kernel)");
  run<2>("f32x2, no poly");
  run<7>("f32x2 + scalar 25% poly");
  run<3>("f32x2, 12.5% poly2");
  run<4>("f32x2, 25% poly2");
  run<5>("f32x2, 37.5% poly2");
  run<6>("f32x2, 50% poly2");
  return 0;
}
