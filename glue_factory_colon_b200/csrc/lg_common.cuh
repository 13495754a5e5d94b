// Shared helpers for the LightGlue B200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/lightglue_b200.h"

#define LG_HEADS 4
#define LG_DH 64
#define LG_D 256

// Every kernel launch is followed by this macro: it surfaces launch errors and bumps the
// process-wide launch counter that bench.py reports as "gpu_launches" (statistics only).
extern unsigned long long lg_launch_counter;
#define LG_LAUNCH_CHECK()                       \
  do {                                          \
    cudaError_t e__ = cudaGetLastError();       \
    if (e__ != cudaSuccess) return (int)e__;    \
    ++lg_launch_counter;                        \
  } while (0)

static inline cudaStream_t lg_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float lg_logsigmoid(float x) {
  // log(sigmoid(x)) = min(x,0) - log1p(exp(-|x|))   (same form as ATen's log_sigmoid)
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float lg_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float lg_gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

// filter_matches: (value, index) packed so that a 64-bit max picks the largest value and, among equal
// values, the LOWEST index (torch.max tie rule); NaN sorts above everything (torch.max propagates NaN).
// Known deviation: -0.0 sorts below +0.0 here, torch.max treats them as equal (tests/test_filter_keys.py).
__device__ __forceinline__ unsigned long long fm_pack(float v, int idx) {
  const unsigned u = __float_as_uint(v);
  unsigned key;
  if (v != v) key = 0xffffffffu;
  else key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)key << 32) | (unsigned)(0xffffffffu - (unsigned)idx);
}
__device__ __forceinline__ float fm_value(unsigned long long p) {
  const unsigned key = (unsigned)(p >> 32);
  if (key == 0xffffffffu) return __uint_as_float(0x7fc00000u);
  const unsigned u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  return __uint_as_float(u);
}
__device__ __forceinline__ int fm_index(unsigned long long p) {
  return (int)(0xffffffffu - (unsigned)(p & 0xffffffffu));
}

// Epilogue description shared by the fp32 (CUDA-core) and bf16 (tcgen05) GEMMs.
struct LgEpi {
  int mode;  // LGB200_EPI_*
  int N;     // output columns
  int Lp;
  const float* bias;
  float scale[3];
  const float* resid32;
  float* out32;
  __nv_bfloat16* out16;
  const float* rot;  // [T,64] (cos,sin) pairs
  int n_rot;
  void* outp[3];  // head-major part outputs
  const float* gamma;
  const float* beta;
};

// Apply the ROWMAJOR / HEADS epilogue to 4 consecutive columns [c, c+4) of row r.
// `v` holds raw accumulators.  OutT is the element type of the head-major buffers.
template <typename OutT>
__device__ __forceinline__ void lg_epi_apply4(const LgEpi& e, int r, int c, float v[4]) {
  const float4 b4 = *reinterpret_cast<const float4*>(e.bias + c);
  v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
  if (e.mode == LGB200_EPI_ROWMAJOR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] *= e.scale[0];
    const size_t off = (size_t)r * e.N + c;
    if (e.resid32) {
      const float4 r4 = *reinterpret_cast<const float4*>(e.resid32 + off);
      v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
    }
    if (e.out32) *reinterpret_cast<float4*>(e.out32 + off) = make_float4(v[0], v[1], v[2], v[3]);
    if (e.out16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(e.out16 + off) = pk;
    }
  } else {  // LGB200_EPI_HEADS
    const int part = c >> 8, h = (c >> 6) & 3, d = c & 63;
    if (part < e.n_rot) {
      // rotary pairs (d, d+1), (d+2, d+3) use frequencies d/2, d/2+1
      const float4 cs = *reinterpret_cast<const float4*>(e.rot + (size_t)r * 64 + d);
      const float a0 = v[0] * cs.x - v[1] * cs.y, a1 = v[1] * cs.x + v[0] * cs.y;
      const float a2 = v[2] * cs.z - v[3] * cs.w, a3 = v[3] * cs.z + v[2] * cs.w;
      v[0] = a0; v[1] = a1; v[2] = a2; v[3] = a3;
    }
    const float sc = e.scale[part];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] *= sc;
    const int s = r / e.Lp, l = r - s * e.Lp;
    const size_t off = (((size_t)s * LG_HEADS + h) * e.Lp + l) * LG_DH + d;
    OutT* base = reinterpret_cast<OutT*>(e.outp[part]);
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) =
          make_float4(v[0], v[1], v[2], v[3]);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = pk;
    }
  }
}
