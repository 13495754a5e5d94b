// fp32 CUDA-core kernels: the parity mode (LGB200_F32) of the LightGlue hot path.
// Same data layout, same epilogues and same entry points as the tcgen05 path;
// every contraction accumulates in fp32 in a fixed order, which is what lets
// log_assignment land within 1e-3 of the reference's fp32 run.
#include "lg_common.cuh"

// ---------------------------------------------------------------------------
// Tiled fp32 GEMM mainloop: acc[8][4] += A[128 rows, K] . B[64 rows, K]^T
// 256 threads; thread (ty, tx) owns rows ty*8..+8, cols tx*4..+4.
// A may be the column-wise concatenation of two matrices (A0: first K0 cols).
// ---------------------------------------------------------------------------
#define SG_BM 128
#define SG_BN 64
#define SG_BK 16

struct SimtSmem {
  float As[SG_BK][SG_BM + 4];
  float Bs[SG_BK][SG_BN + 4];
};

__device__ __forceinline__ void simt_mainloop(const float* __restrict__ A0, int lda0,
                                              const float* __restrict__ A1, int lda1, int K0,
                                              const float* __restrict__ Bm, int ldb, int K,
                                              SimtSmem& sm, float acc[8][4]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  // loaders: A tile 128x16 -> 512 float4, 2 per thread; B tile 64x16 -> 256 float4
  const int a_row0 = tid >> 2, a_k4 = (tid & 3) * 4;  // rows a_row0 and a_row0+64
  const int b_row = tid >> 2, b_k4 = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    const float* Ap;
    int lda, kk;
    if (k0 < K0) { Ap = A0; lda = lda0; kk = k0; } else { Ap = A1; lda = lda1; kk = k0 - K0; }
    const float4 a0 = *reinterpret_cast<const float4*>(Ap + (size_t)a_row0 * lda + kk + a_k4);
    const float4 a1 = *reinterpret_cast<const float4*>(Ap + (size_t)(a_row0 + 64) * lda + kk + a_k4);
    const float4 b0 = *reinterpret_cast<const float4*>(Bm + (size_t)b_row * ldb + k0 + b_k4);
    __syncthreads();
    sm.As[a_k4 + 0][a_row0] = a0.x; sm.As[a_k4 + 1][a_row0] = a0.y;
    sm.As[a_k4 + 2][a_row0] = a0.z; sm.As[a_k4 + 3][a_row0] = a0.w;
    sm.As[a_k4 + 0][a_row0 + 64] = a1.x; sm.As[a_k4 + 1][a_row0 + 64] = a1.y;
    sm.As[a_k4 + 2][a_row0 + 64] = a1.z; sm.As[a_k4 + 3][a_row0 + 64] = a1.w;
    sm.Bs[b_k4 + 0][b_row] = b0.x; sm.Bs[b_k4 + 1][b_row] = b0.y;
    sm.Bs[b_k4 + 2][b_row] = b0.z; sm.Bs[b_k4 + 3][b_row] = b0.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sm.As[k][ty * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sm.As[k][ty * 8 + 4]);
      const float4 y = *reinterpret_cast<const float4*>(&sm.Bs[k][tx * 4]);
      const float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float b[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------
// Linear layer
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simt_linear_kernel(const float* __restrict__ A0,
                                                          const float* __restrict__ A1, int K0,
                                                          const float* __restrict__ W, int K,
                                                          const int32_t* __restrict__ lens, LgEpi epi) {
  __shared__ SimtSmem sm;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  if (lens) {
    const int s = m0 / epi.Lp;
    if (m0 - s * epi.Lp >= lens[s]) return;
  }
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lda1 = K - K0;
  simt_mainloop(A0 + (size_t)m0 * K0, K0, A1 ? A1 + (size_t)m0 * lda1 : nullptr, lda1, K0,
                W + (size_t)n0 * K, K, K, sm, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 8; ++i) lg_epi_apply4<float>(epi, m0 + ty * 8 + i, n0 + tx * 4, acc[i]);
}

// LayerNorm(512) + GELU(erf), one warp per row, in place on fp32 [T,512].
__global__ void ln_gelu_kernel(float* __restrict__ h, int T, int Lp, const int32_t* __restrict__ lens,
                               const float* __restrict__ gamma, const float* __restrict__ beta) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= T) return;
  if (lens) {
    const int s = row / Lp;
    if (row - s * Lp >= ((lens[s] + 127) & ~127)) return;  // whole tile skipped by the GEMM
  }
  float4* p = reinterpret_cast<float4*>(h + (size_t)row * 512);
  float4 v[4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = p[lane + 32 * i];
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.f / 512.f);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += a * a + b * b + c * c + d * d;
  }
  for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.f / sqrtf(sq * (1.f / 512.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
    const float4 b = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
    float4 o;
    o.x = lg_gelu_erf((v[i].x - mean) * rstd * g.x + b.x);
    o.y = lg_gelu_erf((v[i].y - mean) * rstd * g.y + b.y);
    o.z = lg_gelu_erf((v[i].z - mean) * rstd * g.z + b.z);
    o.w = lg_gelu_erf((v[i].w - mean) * rstd * g.w + b.w);
    p[lane + 32 * i] = o;
  }
}

int lg_simt_linear(int epilogue, const float* A0, const float* A1, int K0, const float* W, int T, int N,
                   int K, const int32_t* lens, LgEpi epi, cudaStream_t st) {
  if (T % SG_BM || N % SG_BN || K % SG_BK || K0 % SG_BK) return LGB200_ERR_SHAPE;
  LgEpi e = epi;
  if (epilogue == LGB200_EPI_LN_GELU) {
    if (N != 512 || !epi.out32 || !epi.gamma || !epi.beta) return LGB200_ERR_SHAPE;
    e.mode = LGB200_EPI_ROWMAJOR;
    e.scale[0] = 1.f;
    e.resid32 = nullptr;
    e.out16 = nullptr;
  }
  dim3 grid(N / SG_BN, T / SG_BM);
  simt_linear_kernel<<<grid, 256, 0, st>>>(A0, A1, K0, W, K, lens, e);
  LG_LAUNCH_CHECK();
  if (epilogue == LGB200_EPI_LN_GELU) {
    ln_gelu_kernel<<<(T * 32 + 255) / 256, 256, 0, st>>>(epi.out32, T, epi.Lp, lens, epi.gamma, epi.beta);
    LG_LAUNCH_CHECK();
  }
  return LGB200_OK;
}

// ---------------------------------------------------------------------------
// fp32 flash attention.  CTA = 64 queries of one (sequence, head); 256 threads as
// a 16x16 grid, thread (ty,tx) owns rows ty*4..+4 and columns tx*4..+4 of both
// the 64x64 score tile and the 64x64 output tile.
// ---------------------------------------------------------------------------
#define FA_BM 64
#define FA_BN 64
struct FaSmem {
  float Qt[LG_DH][FA_BM + 4];  // [d][m]
  float Kt[LG_DH][FA_BN + 4];  // [d][n]
  float Vs[FA_BN][LG_DH + 4];  // [n][d]
  float Pt[FA_BN][FA_BM + 4];  // [n][m]
};

__global__ void __launch_bounds__(256) simt_attention_kernel(const float* __restrict__ Q,
                                                             const float* __restrict__ K,
                                                             const float* __restrict__ V, int Lp,
                                                             const int32_t* __restrict__ lens, int kv_xor,
                                                             float* __restrict__ ctx) {
  extern __shared__ __align__(16) unsigned char fa_raw[];
  FaSmem& sm = *reinterpret_cast<FaSmem*>(fa_raw);
  const int s = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BM;
  const int nq = lens ? lens[s] : Lp;
  if (q0 >= nq) return;
  const int skv = s ^ kv_xor;
  const int nk = lens ? lens[skv] : Lp;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float* Qh = Q + (((size_t)s * LG_HEADS + h) * Lp + q0) * LG_DH;
  const float* Kh = K + ((size_t)skv * LG_HEADS + h) * Lp * LG_DH;
  const float* Vh = V + ((size_t)skv * LG_HEADS + h) * Lp * LG_DH;
  // load Q tile transposed: 64x64 floats = 1024 float4, 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + 256 * i, m = idx >> 4, d4 = (idx & 15) * 4;
    const float4 v = *reinterpret_cast<const float4*>(Qh + (size_t)m * LG_DH + d4);
    sm.Qt[d4 + 0][m] = v.x; sm.Qt[d4 + 1][m] = v.y; sm.Qt[d4 + 2][m] = v.z; sm.Qt[d4 + 3][m] = v.w;
  }
  float o[4][4], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY;
    lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int n0 = 0; n0 < nk; n0 += FA_BN) {
    __syncthreads();  // previous tile fully consumed (also orders the Q stores on iteration 0)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i, n = idx >> 4, d4 = (idx & 15) * 4;
      const float4 kv = *reinterpret_cast<const float4*>(Kh + (size_t)(n0 + n) * LG_DH + d4);
      const float4 vv = *reinterpret_cast<const float4*>(Vh + (size_t)(n0 + n) * LG_DH + d4);
      sm.Kt[d4 + 0][n] = kv.x; sm.Kt[d4 + 1][n] = kv.y; sm.Kt[d4 + 2][n] = kv.z; sm.Kt[d4 + 3][n] = kv.w;
      *reinterpret_cast<float4*>(&sm.Vs[n][d4]) = vv;
    }
    __syncthreads();
    float sc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < LG_DH; ++d) {
      const float4 a = *reinterpret_cast<const float4*>(&sm.Qt[d][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sm.Kt[d][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(av[i], bv[j], sc[i][j]);
    }
    // mask + online softmax (scores are already in the log2 domain)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + tx * 4 + j >= nk) sc[i][j] = -INFINITY;
        mx = fmaxf(mx, sc[i][j]);
      }
      for (int ofs = 8; ofs; ofs >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, ofs));
      const float mnew = fmaxf(mrow[i], mx);
      const float msafe = mnew == -INFINITY ? 0.f : mnew;
      const float corr = exp2f(mrow[i] - msafe);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[i][j] = exp2f(sc[i][j] - msafe);
        rs += sc[i][j];
      }
      for (int ofs = 8; ofs; ofs >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, ofs);
      lrow[i] = lrow[i] * corr + rs;
      mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[i][j] *= corr;
        sm.Pt[tx * 4 + j][ty * 4 + i] = sc[i][j];
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < FA_BN; ++n) {
      const float4 a = *reinterpret_cast<const float4*>(&sm.Pt[n][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sm.Vs[n][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(av[i], bv[j], o[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float inv = lrow[i] > 0.f ? 1.f / lrow[i] : 0.f;
    const int l = q0 + ty * 4 + i;
    float4 r = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    *reinterpret_cast<float4*>(ctx + ((size_t)s * Lp + l) * LG_D + h * LG_DH + tx * 4) = r;
  }
}

int lg_simt_attention(const float* Q, const float* K, const float* V, int S, int Lp, const int32_t* lens,
                      int kv_xor, float* ctx, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(simt_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FaSmem));
  if (e != cudaSuccess) return (int)e;
  dim3 grid(Lp / FA_BM, LG_HEADS, S);
  simt_attention_kernel<<<grid, 256, sizeof(FaSmem), st>>>(Q, K, V, Lp, lens, kv_xor, ctx);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

// ---------------------------------------------------------------------------
// log assignment, fp32.  Both passes recompute the 128x64 similarity tile with
// the same mainloop, so the values that feed the normalisers are bit-identical
// to the values written out.
// ---------------------------------------------------------------------------
// pass 1: lse[s, l] = logsumexp_j <md[s,l], md[s^1,j]>   (natural log)
__global__ void __launch_bounds__(256) simt_assign_lse_kernel(const float* __restrict__ md, int Lp,
                                                              const int32_t* __restrict__ lens,
                                                              float* __restrict__ lse) {
  __shared__ SimtSmem sm;
  const int s = blockIdx.y, m0 = blockIdx.x * SG_BM;
  const int nq = lens ? lens[s] : Lp, nk = lens ? lens[s ^ 1] : Lp;
  if (m0 >= nq) return;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float* A = md + ((size_t)s * Lp + m0) * LG_D;
  const float* Bm = md + (size_t)(s ^ 1) * Lp * LG_D;
  float mrow[8], lrow[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mrow[i] = -INFINITY; lrow[i] = 0.f; }
  for (int n0 = 0; n0 < nk; n0 += SG_BN) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    simt_mainloop(A, LG_D, nullptr, 0, LG_D, Bm + (size_t)n0 * LG_D, LG_D, LG_D, sm, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + tx * 4 + j >= nk) acc[i][j] = -INFINITY;
        mx = fmaxf(mx, acc[i][j]);
      }
      for (int ofs = 8; ofs; ofs >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, ofs));
      const float mnew = fmaxf(mrow[i], mx);
      const float msafe = mnew == -INFINITY ? 0.f : mnew;
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) rs += expf(acc[i][j] - msafe);
      for (int ofs = 8; ofs; ofs >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, ofs);
      lrow[i] = lrow[i] * expf(mrow[i] - msafe) + rs;
      mrow[i] = mnew;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int l = m0 + ty * 8 + i;
      lse[(size_t)s * Lp + l] = (nk > 0) ? mrow[i] + logf(lrow[i]) : 0.f;
    }
  }
}

// pass 2: write scores [B,R,C].  Grid covers ceil(R/128) x ceil(C/64) tiles so the
// dustbin row/column and the zero padding are written by the same kernel.
__global__ void __launch_bounds__(256) simt_assign_scores_kernel(const float* __restrict__ md,
                                                                 const float* __restrict__ z,
                                                                 const float* __restrict__ lse, int Lp,
                                                                 const int32_t* __restrict__ lens, int R,
                                                                 int C, float* __restrict__ scores,
                                                                 float* __restrict__ sim_out) {
  // z == nullptr: nearest-neighbour matcher mode (no matchability terms, dustbins stay 0,
  // nearest_neighbor_matcher.py:73-74); sim_out (nullable): the raw similarity [B, R-1, C-1] (:64)
  __shared__ SimtSmem sm;
  const int b = blockIdx.z, m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int s0 = 2 * b, s1 = 2 * b + 1;
  const int na = lens ? lens[s0] : R - 1, nb = lens ? lens[s1] : C - 1;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool has_valid = (m0 < na) && (n0 < nb) && (m0 < Lp) && (n0 < Lp);
  if (has_valid)
    simt_mainloop(md + ((size_t)s0 * Lp + m0) * LG_D, LG_D, nullptr, 0, LG_D,
                  md + ((size_t)s1 * Lp + n0) * LG_D, LG_D, LG_D, sm, acc);
  float* out = scores + (size_t)b * R * C;
  float cz[4], cl[4], cneg[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = n0 + tx * 4 + j;
    cz[j] = cl[j] = cneg[j] = 0.f;
    if (c < nb) {
      if (z) {
        const float zz = z[(size_t)s1 * Lp + c];
        cz[j] = lg_logsigmoid(zz);
        cneg[j] = lg_logsigmoid(-zz);
      }
      cl[j] = lse[(size_t)s1 * Lp + c];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + ty * 8 + i;
    if (r >= R) continue;
    float rz = 0.f, rl = 0.f, rneg = 0.f;
    if (r < na) {
      if (z) {
        const float zz = z[(size_t)s0 * Lp + r];
        rz = lg_logsigmoid(zz);
        rneg = lg_logsigmoid(-zz);
      }
      rl = lse[(size_t)s0 * Lp + r];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= C) continue;
      float v = 0.f;
      if (r < na && c < nb) v = (acc[i][j] - rl) + (acc[i][j] - cl[j]) + (rz + cz[j]);
      else if (r < na && c == C - 1) v = rneg;
      else if (r == R - 1 && c < nb) v = cneg[j];
      out[(size_t)r * C + c] = v;
      if (sim_out && r < R - 1 && c < C - 1)
        sim_out[((size_t)b * (R - 1) + r) * (C - 1) + c] = (r < na && c < nb) ? acc[i][j] : 0.f;
    }
  }
}

int lg_simt_assign_lse(const float* md, int S, int Lp, const int32_t* lens, float* lse, cudaStream_t st) {
  dim3 grid(Lp / SG_BM, S);
  simt_assign_lse_kernel<<<grid, 256, 0, st>>>(md, Lp, lens, lse);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

int lg_simt_assign_scores(const float* md, const float* z, const float* lse, int B, int Lp,
                          const int32_t* lens, int R, int C, float* scores, float* sim_out, cudaStream_t st) {
  dim3 grid((C + SG_BN - 1) / SG_BN, (R + SG_BM - 1) / SG_BM, B);
  simt_assign_scores_kernel<<<grid, 256, 0, st>>>(md, z, lse, Lp, lens, R, C, scores, sim_out);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}
