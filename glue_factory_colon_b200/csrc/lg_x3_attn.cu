// fp32-accurate flash attention on tcgen05 (sm_100a): split-fp16 operands ("x3", see lg_x3.cu), head_dim 64.
//
//   S = Qh.Kh^T + Qh.Kl^T + Ql.Kh^T     (12 SS MMAs of 128x128x16 per key tile, fp32 accumulation in TMEM)
//   P = 2^(S - m) in fp32, split into two fp16 planes Ph + Pl, written to TMEM
//   O += Ph.Vh + Ph.Vl + Pl.Vh          (24 TS MMAs of 128x64x16, A operand from TMEM)
//
// One CTA = 128 queries of one (sequence, head), one CTA per SM (all 512 TMEM columns: two score buffers so that
// Q.K^T of tile j+1 runs during the softmax of tile j | Ph | Pl | O_tile).  warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2-9 = softmax (two threads per query row).  The softmax is the exact online one (row maximum before the
// exponentials on every tile, fp32 throughout); the reference moves only when the maximum grows by more than 4 (log2).
// TWO-LEVEL ACCUMULATION: the tensor core rounds its fp32 accumulator toward zero after every MMA (measured,
// tools/x3_micro.py: one accumulator over 2048 keys = 384 MMAs per row loses 5.8e-6 relative, one-sidedly, against
// 2.5e-8 for the CUDA-core kernel).  So P.V of ONE key tile (24 MMAs) goes into a fresh TMEM accumulator, and the
// softmax threads add the tile results to the running output in registers (round-to-nearest FADD), FA2-style; the
// rescaling on a moving maximum happens in those registers too -- no read-modify-write of O in TMEM.
// Replaces Attention.forward (lightglue.py:112-129) in the precision = fp32 mode; scores arrive in the log2 domain
// (the projection epilogue folds log2(e) / sqrt(64) into q).
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int TB = 128 * 64 * 2;                 // one plane of a 128 x 64 fp16 tile, 128-byte swizzle (16 KB)
constexpr int XA_KST = 2, XA_VST = 2;
constexpr int XA_Q = 0;                          // Qh | Ql
constexpr int XA_K = 2 * TB;                     // XA_KST stages of (Kh | Kl)
constexpr int XA_V = XA_K + XA_KST * 2 * TB;     // XA_VST stages of (Vh | Vl)
constexpr int XA_BAR = XA_V + XA_VST * 2 * TB;   // 160 KB of tiles
constexpr int XA_NBAR = 1 + 2 * XA_KST + 2 * XA_VST + 2 + 2 + 1 + 1;
constexpr int XA_XCH = XA_BAR + 256;             // [2 slots][128 rows][NP parts] fp32 exchange
constexpr int XA_SMEM = XA_XCH + 2 * 128 * 4 * 4;
constexpr uint32_t XT_S = 0, XT_PH = 256, XT_PL = 320, XT_O = 384, XT_OL = 448;  // O / OL: P.V of ONE key tile (large / small products)

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn_major) {
  return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// NP = softmax threads per query row: 2 (8 softmax warps, 64 key columns each) or 4 (16 warps, 32 columns each)
template <int NP>
__global__ void __launch_bounds__(64 + 128 * NP, 1)
x3_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, int Lp, const int32_t* __restrict__ lens, int kv_xor,
                    __half* __restrict__ ctx, size_t ctx_plane) {
  const int h = blockIdx.y, q0 = blockIdx.x * 128, s = blockIdx.z;
  const int nq = lens ? lens[s] : Lp;
  if (q0 >= nq) return;
  const int skv = s ^ kv_xor;
  const int nk = lens ? lens[skv] : Lp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (nk + 127) / 128;
  if (n_tiles == 0) {  // no keys: the attention output is zero (nan_to_num of the reference's masked rows)
    if (warp >= 2 && warp < 6) {
      const int r = (warp & 3) * 32 + lane;
      if (q0 + r < nq) {
        for (int pl = 0; pl < 2; ++pl) {
          uint4* dst = reinterpret_cast<uint4*>(ctx + pl * ctx_plane + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
        }
      }
    }
    return;
  }
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XA_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + XA_KST;
  uint64_t* v_full = k_empty + XA_KST;
  uint64_t* v_empty = v_full + XA_VST;
  uint64_t* s_full = v_empty + XA_VST;  // [2]
  uint64_t* s_free = s_full + 2;        // [2]
  uint64_t* p_ready = s_free + 2;
  uint64_t* pv_done = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + XA_NBAR);
  float* xch = reinterpret_cast<float*>(smem + XA_XCH);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < XA_KST; ++i) { tc::mbar_init(&k_full[i], 1); tc::mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < XA_VST; ++i) { tc::mbar_init(&v_full[i], 1); tc::mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_free[i], 4 * NP); }
    tc::mbar_init(p_ready, 4 * NP);
    tc::mbar_init(pv_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = (s * LG_HEADS + h) * Lp + q0;
      const int kvrow = (skv * LG_HEADS + h) * Lp;
      tc::mbar_arrive_expect_tx(q_full, 2 * TB);
      tc::tma_load_3d(smem + XA_Q, &tmQ, q_full, 0, qrow, 0);
      tc::tma_load_3d(smem + XA_Q + TB, &tmQ, q_full, 0, qrow, 1);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % XA_KST, vs = j % XA_VST;
        tc::mbar_wait(&k_empty[ks], ((j / XA_KST) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&k_full[ks], 2 * TB);
        tc::tma_load_3d(smem + XA_K + ks * 2 * TB, &tmK, &k_full[ks], 0, kvrow + j * 128, 0);
        tc::tma_load_3d(smem + XA_K + ks * 2 * TB + TB, &tmK, &k_full[ks], 0, kvrow + j * 128, 1);
        tc::mbar_wait(&v_empty[vs], ((j / XA_VST) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&v_full[vs], 2 * TB);
        tc::tma_load_3d(smem + XA_V + vs * 2 * TB, &tmV, &v_full[vs], 0, kvrow + j * 128, 0);
        tc::tma_load_3d(smem + XA_V + vs * 2 * TB + TB, &tmV, &v_full[vs], 0, kvrow + j * 128, 1);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = idesc_f16(128, 128, 0);
    constexpr uint32_t idesc_pv = idesc_f16(128, 64, 1);
    const uint64_t dQh = tc::smem_desc_sw128(tc::smem_u32(smem + XA_Q), 0, 1024);
    const uint64_t dQl = tc::smem_desc_sw128(tc::smem_u32(smem + XA_Q + TB), 0, 1024);
    auto issue_qk = [&](int j) {  // S[j & 1] = Q . K(j)^T
      const int ks = j % XA_KST;
      tc::mbar_wait(&k_full[ks], (j / XA_KST) & 1);
      tc::mbar_wait(&s_free[j & 1], ((j >> 1) & 1) ^ 1);  // the softmax has read tile j-2 out of this buffer
      tc::fence_after_sync();
      const uint64_t dKh = tc::smem_desc_sw128(tc::smem_u32(smem + XA_K + ks * 2 * TB), 0, 1024);
      const uint64_t dKl = tc::smem_desc_sw128(tc::smem_u32(smem + XA_K + ks * 2 * TB + TB), 0, 1024);
      const uint32_t tS = tmem + XT_S + (j & 1) * 128;
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc::umma_ss(tS, dQl + 2 * k, dKh + 2 * k, idesc_qk, k != 0);
          tc::umma_ss(tS, dQh + 2 * k, dKl + 2 * k, idesc_qk, 1);
          tc::umma_ss(tS, dQh + 2 * k, dKh + 2 * k, idesc_qk, 1);
        }
        tc::umma_commit(&s_full[j & 1]);
        tc::umma_commit(&k_empty[ks]);
      }
      __syncwarp();
    };
    tc::mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int vs = j % XA_VST;
      tc::mbar_wait(&v_full[vs], (j / XA_VST) & 1);
      tc::mbar_wait(p_ready, j & 1);
      tc::fence_after_sync();
      const uint64_t dVh = tc::smem_desc_sw128(tc::smem_u32(smem + XA_V + vs * 2 * TB), TB, 1024);
      const uint64_t dVl = tc::smem_desc_sw128(tc::smem_u32(smem + XA_V + vs * 2 * TB + TB), TB, 1024);
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 keys per MMA: P columns k*8.., V rows k*16.. (2048 B)
          // fresh accumulators per tile; the two small products have their own (one rounding per step on the large one)
          tc::umma_ts(tmem + XT_OL, tmem + XT_PL + k * 8, dVh + k * (2048 >> 4), idesc_pv, k != 0);
          tc::umma_ts(tmem + XT_OL, tmem + XT_PH + k * 8, dVl + k * (2048 >> 4), idesc_pv, 1);
          tc::umma_ts(tmem + XT_O, tmem + XT_PH + k * 8, dVh + k * (2048 >> 4), idesc_pv, k != 0);
        }
        tc::umma_commit(&v_empty[vs]);
        tc::umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    constexpr int COLS = 128 / NP;  // key columns of every tile owned by this thread
    constexpr int OC = 64 / NP;     // output columns owned by this thread
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    float m_ref = -INFINITY, l_part = 0.f;
    float oacc[OC];  // running output of this thread's columns (sum over the tiles before the previous one)
#pragma unroll
    for (int i = 0; i < OC; ++i) oacc[i] = 0.f;
    float* my = xch + r * NP;  // [slot][row][part]
    auto row_max = [&](int slot) -> float {
      if constexpr (NP == 2) {
        const float2 v = *reinterpret_cast<const float2*>(my + slot * (128 * NP));
        return fmaxf(v.x, v.y);
      } else {
        const float4 v = *reinterpret_cast<const float4*>(my + slot * (128 * NP));
        return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      }
    };
    auto ld_o = [&](uint32_t col, uint32_t* dst) {
      if constexpr (OC == 32) tc::tmem_ld32(tmem + lane_base + col + part * OC, dst);
      else tc::tmem_ld16(tmem + lane_base + col + part * OC, dst);
    };
    for (int j = 0; j < n_tiles; ++j) {
      tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc::fence_after_sync();
      uint32_t sv[COLS];
#pragma unroll
      for (int c = 0; c < COLS; c += 32) tc::tmem_ld32(tmem + lane_base + XT_S + (j & 1) * 128 + part * COLS + c, sv + c);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_free[j & 1]);
      const int valid = nk - j * 128 - part * COLS;
      if (valid < COLS) {
#pragma unroll
        for (int i = 0; i < COLS; ++i)
          if (i >= valid) sv[i] = 0xff800000u;  // -inf
      }
      constexpr float s_us = 1.f / (LG_X3_EA * LG_X3_EA);  // S arrives scaled by the plane scaling of q and k
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < COLS; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * s_us;
      my[(j & 1) * (128 * NP) + part] = mx;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NP) : "memory");
      mx = row_max(j & 1);  // the row's maximum over the whole tile (tile 0 holds at least one valid key)
      // lazy rescale: the reference moves only when the maximum grew by more than 4 (P stays below 2^4)
      float alpha = 1.f;
      const bool move = mx > m_ref + 4.f || m_ref == -INFINITY;  // P <= 2^4, P * LG_X3_EP stays inside fp16
      if (move) {
        alpha = m_ref == -INFINITY ? 0.f : ex2f(m_ref - mx);
        m_ref = mx;
      }
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t ph[COLS / 2], pl[COLS / 2];
#pragma unroll
      for (int i = 0; i < COLS / 2; ++i) {
        const float p0 = ex2f(fmaf(__uint_as_float(sv[2 * i]), s_us, -m_ref)), p1 = ex2f(fmaf(__uint_as_float(sv[2 * i + 1]), s_us, -m_ref));
        rs4[i & 3] += p0 + p1;
        const float e0 = p0 * LG_X3_EP, e1 = p1 * LG_X3_EP;
        const __half2 hh = __floats2half2_rn(e0, e1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(e0 - hf.x, e1 - hf.y);
        ph[i] = *reinterpret_cast<const uint32_t*>(&hh);
        pl[i] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      l_part = l_part * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      if (j > 0) {
        tc::mbar_wait(pv_done, (j - 1) & 1);  // P.V(j-1) retired: P is free, its tile result is complete
        tc::fence_after_sync();
        uint32_t o[OC], ol[OC];
        ld_o(XT_O, o);
        ld_o(XT_OL, ol);
        tc::tmem_ld_wait();
        // tile j-1 was formed against the reference that held before this tile's move: add, then rescale
#pragma unroll
        for (int i = 0; i < OC; ++i) oacc[i] = (oacc[i] + (__uint_as_float(o[i]) + __uint_as_float(ol[i]))) * alpha;
      }
      if constexpr (NP == 2) {
        tc::tmem_st32(tmem + lane_base + XT_PH + part * (COLS / 2), ph);
        tc::tmem_st32(tmem + lane_base + XT_PL + part * (COLS / 2), pl);
      } else {
        tc::tmem_st16(tmem + lane_base + XT_PH + part * (COLS / 2), ph);
        tc::tmem_st16(tmem + lane_base + XT_PL + part * (COLS / 2), pl);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_ready);
    }
    // row sum over the NP parts, normalise, store the two output planes
    my[(n_tiles & 1) * (128 * NP) + part] = l_part;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NP) : "memory");
    float l_sum;
    if constexpr (NP == 2) {
      const float2 lv = *reinterpret_cast<const float2*>(my + (n_tiles & 1) * (128 * NP));
      l_sum = lv.x + lv.y;
    } else {
      const float4 lv = *reinterpret_cast<const float4*>(my + (n_tiles & 1) * (128 * NP));
      l_sum = (lv.x + lv.y) + (lv.z + lv.w);
    }
    tc::mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc::fence_after_sync();
    // O carries the plane scaling of P and V; the output planes carry the activation scaling again
    const float inv = l_sum > 0.f ? (1.f / l_sum) * (1.f / LG_X3_EP) : 0.f;
    {
      uint32_t o[OC], ol[OC];
      ld_o(XT_O, o);
      ld_o(XT_OL, ol);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < OC; ++i) oacc[i] += __uint_as_float(o[i]) + __uint_as_float(ol[i]);
    }
    if (q0 + r < nq) {
      uint32_t oh[OC / 2], ol[OC / 2];
#pragma unroll
      for (int i = 0; i < OC / 2; ++i) {
        const float a = oacc[2 * i] * inv, b = oacc[2 * i + 1] * inv;
        const __half2 hh = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
        oh[i] = *reinterpret_cast<const uint32_t*>(&hh);
        ol[i] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      __half* dst = ctx + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH + part * OC;
      uint4* dh = reinterpret_cast<uint4*>(dst);
      uint4* dl = reinterpret_cast<uint4*>(dst + ctx_plane);
#pragma unroll
      for (int i = 0; i < OC / 8; ++i) {
        dh[i] = make_uint4(oh[4 * i], oh[4 * i + 1], oh[4 * i + 2], oh[4 * i + 3]);
        dl[i] = make_uint4(ol[4 * i], ol[4 * i + 1], ol[4 * i + 2], ol[4 * i + 3]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// Q, K, V: split planes [2][S*4*Lp][64] fp16 (head-major, as written by the X_HEADS epilogue); ctx: [2][S*Lp][256]
int lg_x3_attention(const void* Q, const void* K, const void* V, int S, int Lp, const int32_t* lens, int kv_xor,
                    void* ctx, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  const uint64_t rows = (uint64_t)S * LG_HEADS * Lp;
  const uint64_t d[3] = {64, rows, 2}, sb[2] = {128, rows * 128};
  const uint32_t box[3] = {64, 128, 1};
  int rc;
  if ((rc = lg_make_tmap_bf16(&tq, Q, 3, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tk, K, 3, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tv, V, 3, d, sb, box))) return rc;
  // 16 softmax warps (NP = 4) measured equal to 8 (1031 vs 1037 pairs/s at 64 x 2048): the kernel is bound by its 36 MMAs per tile
  static const int np = getenv("LGB200_X3_ATTN_NP") ? atoi(getenv("LGB200_X3_ATTN_NP")) : 2;
  cudaError_t e;
  if (np == 2) {
    if ((e = cudaFuncSetAttribute(x3_attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM)) != cudaSuccess) return (int)e;
    x3_attention_kernel<2><<<dim3(Lp / 128, LG_HEADS, S), 320, XA_SMEM, st>>>(tq, tk, tv, Lp, lens, kv_xor, (__half*)ctx,
                                                                               (size_t)S * Lp * LG_D);
  } else {
    if ((e = cudaFuncSetAttribute(x3_attention_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM)) != cudaSuccess) return (int)e;
    x3_attention_kernel<4><<<dim3(Lp / 128, LG_HEADS, S), 576, XA_SMEM, st>>>(tq, tk, tv, Lp, lens, kv_xor, (__half*)ctx,
                                                                               (size_t)S * Lp * LG_D);
  }
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}
