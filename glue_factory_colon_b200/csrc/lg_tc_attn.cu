// Flash attention forward on tcgen05 / TMEM / TMA (sm_100a), head_dim 64, bf16 in, fp32 accumulate.
//
// One CTA = 128 query rows of one (sequence, head).  Two CTAs are resident per SM (80 KB smem,
// 256 TMEM columns each) so that one CTA's softmax overlaps the other's MMAs.
//   warp 0     TMA producer: Q once, then K/V tiles of 128 keys into a 2-stage ring
//   warp 1     MMA issuer:   S = Q.K^T (SS, N=128) into TMEM, O += P.V (TS: P read from TMEM, V MN-major)
//   warps 2-5  softmax:      thread = query row.  tcgen05.ld S -> registers, online max with lazy
//                            rescale (O is only touched when the max grows by > 8 in log2 units),
//                            ex2, row sum, bf16 P written back to TMEM with tcgen05.st
// Issue order QK(j+1) before PV(j): the next score tile is produced while the softmax warps are
// still exponentiating tile j, and PV(j) runs while they work on tile j+1.
// TMEM columns: [0,128) S fp32 | [128,192) P bf16x2 | [192,256) O fp32.
// Scores arrive in the log2 domain (the Q projection epilogue folds log2(e)/sqrt(64)).
// Keys >= lens[kv sequence] are masked to -inf; query rows >= lens[q sequence] are not stored.
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"

namespace {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per step
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int AT_SMEM = TILE_BYTES * 5 + 1024 + 128;

constexpr uint32_t TM_S = 0, TM_P = 128, TM_O = 192, TM_COLS = 256;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(192, 2)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, int Lp, const int32_t* __restrict__ lens,
                    int kv_xor, __nv_bfloat16* __restrict__ ctx) {
  const int s = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_BM;
  const int nq = lens ? lens[s] : Lp;
  if (q0 >= nq) return;
  const int skv = s ^ kv_xor;
  const int nk = lens ? lens[skv] : Lp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (nk + AT_BN - 1) / AT_BN;

  if (n_tiles == 0) {  // no keys: attention output is defined as zero (nan_to_num)
    if (warp >= 2) {
      const int r = (warp & 3) * 32 + lane;
      if (q0 + r < nq) {
        uint4* dst = reinterpret_cast<uint4*>(ctx + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;       // 2 stages
  uint8_t* sV = smem + 3 * TILE_BYTES;   // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&kv_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(s_free, 4);
    tc::mbar_init(p_ready, 4);
    tc::mbar_init(pv_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = (s * LG_HEADS + h) * Lp + q0;
      const int kvrow = (skv * LG_HEADS + h) * Lp;
      tc::mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tc::tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        tc::mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[st], 2 * TILE_BYTES);
        tc::tma_load_2d(sK + st * TILE_BYTES, &tmK, &kv_full[st], 0, kvrow + j * AT_BN);
        tc::tma_load_2d(sV + st * TILE_BYTES, &tmV, &kv_full[st], 0, kvrow + j * AT_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = tc::idesc_bf16(128, 128, 0);
      constexpr uint32_t idesc_pv = tc::idesc_bf16(128, 64, 1);
      const uint32_t aQ = tc::smem_u32(sQ);
      auto issue_qk = [&](int j) {
        const uint32_t aK = tc::smem_u32(sK + (j & 1) * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_ss(tmem + TM_S, tc::smem_desc_sw128(aQ + k * 32, 0, 1024),
                      tc::smem_desc_sw128(aK + k * 32, 0, 1024), idesc_qk, k != 0);
        tc::umma_commit(s_full);
      };
      tc::mbar_wait(q_full, 0);
      tc::mbar_wait(&kv_full[0], 0);
      tc::fence_after_sync();
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          tc::mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc::mbar_wait(s_free, j & 1);  // softmax holds S(j) in registers
          tc::fence_after_sync();
          issue_qk(j + 1);
        }
        tc::mbar_wait(p_ready, j & 1);
        tc::fence_after_sync();
        const uint32_t aV = tc::smem_u32(sV + (j & 1) * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)  // 16 keys per MMA: P columns k*8.., V rows k*16..
          tc::umma_ts(tmem + TM_O, tmem + TM_P + k * 8, tc::smem_desc_sw128(aV + k * 2048, TILE_BYTES, 1024),
                      idesc_pv, (j | k) != 0);
        tc::umma_commit(&kv_empty[j & 1]);
        tc::umma_commit(pv_done);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      tc::mbar_wait(s_full, j & 1);
      tc::fence_after_sync();
      uint32_t sv[128];
      tc::tmem_ld32(tmem + lane_base + TM_S + 0, sv + 0);
      tc::tmem_ld32(tmem + lane_base + TM_S + 32, sv + 32);
      tc::tmem_ld32(tmem + lane_base + TM_S + 64, sv + 64);
      tc::tmem_ld32(tmem + lane_base + TM_S + 96, sv + 96);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(s_free);
      const int valid = nk - j * AT_BN;  // >= 1
      float mx = -INFINITY;
      if (valid < AT_BN) {
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          if (i >= valid) sv[i] = 0xff800000u;  // -inf
        }
      }
#pragma unroll
      for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      // lazy rescale: keep the reference max unless it grows by more than 8 (factor 256)
      float m_new = m_ref;
      if (mx > m_ref + 8.f) m_new = mx;
      const float alpha = ex2(m_ref - m_new);  // 1 when unchanged, 0 on the first tile
      float rs = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float p0 = ex2(__uint_as_float(sv[2 * i]) - m_new);
        const float p1 = ex2(__uint_as_float(sv[2 * i + 1]) - m_new);
        rs += p0 + p1;
        pk[i] = tc::pack_bf16(p0, p1);
      }
      l_sum = l_sum * alpha + rs;
      if (j > 0) {
        tc::mbar_wait(pv_done, (j - 1) & 1);  // PV(j-1) retired: P is free, O is up to date
        tc::fence_after_sync();
        const bool need = m_new != m_ref;
        if (__any_sync(0xffffffffu, need)) {
          uint32_t o[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tc::tmem_ld32(tmem + lane_base + TM_O + half * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st32(tmem + lane_base + TM_O + half * 32, o);
          }
        }
      }
      m_ref = m_new;
      tc::tmem_st32(tmem + lane_base + TM_P + 0, pk + 0);
      tc::tmem_st32(tmem + lane_base + TM_P + 32, pk + 32);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_ready);
    }
    tc::mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc::fence_after_sync();
    const float inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
    uint32_t o[64];
    tc::tmem_ld32(tmem + lane_base + TM_O + 0, o);
    tc::tmem_ld32(tmem + lane_base + TM_O + 32, o + 32);
    tc::tmem_ld_wait();
    if (q0 + r < nq) {
      uint4* dst = reinterpret_cast<uint4*>(ctx + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 w;
        w.x = tc::pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        w.y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        w.z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        w.w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        dst[i] = w;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
}

}  // namespace

int lg_tc_attention(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                    const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  const uint64_t d[2] = {64, (uint64_t)S * LG_HEADS * Lp}, sb[1] = {128};
  const uint32_t box[2] = {64, 128};
  int rc;
  if ((rc = lg_make_tmap_bf16(&tq, Q, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tk, K, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tv, V, 2, d, sb, box))) return rc;
  cudaError_t e = cudaFuncSetAttribute(tc_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(Lp / AT_BM, LG_HEADS, S);
  tc_attention_kernel<<<grid, 192, AT_SMEM, st>>>(tq, tk, tv, Lp, lens, kv_xor, ctx);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}
