// bf16 linear layers on the 5th-gen tensor cores (tcgen05), sm_100a.
//
//   Y[T,N] = A[T,K] . W[N,K]^T  (+ fused epilogue)
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   A/W tiles -> 128B-swizzled shared memory ring (mbarrier full/empty)
//   warp 1      MMA issuer     one thread issues tcgen05.mma (M=128, N<=256, K=16), accumulators in TMEM;
//                              tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2-5   epilogue       tcgen05.ld (thread = row), bias/scale/rotary/residual or LayerNorm+GELU,
//                              global stores; TMEM accumulators are double buffered so the epilogue of
//                              tile i overlaps the MMAs of tile i+1
// A can be the column-wise concatenation of two tensors (two tensor maps) -- the FFN's cat([x, msg]).
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"

#include <mutex>

// ----------------------------------------------------------------------------- tensor map encode
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled lg_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int lg_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
  return lg_make_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, 128);
}

int lg_make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled enc = lg_get_encode();
  if (!enc) return LGB200_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                   bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 0    ? CU_TENSOR_MAP_SWIZZLE_NONE
                   : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                         : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LGB200_OK : LGB200_ERR_DRIVER;
}

// ----------------------------------------------------------------------------- kernel
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 128 ? 6 : (BN == 256 ? 4 : 2);
  static constexpr int ACC = BN == 512 ? 1 : 2;
  static constexpr int TMEM_COLS = BN * ACC;  // 256 or 512 (power of two)
  static constexpr int UMMA_N = BN > 256 ? 256 : BN;
  static constexpr int N_SPLIT = BN / UMMA_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 3 * 512 * 4;
};

struct GemmArgs {
  int kb_total;   // K / 64
  int kb_a0;      // K0 / 64
  int n_tiles;    // N / BN
  int m_tiles;    // T / 128
  const int32_t* lens;
};

__device__ __forceinline__ bool tile_skipped(const GemmArgs& g, const LgEpi& e, int m_tile) {
  if (!g.lens) return false;
  const int r0 = m_tile * BM;
  const int s = r0 / e.Lp;
  return r0 - s * e.Lp >= g.lens[s];
}

template <int BN, bool LN>
__global__ void __launch_bounds__(192, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, GemmArgs g, LgEpi epi) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + C::ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC);
  float* s_par = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);  // bias | gamma | beta

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA0);
    tc::prefetch_tmap(&tmA1);
    tc::prefetch_tmap(&tmW);
    for (int i = 0; i < C::STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < C::ACC; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
  if (LN) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      s_par[i] = epi.bias[i];
      s_par[512 + i] = epi.gamma[i];
      s_par[1024 + i] = epi.beta[i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m_tile = t / g.n_tiles, n_tile = t - m_tile * g.n_tiles;
        if (tile_skipped(g, epi, m_tile)) continue;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          tc::mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
          if (kb < g.kb_a0) tc::tma_load_2d(sa, &tmA0, &full[stage], kb * BK, m_tile * BM);
          else tc::tma_load_2d(sa, &tmA1, &full[stage], (kb - g.kb_a0) * BK, m_tile * BM);
#pragma unroll
          for (int h = 0; h < C::N_SPLIT; ++h)
            tc::tma_load_2d(sb + h * C::UMMA_N * BK * 2, &tmW, &full[stage], kb * BK, n_tile * BN + h * C::UMMA_N);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_bf16(BM, C::UMMA_N, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m_tile = t / g.n_tiles;
        if (tile_skipped(g, epi, m_tile)) continue;
        tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = tc::smem_desc_sw128(sa + k * 32, 0, 1024);
#pragma unroll
            for (int h = 0; h < C::N_SPLIT; ++h) {
              const uint64_t bdesc = tc::smem_desc_sw128(sb + h * C::UMMA_N * BK * 2 + k * 32, 0, 1024);
              tc::umma_ss(d_tmem + h * C::UMMA_N, adesc, bdesc, idesc, (kb | k) != 0);
            }
          }
          tc::umma_commit(&empty[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tfull[acc]);      // accumulator complete
        if (++acc == C::ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m_tile = t / g.n_tiles, n_tile = t - m_tile * g.n_tiles;
      if (tile_skipped(g, epi, m_tile)) continue;
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      const int row = m_tile * BM + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      if constexpr (!LN) {
        const int n0 = n_tile * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float v[4] = {__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                          __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])};
            lg_epi_apply4<__nv_bfloat16>(epi, row, n0 + c * 32 + 4 * q, v);
          }
        }
      } else {
        // pass 1: row statistics over all 512 columns (thread-local: one thread = one row)
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]) + s_par[c * 32 + j];
            sum += v;
            sq = fmaf(v, v, sq);
          }
        }
        const float mean = sum * (1.f / 512.f);
        const float var = fmaxf(sq * (1.f / 512.f) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        // pass 2: normalise, GELU, store bf16 (and/or fp32)
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int col = c * 32 + j;
            float a = (__uint_as_float(r[j]) + s_par[col] - mean) * rstd * s_par[512 + col] + s_par[1024 + col];
            float b = (__uint_as_float(r[j + 1]) + s_par[col + 1] - mean) * rstd * s_par[512 + col + 1] +
                      s_par[1024 + col + 1];
            a = lg_gelu_erf(a);
            b = lg_gelu_erf(b);
            pk[j >> 1] = tc::pack_bf16(a, b);
            if (epi.out32) {
              epi.out32[(size_t)row * 512 + col] = a;
              epi.out32[(size_t)row * 512 + col + 1] = b;
            }
          }
          if (epi.out16) {
            uint4* dst = reinterpret_cast<uint4*>(epi.out16 + (size_t)row * 512 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == C::ACC) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, bool LN>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, GemmArgs g, LgEpi epi,
           cudaStream_t st) {
  using C = Cfg<BN>;
  auto kern = tc_linear_kernel<BN, LN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = g.m_tiles * g.n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, 192, C::SMEM_BYTES, st>>>(a0, a1, w, g, epi);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

}  // namespace

int lg_tc_linear(int epilogue, const __nv_bfloat16* A0, const __nv_bfloat16* A1, int K0,
                 const __nv_bfloat16* W, int T, int N, int K, const int32_t* lens, LgEpi epi,
                 cudaStream_t st) {
  if (T % BM || K % BK || K0 % BK || (N != 256 && N != 512 && N != 768)) return LGB200_ERR_SHAPE;
  const bool ln = epilogue == LGB200_EPI_LN_GELU;
  if (ln && !epi.out16 && !epi.out32) return LGB200_ERR_NULL;
  const int BN = ln ? 512 : 256;
  CUtensorMap tA0, tA1, tW;
  {
    const uint64_t d[2] = {(uint64_t)K0, (uint64_t)T}, s[1] = {(uint64_t)K0 * 2};
    const uint32_t b[2] = {BK, BM};
    int rc = lg_make_tmap_bf16(&tA0, A0, 2, d, s, b);
    if (rc) return rc;
  }
  if (K0 < K) {
    const uint64_t d[2] = {(uint64_t)(K - K0), (uint64_t)T}, s[1] = {(uint64_t)(K - K0) * 2};
    const uint32_t b[2] = {BK, BM};
    int rc = lg_make_tmap_bf16(&tA1, A1, 2, d, s, b);
    if (rc) return rc;
  } else {
    tA1 = tA0;
  }
  {
    const uint64_t d[2] = {(uint64_t)K, (uint64_t)N}, s[1] = {(uint64_t)K * 2};
    const uint32_t b[2] = {BK, 256};
    int rc = lg_make_tmap_bf16(&tW, W, 2, d, s, b);
    if (rc) return rc;
  }
  GemmArgs g;
  g.kb_total = K / BK;
  g.kb_a0 = K0 / BK;
  g.n_tiles = N / BN;
  g.m_tiles = T / BM;
  g.lens = lens;
  if (ln) return launch<512, true>(tA0, tA1, tW, g, epi, st);
  return launch<256, false>(tA0, tA1, tW, g, epi, st);
}
