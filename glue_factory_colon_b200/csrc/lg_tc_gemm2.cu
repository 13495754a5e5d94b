// bf16 linear layers, v2: weight-stationary tcgen05 GEMM with cluster multicast (sm_100a).
//
//   Y[T, N] = A[T, K] . W[N, K]^T  (+ fused epilogue),  K <= 512, N in {256, 512, 768}
//
// Why this shape.  With K <= 512 a 128-row output tile needs as many operand bytes as it has
// tensor-core cycles; streaming W per tile asks the L2 for 80-95 B/clk/SM, far more than it
// delivers (~35-40 B/clk/SM).  So every CTA keeps ONE column block of W resident in shared
// memory for its whole life (BN x K bf16 = 128 KB: BN = 256 at K <= 256, BN = 128 at K = 512) and
// only A tiles stream.  When N/BN > 1 column blocks exist at K = 512, the CL = N/BN CTAs that
// share an M tile form a thread-block cluster and each loads 1/CL of every A stage and TMA-
// multicasts it to its peers, which divides A traffic by CL as well (16-32 B/clk/SM in total).
//
// Roles (320 threads):  warp 0 TMA producer | warp 1 tcgen05.mma issuer + TMEM owner |
// warps 2-9 epilogue (two warps per TMEM lane quarter, each owning half of the columns).
// TMEM holds two BN-column accumulators: the epilogue copies a finished accumulator to
// registers, releases it at once, and does its math while the next tile's MMAs already run.
//
// Epilogues (compile-time):  ROW (bias, scale, optional fp32 residual, fp32 and/or bf16 out),
// HEADS (rotary + per-part scale, head-major bf16 out), LN (LayerNorm over the full 512-wide
// row + GELU; the row's columns live in CL = 4 CTAs, so per-row (sum, sum^2) partials are
// exchanged through distributed shared memory with cluster-scope mbarriers).
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_STAGE = BM * BK * 2;  // 16 KB
constexpr int NSTAGE = 4;
constexpr int W_BYTES = 128 * 1024;   // BN x K bf16
constexpr int OFF_A = W_BYTES;
constexpr int OFF_PAR = OFF_A + NSTAGE * A_STAGE;       // bias | gamma | beta, 3 x 256 floats
constexpr int OFF_STATS = OFF_PAR + 3 * 256 * 4;        // [2 buffers][8 slots][128 rows] float2
constexpr int OFF_BAR = OFF_STATS + 2 * 8 * 128 * 8;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

enum { MODE_ROW = 0, MODE_HEADS = 1, MODE_LN = 2 };

struct Args {
  int kb_total, kb_a0;
  int n_blocks;  // N / BN
  int m_tiles;
  const int32_t* lens;
};

// ---- cluster / DSMEM primitives ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAITC_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAITC_DONE;\n\t"
      "bra.uni WAITC_LOOP;\n\t"
      "WAITC_DONE:\n\t}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(tc::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// erf-GELU with the Abramowitz-Stegun 7.1.26 rational/exponential form (|err| < 2e-7, far
// below bf16 resolution): 2 MUFU + ~12 FMA instead of the ~30-instruction erff().
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * u * u));
  const float erf_abs = fmaf(-p, e, 1.f);
  const float h = 0.5f * x;
  return fmaf(copysignf(erf_abs, x), h, h);
}

__device__ __forceinline__ bool tile_skipped(const Args& g, int Lp, int m_tile) {
  if (!g.lens) return false;
  const int r0 = m_tile * BM;
  const int s = r0 / Lp;
  return r0 - s * Lp >= g.lens[s];
}

template <int MODE, int BN, int CL>
__global__ void __launch_bounds__(320, 1)
tc_ws_linear_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmW, Args g, LgEpi epi) {
  constexpr int TMEM_COLS = 2 * BN;             // 256 or 512
  constexpr int WB_BYTES = BN * BK * 2;         // one K block of the resident W
  constexpr int CPW = BN / 2;                   // columns per epilogue warp
  constexpr int NCH = CPW / 32;                 // 32-column chunks per epilogue warp
  constexpr int SLICE_ROWS = BM / CL;           // rows of an A stage this CTA loads (and multicasts)
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + OFF_A;
  float* s_par = reinterpret_cast<float*>(smem + OFF_PAR);
  float2* s_stats = reinterpret_cast<float2*>(smem + OFF_STATS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;   // [2]
  uint64_t* tempty = tfull + 2;       // [2]
  uint64_t* w_full = tempty + 2;
  uint64_t* stats_bar = w_full + 1;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stats_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CL > 1 ? cluster_rank() : 0;
  // column block of this CTA and its walk over M tiles
  int nb, m_first, m_step;
  if (CL > 1) {
    nb = (int)crank;
    m_first = blockIdx.x / CL;
    m_step = gridDim.x / CL;
  } else {
    nb = blockIdx.x % g.n_blocks;
    m_first = blockIdx.x / g.n_blocks;
    m_step = gridDim.x / g.n_blocks;
  }
  const int n0 = nb * BN;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA0);
    tc::prefetch_tmap(&tmA1);
    tc::prefetch_tmap(&tmW);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], CL); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 8);
      tc::mbar_init(&stats_bar[i], 8 * CL);
    }
    tc::mbar_init(w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < BN; i += blockDim.x) {
    s_par[i] = epi.bias[n0 + i];
    if (MODE == MODE_LN) {
      s_par[256 + i] = epi.gamma[n0 + i];
      s_par[512 + i] = epi.beta[n0 + i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(w_full, g.kb_total * WB_BYTES);
      for (int kb = 0; kb < g.kb_total; ++kb) tc::tma_load_2d(sW + kb * WB_BYTES, &tmW, w_full, kb * BK, n0);
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = m_first; mt < g.m_tiles; mt += m_step) {
        if (tile_skipped(g, epi.Lp, mt)) continue;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);  // all CL consumers released this stage
          tc::mbar_arrive_expect_tx(&full[stage], A_STAGE);
          uint8_t* dst = sA + stage * A_STAGE + crank * (SLICE_ROWS * 128);
          const CUtensorMap* tm = kb < g.kb_a0 ? &tmA0 : &tmA1;
          const int kc = (kb < g.kb_a0 ? kb : kb - g.kb_a0) * BK;
          const int row = mt * BM + crank * SLICE_ROWS;
          if (CL > 1) tma_load_2d_mc(dst, tm, &full[stage], kc, row, MC_MASK);
          else tc::tma_load_2d(dst, tm, &full[stage], kc, row);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_bf16(BM, BN, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      tc::mbar_wait(w_full, 0);
      const uint32_t aW = tc::smem_u32(sW);
      for (int mt = m_first; mt < g.m_tiles; mt += m_step) {
        if (tile_skipped(g, epi.Lp, mt)) continue;
        tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint32_t aA = tc::smem_u32(sA + stage * A_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_ss(d_tmem, tc::smem_desc_sw128(aA + k * 32, 0, 1024),
                        tc::smem_desc_sw128(aW + kb * WB_BYTES + k * 32, 0, 1024), idesc, (kb | k) != 0);
          if (CL > 1) umma_commit_mc(&empty[stage], MC_MASK);
          else tc::umma_commit(&empty[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;            // 0..7
    const int quarter = warp & 3;       // TMEM lane quarter accessible to this warp
    const int half = ew >> 2;           // which half of the BN columns
    const int r_in_tile = quarter * 32 + lane;
    const int c_warp = half * CPW;      // first column (within the block) of this warp
    int acc = 0, iter = 0;
    uint32_t acc_phase = 0;
    for (int mt = m_first; mt < g.m_tiles; mt += m_step) {
      if (tile_skipped(g, epi.Lp, mt)) continue;
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      uint32_t v[NCH][32];
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c_warp;
#pragma unroll
      for (int c = 0; c < NCH; ++c) tc::tmem_ld32(t_addr + c * 32, v[c]);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);  // accumulator is in registers: release it now
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      const int row = mt * BM + r_in_tile;
      if constexpr (MODE == MODE_ROW) {
        const float sc = epi.scale[0];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int cb = c_warp + c * 32;  // column within block
          const size_t off = (size_t)row * epi.N + n0 + cb;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = (__uint_as_float(v[c][j]) + s_par[cb + j]) * sc;
          if (epi.resid32) {
            const float4* rp = reinterpret_cast<const float4*>(epi.resid32 + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r4 = rp[j];
              f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
            }
          }
          if (epi.out32) {
            float4* op = reinterpret_cast<float4*>(epi.out32 + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
          if (epi.out16) {
            uint4* op = reinterpret_cast<uint4*>(epi.out16 + off);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              op[j] = make_uint4(tc::pack_bf16(f[8 * j], f[8 * j + 1]), tc::pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 tc::pack_bf16(f[8 * j + 4], f[8 * j + 5]), tc::pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
        }
      } else if constexpr (MODE == MODE_HEADS) {
        const int s = row / epi.Lp, l = row - s * epi.Lp;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int cb = c_warp + c * 32;
          const int col = n0 + cb;  // global output column: part*256 + head*64 + d
          const int part = col >> 8, h = (col >> 6) & 3, d0 = col & 63;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[c][j]) + s_par[cb + j];
          if (part < epi.n_rot) {
            const float4* rp = reinterpret_cast<const float4*>(epi.rot + (size_t)row * 64 + d0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 cs = rp[j];  // (cos, sin) of two rotary pairs
              const float a0 = f[4 * j] * cs.x - f[4 * j + 1] * cs.y, a1 = f[4 * j + 1] * cs.x + f[4 * j] * cs.y;
              const float a2 = f[4 * j + 2] * cs.z - f[4 * j + 3] * cs.w, a3 = f[4 * j + 3] * cs.z + f[4 * j + 2] * cs.w;
              f[4 * j] = a0; f[4 * j + 1] = a1; f[4 * j + 2] = a2; f[4 * j + 3] = a3;
            }
          }
          const float sc = part == 0 ? epi.scale[0] : (part == 1 ? epi.scale[1] : epi.scale[2]);
          void* basep = part == 0 ? epi.outp[0] : (part == 1 ? epi.outp[1] : epi.outp[2]);
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(basep) +
                                               (((size_t)s * LG_HEADS + h) * epi.Lp + l) * LG_DH + d0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            op[j] = make_uint4(tc::pack_bf16(f[8 * j] * sc, f[8 * j + 1] * sc), tc::pack_bf16(f[8 * j + 2] * sc, f[8 * j + 3] * sc),
                               tc::pack_bf16(f[8 * j + 4] * sc, f[8 * j + 5] * sc), tc::pack_bf16(f[8 * j + 6] * sc, f[8 * j + 7] * sc));
        }
      } else {  // MODE_LN
        // pass 1: partial row statistics over this warp's columns
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[c][j]) + s_par[c_warp + c * 32 + j];
            v[c][j] = __float_as_uint(x);
            sum += x;
            sq = fmaf(x, x, sq);
          }
        // publish to every CTA of the cluster (slot = rank*2 + half), then signal their barriers
        const int buf = iter & 1;
        const uint32_t slot_addr =
            tc::smem_u32(&s_stats[(buf * 8 + (int)crank * 2 + half) * 128 + r_in_tile]);
#pragma unroll
        for (int p = 0; p < CL; ++p) st_cluster_f2(CL > 1 ? map_to_rank(slot_addr, p) : slot_addr, sum, sq);
        fence_acq_rel_cluster();
        __syncwarp();
        if (lane == 0) {
          const uint32_t bar_addr = tc::smem_u32(&stats_bar[buf]);
#pragma unroll
          for (int p = 0; p < CL; ++p) mbar_arrive_remote(CL > 1 ? map_to_rank(bar_addr, p) : bar_addr);
        }
        mbar_wait_cluster(&stats_bar[buf], (iter >> 1) & 1);
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int p = 0; p < 2 * CL; ++p) {
          const float2 st = s_stats[(buf * 8 + p) * 128 + r_in_tile];
          ts += st.x;
          tq += st.y;
        }
        const float inv_n = 1.f / (float)(BN * CL);
        const float mean = ts * inv_n;
        const float rstd = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + 1e-5f);
        // pass 2: normalise + GELU on the register copy
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int cb = c_warp + c * 32;
          const size_t off = (size_t)row * epi.N + n0 + cb;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            f[j] = gelu_fast(fmaf((__uint_as_float(v[c][j]) - mean) * rstd, s_par[256 + cb + j], s_par[512 + cb + j]));
          if (epi.out16) {
            uint4* op = reinterpret_cast<uint4*>(epi.out16 + off);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              op[j] = make_uint4(tc::pack_bf16(f[8 * j], f[8 * j + 1]), tc::pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 tc::pack_bf16(f[8 * j + 4], f[8 * j + 5]), tc::pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
          if (epi.out32) {
            float4* op = reinterpret_cast<float4*>(epi.out32 + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      }
      ++iter;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody exits while peers may still multicast / arrive here
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int MODE, int BN, int CL>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, Args g, LgEpi epi, cudaStream_t st) {
  auto kern = tc_ws_linear_kernel<MODE, BN, CL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per = CL > 1 ? CL : g.n_blocks;   // CTAs that share an M tile
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(per);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int groups = sms / per;
  if (CL > 1) {
    // persistent clusters with static striding: never launch more clusters than can be co-resident
    static int max_clusters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (max_clusters[CL] == 0) {
      int n = 0;
      cfg.gridDim = dim3(sms / CL * CL);
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = sms / CL;
      max_clusters[CL] = n;
    }
    if (groups > max_clusters[CL]) groups = max_clusters[CL];
  }
  if (groups > g.m_tiles) groups = g.m_tiles;
  if (groups < 1) groups = 1;
  cfg.gridDim = dim3(groups * per);
  e = cudaLaunchKernelEx(&cfg, kern, a0, a1, w, g, epi);
  if (e != cudaSuccess) return (int)e;
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

}  // namespace

int lg_tc_linear_v2(int epilogue, const __nv_bfloat16* A0, const __nv_bfloat16* A1, int K0,
                    const __nv_bfloat16* W, int T, int N, int K, const int32_t* lens, LgEpi epi,
                    cudaStream_t st) {
  if (T % BM || K % BK || K0 % BK || K > 512 || (N != 256 && N != 512 && N != 768)) return LGB200_ERR_SHAPE;
  const bool ln = epilogue == LGB200_EPI_LN_GELU;
  if (ln && (N != 512 || K != 512)) return LGB200_ERR_SHAPE;
  if (K > 256 && N == 768) return LGB200_ERR_SHAPE;
  const int BN = K <= 256 ? 256 : 128;
  const int CL = K <= 256 ? 1 : N / BN;  // 2 (N = 256) or 4 (N = 512)
  CUtensorMap tA0, tA1, tW;
  const uint32_t boxA[2] = {BK, (uint32_t)(BM / CL)};
  {
    const uint64_t d[2] = {(uint64_t)K0, (uint64_t)T}, s[1] = {(uint64_t)K0 * 2};
    int rc = lg_make_tmap_bf16(&tA0, A0, 2, d, s, boxA);
    if (rc) return rc;
  }
  if (K0 < K) {
    const uint64_t d[2] = {(uint64_t)(K - K0), (uint64_t)T}, s[1] = {(uint64_t)(K - K0) * 2};
    int rc = lg_make_tmap_bf16(&tA1, A1, 2, d, s, boxA);
    if (rc) return rc;
  } else {
    tA1 = tA0;
  }
  {
    const uint64_t d[2] = {(uint64_t)K, (uint64_t)N}, s[1] = {(uint64_t)K * 2};
    const uint32_t b[2] = {BK, (uint32_t)BN};
    int rc = lg_make_tmap_bf16(&tW, W, 2, d, s, b);
    if (rc) return rc;
  }
  Args g;
  g.kb_total = K / BK;
  g.kb_a0 = K0 / BK;
  g.n_blocks = N / BN;
  g.m_tiles = T / BM;
  g.lens = lens;
  if (K <= 256) {
    if (epilogue == LGB200_EPI_ROWMAJOR) return launch<MODE_ROW, 256, 1>(tA0, tA1, tW, g, epi, st);
    if (epilogue == LGB200_EPI_HEADS) return launch<MODE_HEADS, 256, 1>(tA0, tA1, tW, g, epi, st);
    return LGB200_ERR_SHAPE;
  }
  if (ln) return launch<MODE_LN, 128, 4>(tA0, tA1, tW, g, epi, st);
  if (epilogue == LGB200_EPI_ROWMAJOR && N == 256) return launch<MODE_ROW, 128, 2>(tA0, tA1, tW, g, epi, st);
  return LGB200_ERR_SHAPE;
}
