// bf16 linear layers, v2: weight-stationary tcgen05 GEMM with cluster multicast (sm_100a).
//
//   Y[T, N] = A[T, K] . W[N, K]^T  (+ fused epilogue),  K <= 512, N in {256, 512, 768}
//
// Why this shape.  With K <= 512 a 128-row output tile needs as many operand bytes as it has
// tensor-core cycles; streaming W per tile would ask the L2 for 80-95 B/clk/SM, far more than it
// delivers.  So every CTA keeps ONE 128-column block of W resident in shared memory for its whole
// life (128 x K bf16 = 64 KB at K = 256, 128 KB at K = 512) and only A tiles stream.  The CL CTAs
// that work on the same M tile form a thread-block cluster; each loads 1/CL of every A stage and
// TMA-multicasts it to its peers (A traffic / CL).
//
// Roles (320 threads):  warp 0 TMA producer | warp 1 tcgen05.mma issuer + TMEM owner |
// warps 2-9 epilogue: warp (q, h) owns TMEM lane quarter q (32 rows) and column half h (64 cols).
// TMEM holds two 128-column accumulators: the epilogue copies a finished accumulator to registers,
// releases it at once, and does its math while the next tile's MMAs already run.
//
// Epilogue I/O never touches global memory from a thread-per-row pattern (that was measured at
// 32 sectors/request and 70 % L1 utilisation in v1).  Each epilogue warp owns 4 KB staging tiles
// (32 rows x 128 B, 128-byte swizzle): inputs (rotary table rows / residual rows) arrive by TMA
// load, outputs leave by TMA store.
//   ROW    bias, scale, optional bf16 residual (x16 updated in place), bf16 out
//   HEADS  optional rotary (fp16 cos/sin pairs), per-part scale, head-major bf16 out [S,4,Lp,64]
//   LN     LayerNorm over the 512-wide row + erf-GELU; the row's columns live in the 4 CTAs of the
//          cluster, so per-row (sum, sum^2) partials are exchanged with st.async into every peer's
//          shared memory, completion counted by the peer's mbarrier (no fences, no cluster barrier)
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

// debug timeline (compile with -DLG_GEMM_DEBUG): accumulated wait cycles of cluster 0 / CTA rank 0
//   [0] MMA warp waiting for an accumulator (epilogue too slow)   [1] MMA warp waiting for A stages (loads too slow)
//   [2] MMA warp total   [3] epilogue warp 0 waiting for tfull   [4] epilogue warp 0 total   [5] tiles   [6] epilogue stats wait
__device__ long long g_gemm_times[16];

// waits of the TMA-producer and MMA-issuer warps: -DLG_GEMM_RELAXED polls with a sleep in between instead of spinning
// (measured on the whole step: 20.4-20.6 ms against 20.2-20.4 ms spinning, and one 48 ms outlier -- not adopted; the
// attention kernel, whose softmax warps compete with the waiting warps for issue slots, does gain from it)
#ifdef LG_GEMM_RELAXED
#define LG_PI_WAIT(...) tc::mbar_wait_relaxed(__VA_ARGS__)
#else
#define LG_PI_WAIT(...) tc::mbar_wait(__VA_ARGS__)
#endif

namespace {

constexpr int BM = 128, BK = 64, BN = 128;
constexpr int A_STAGE = BM * BK * 2;  // 16 KB
constexpr int WB_BYTES = BN * BK * 2; // one K block of the resident W (16 KB)
constexpr int STG = 4096;             // one staging tile: 32 rows x 128 B

enum { MODE_ROW = 0, MODE_HEADS = 1, MODE_LN = 2 };

template <bool KBIG>
struct Lay {
  static constexpr int NSTAGE = KBIG ? 3 : 4;
  static constexpr int W_BYTES = KBIG ? 128 * 1024 : 64 * 1024;
  static constexpr int OFF_A = W_BYTES;
  static constexpr int OFF_SOUT = OFF_A + NSTAGE * A_STAGE;
  static constexpr int OFF_SIN = KBIG ? OFF_SOUT : OFF_SOUT + 8 * STG;  // K = 512: input aliases output tile
  static constexpr int OFF_PAR = OFF_SIN + (KBIG ? 8 : 16) * STG;       // (K <= 256: two input tiles per warp, see the
                                                                        //  rotary prefetch) then bias | gamma | beta
  static constexpr int OFF_STATS = OFF_PAR + 3 * BN * 4;                // [2][8 slots][128 rows] float2
  static constexpr int OFF_BAR = OFF_STATS + (KBIG ? 2 * 8 * 128 * 8 : 0);
  static constexpr int SMEM = OFF_BAR + 448;  // up to 50 mbarriers (pair HEADS kernel: 4 stages, 32 input barriers) + the TMEM slot
};

struct Maps {
  CUtensorMap a0, a1, w, out0, out1, out2, in;
};

struct Args {
  int kb_total, kb_a0;
  int n_groups;  // column-block groups (clusters) per M tile: N / (128 * CL)
  int m_tiles;
  int Lp;
  int prefetch_tiles;  // L2 prefetch distance in M tiles (0 = off)
  int reverse;   // 1: the M tiles are walked from the last to the first.  Consecutive layers alternate the direction, so
                 // that a kernel starts on the rows its predecessor wrote last -- the part of a 134-268 MB activation
                 // that is still in the 126 MB L2
  int has_in;    // rotary table (HEADS) or residual (ROW) present
  int n_rot;
  float scale[3];
  const int32_t* lens;
  const float* bias;
  const float* gamma;
  const float* beta;
  void* out16;   // bf16 [T, N] row-major output (the LayerNorm pair kernel stores from registers)
};

// ---- cluster / DSMEM / bulk-copy primitives --------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t n_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
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
// remote 8-byte store whose completion is counted (in bytes) by the destination CTA's mbarrier
__device__ __forceinline__ void st_async_f2(uint32_t addr, float a, float b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr),
               "f"(a), "f"(b), "r"(mbar)
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
#ifndef LG_PAIR_2SM_TMA
// 1 (experiment, parity-green, no gain): in the CTA-pair kernels the A loads of BOTH CTAs of a pair complete on the LEADER's
// `full` barrier (cp.async.bulk.tensor ... .cta_group::2 with the pair bit of the barrier address cleared -- with multicast
// each destination's copy signals the leader of THAT destination's pair; the leader expects the bytes of both stages), so
// the issuer does not wait for a relay warp in the peer CTA to poll its own barrier and arrive remotely.  B200: the step is
// unchanged (20.93 / 20.94 vs 20.94 ms; FFN1 149 us, QKV 119 us under ncu) -- the relay's hand-over (1 047 of the FFN1
// issuer's 7 227 cycles per super-tile) overlaps the wait for the leader's own stage.  0: the relay (default).
#define LG_PAIR_2SM_TMA 0
#endif
constexpr uint32_t PAIR_LEADER_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(tc::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar) & PAIR_LEADER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(tc::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar) & PAIR_LEADER_MASK), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(tc::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// L2 prefetch of a tensor box: no shared memory, no barrier -- shortens the later smem fill from HBM
// latency to L2-hit latency (the A ring holds only 48-64 KB, ~40-50 % of what HBM latency would need)
__device__ __forceinline__ void tma_prefetch_l2(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// erf-GELU with the Abramowitz-Stegun 7.1.26 form (|err| < 2e-7, far below bf16 resolution):
// 2 MUFU + ~12 FMA instead of the ~30-instruction erff().
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

// tanh-form GELU, 5 FMA-pipe ops + 1 MUFU.  |gelu_tanh - gelu_erf| <= 4.7e-4 absolute (at |x| ~ 2), an eighth
// of a bf16 ulp there; with the erf form the FFN1 epilogue needed ~2560 issue cycles per 128x128 tile
// against 2048 tensor cycles.  bf16 mode only -- the fp32 parity path uses erff().
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = fmaf(x * x, 0.0356774081f, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * k));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

#ifdef LG_GELU_ERF
__device__ __forceinline__ float gelu_act(float x) { return gelu_fast(x); }
#else
__device__ __forceinline__ float gelu_act(float x) { return gelu_tanh(x); }
#endif

// tile index inside [0, m_tiles): one unsigned compare serves both walking directions
#define MT_IN(x) ((unsigned)(x) < (unsigned)g.m_tiles)

__device__ __forceinline__ bool tile_skipped(const Args& g, int m_tile) {
  if (!g.lens) return false;
  const int r0 = m_tile * BM;
  const int s = r0 / g.Lp;
  return r0 - s * g.Lp >= g.lens[s];
}

template <int MODE, int CL, bool KBIG>
__global__ void __launch_bounds__(320, 1)
tc_ws_linear_kernel(const __grid_constant__ Maps maps, const Args g) {
  using L = Lay<KBIG>;
  constexpr int NSTAGE = L::NSTAGE;
  constexpr int SLICE_ROWS = BM / CL;  // rows of an A stage this CTA loads (and multicasts)
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::OFF_A;
  float* s_par = reinterpret_cast<float*>(smem + L::OFF_PAR);
  float2* s_stats = reinterpret_cast<float2*>(smem + L::OFF_STATS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;  // [2]
  uint64_t* tempty = tfull + 2;      // [2]
  uint64_t* w_full = tempty + 2;
  uint64_t* stats_bar = w_full + 1;  // [2]
  uint64_t* in_bar = stats_bar + 2;  // [16] two per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bar + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CL > 1 ? cluster_rank() : 0;
  const int cid = (int)cluster_id_x(), ncl = (int)n_clusters_x();
  const int group = cid % g.n_groups;       // which CL-wide set of column blocks
  const int m_first = g.reverse ? g.m_tiles - 1 - cid / g.n_groups : cid / g.n_groups;
  const int m_step = g.reverse ? -(ncl / g.n_groups) : ncl / g.n_groups;
  const int nb = group * CL + (int)crank;   // 128-column block of this CTA
  const int n0 = nb * BN;

  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-byte aligned base

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&maps.a0);
    tc::prefetch_tmap(&maps.a1);
    tc::prefetch_tmap(&maps.w);
    tc::prefetch_tmap(&maps.out0);
    tc::prefetch_tmap(&maps.in);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], CL); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 8);
      tc::mbar_init(&stats_bar[i], 1);
    }
    for (int i = 0; i < 16; ++i) tc::mbar_init(&in_bar[i], 1);
    tc::mbar_init(w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
  for (int i = threadIdx.x; i < BN; i += blockDim.x) {
    s_par[i] = g.bias[n0 + i];
    if (MODE == MODE_LN) {
      s_par[BN + i] = g.gamma[n0 + i];
      s_par[2 * BN + i] = g.beta[n0 + i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the next kernel in the stream may start its own prologue (barriers, TMEM, weight block) while this one runs;
  // this kernel has read nothing of its predecessor's yet: weights, bias and LayerNorm parameters never change
  tc::pdl_launch_dependents();

  // Producer and MMA warps run their loops with warp-uniform control flow (all 32 lanes wait on the
  // barriers, one elected lane issues the TMA / tcgen05 instruction): addresses and descriptors then
  // stay in uniform registers.  Running the loop on lane 0 alone cost ~100 issue cycles per
  // tcgen05.mma (R2UR traffic), more than the 64 cycles the MMA itself takes.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(w_full, g.kb_total * WB_BYTES);
      for (int kb = 0; kb < g.kb_total; ++kb) tc::tma_load_2d(sW + kb * WB_BYTES, &maps.w, w_full, kb * BK, n0);
    }
    __syncwarp();
    tc::pdl_wait();  // activations (and lens) come from the predecessor; the weight block above does not
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (tile_skipped(g, mt)) continue;
      const int row = mt * BM + crank * SLICE_ROWS;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        LG_PI_WAIT(&empty[stage], phase ^ 1);  // all CL consumers released this stage
        uint8_t* dst = sA + stage * A_STAGE + crank * (SLICE_ROWS * 128);
        const CUtensorMap* tm = kb < g.kb_a0 ? &maps.a0 : &maps.a1;
        const int kc = (kb < g.kb_a0 ? kb : kb - g.kb_a0) * BK;
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full[stage], A_STAGE);
          if (CL > 1) tma_load_2d_mc(dst, tm, &full[stage], kc, row, MC_MASK);
          else tc::tma_load_2d(dst, tm, &full[stage], kc, row);
          if (g.prefetch_tiles > 0) {
            const int mp = mt + g.prefetch_tiles * m_step;  // same k-block, a few M tiles ahead
            if (MT_IN(mp)) tma_prefetch_l2(tm, kc, mp * BM + crank * SLICE_ROWS);
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    tc::pdl_wait();  // (reads lens)
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = tc::idesc_bf16(BM, BN, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    LG_PI_WAIT(w_full, 0);
    const uint64_t dW0 = tc::smem_desc_sw128(tc::smem_u32(sW), 0, 1024);
    const uint64_t dA0 = tc::smem_desc_sw128(tc::smem_u32(sA), 0, 1024);
#ifdef LG_GEMM_DEBUG
    long long w_acc = 0, w_full_c = 0, t_begin = clock64(), tt;
    const bool rec = blockIdx.x == 0 && lane == 0;
#define GT0() tt = clock64()
#define GT1(v) v += clock64() - tt
#else
#define GT0()
#define GT1(v)
#endif
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (tile_skipped(g, mt)) continue;
      GT0();
      LG_PI_WAIT(&tempty[acc], acc_phase ^ 1);
      GT1(w_acc);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        GT0();
        LG_PI_WAIT(&full[stage], phase);
        GT1(w_full_c);
        tc::fence_after_sync();
        const uint64_t dA = dA0 + (uint64_t)(stage * (A_STAGE >> 4));
        const uint64_t dW = dW0 + (uint64_t)(kb * (WB_BYTES >> 4));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) tc::umma_ss(d_tmem, dA + 2 * k, dW + 2 * k, idesc, (kb | k) != 0);
          if (CL > 1) umma_commit_mc(&empty[stage], MC_MASK);
          else tc::umma_commit(&empty[stage]);
          if (kb == g.kb_total - 1) tc::umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
#ifdef LG_GEMM_DEBUG
    if (rec) { g_gemm_times[0] = w_acc; g_gemm_times[1] = w_full_c; g_gemm_times[2] = clock64() - t_begin; }
#endif
  } else {
    tc::pdl_wait();  // residual / rotary rows, lens and the output buffers belong to the predecessor until here
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;        // 0..7
    const int quarter = warp & 3;   // TMEM lane quarter accessible to this warp
    const int half = ew >> 2;       // which 64-column half of the block
    const int c_warp = half * 64;   // first column (within the block) of this warp
    uint8_t* stg_out = smem + L::OFF_SOUT + ew * STG;
    uint8_t* stg_in = smem + L::OFF_SIN + ew * STG;
    const uint32_t my_row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    // HEADS: output map and head index are fixed for the CTA / warp
    const int col0 = n0 + c_warp;   // global output column of this warp's first column
    const int part = col0 >> 8, head = (col0 >> 6) & 3;
    const CUtensorMap* out_map = &maps.out0;
    float sc = g.scale[0];
    if (MODE == MODE_HEADS) {
      if (part == 1) { out_map = &maps.out1; sc = g.scale[1]; }
      if (part == 2) { out_map = &maps.out2; sc = g.scale[2]; }
    }
    const bool use_in = g.has_in && (MODE == MODE_ROW || (MODE == MODE_HEADS && part < g.n_rot));
    int acc = 0, iter = 0;
    uint32_t acc_phase = 0;
#ifdef LG_GEMM_DEBUG
    long long e_wait = 0, e_stats = 0, e_begin = clock64(), tt;
    const bool rec = blockIdx.x == 0 && ew == 0 && lane == 0;
#endif
    // HEADS: the rotary rows of the NEXT tile are fetched while this tile is processed (two input tiles per warp).
    // Fetching them at the top of the tile's own iteration left the ~1.5 us load exposed whenever the accumulator
    // was already waiting (the QKV projection was epilogue-bound: the issuer waited 38 % of its time for TMEM).
    constexpr bool PF = MODE == MODE_HEADS && !KBIG;
    auto next_tile = [&](int mt) {
      for (mt += m_step; MT_IN(mt) && tile_skipped(g, mt); mt += m_step) {}
      return mt;
    };
    if (PF && use_in && lane == 0) {
      int mt0 = m_first;
      if (MT_IN(mt0) && tile_skipped(g, mt0)) mt0 = next_tile(mt0);
      if (MT_IN(mt0)) {
        tc::mbar_arrive_expect_tx(&in_bar[ew], STG);
        tc::tma_load_2d(stg_in, &maps.in, &in_bar[ew], 0, mt0 * BM + quarter * 32);
      }
    }
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (tile_skipped(g, mt)) continue;
      const int row0 = mt * BM + quarter * 32;  // first global row of this warp
      if (PF) {
        if (use_in && lane == 0) {
          const int mn = next_tile(mt);
          if (MT_IN(mn)) {  // buffer (iter+1)&1 was read in the previous iteration (all lanes, then __syncwarp)
            const int nb_ = (iter + 1) & 1;
            tc::mbar_arrive_expect_tx(&in_bar[nb_ * 8 + ew], STG);
            tc::tma_load_2d(stg_in + nb_ * 8 * STG, &maps.in, &in_bar[nb_ * 8 + ew], 0, mn * BM + quarter * 32);
          }
        }
      } else if (use_in && lane == 0) {
        if (KBIG) bulk_wait_read0();            // input tile aliases the previous output tile
        tc::mbar_arrive_expect_tx(&in_bar[ew], STG);
        if (MODE == MODE_ROW) tc::tma_load_2d(stg_in, &maps.in, &in_bar[ew], col0, row0);  // residual rows
        else tc::tma_load_2d(stg_in, &maps.in, &in_bar[ew], 0, row0);                       // rotary rows
      }
      if (MODE == MODE_LN && ew == 0 && lane == 0)
        tc::mbar_arrive_expect_tx(&stats_bar[iter & 1], 2 * CL * 128 * 8);  // partials from every peer warp
      GT0();
      LG_PI_WAIT(&tfull[acc], acc_phase);
      GT1(e_wait);
      tc::fence_after_sync();
      uint32_t v[64];
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c_warp;
      tc::tmem_ld32(t_addr, v);
      tc::tmem_ld32(t_addr + 32, v + 32);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);  // accumulator is in registers: release it now
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      uint32_t pk[32];  // 64 bf16 outputs of this thread's row
      if constexpr (MODE == MODE_LN) {
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float x = __uint_as_float(v[j]) + s_par[c_warp + j];
          v[j] = __float_as_uint(x);
          sum += x;
          sq = fmaf(x, x, sq);
        }
        const int buf = iter & 1;
        const int r_in_tile = quarter * 32 + lane;
        const uint32_t slot = tc::smem_u32(&s_stats[(buf * 8 + (int)crank * 2 + half) * 128 + r_in_tile]);
        const uint32_t bar = tc::smem_u32(&stats_bar[buf]);
#pragma unroll
        for (int p = 0; p < CL; ++p) st_async_f2(map_to_rank(slot, p), sum, sq, map_to_rank(bar, p));
        GT0();
        LG_PI_WAIT(&stats_bar[buf], (iter >> 1) & 1);
        GT1(e_stats);
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
        const float nmr = -mean * rstd;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = c_warp + 2 * j;
          const float a = gelu_act(fmaf(fmaf(__uint_as_float(v[2 * j]), rstd, nmr), s_par[BN + c], s_par[2 * BN + c]));
          const float b = gelu_act(fmaf(fmaf(__uint_as_float(v[2 * j + 1]), rstd, nmr), s_par[BN + c + 1], s_par[2 * BN + c + 1]));
          pk[j] = tc::pack_bf16(a, b);
        }
      } else {
        uint32_t in[32];
        if (use_in) {
          const uint8_t* tin = stg_in;
          if (PF) {
            LG_PI_WAIT(&in_bar[(iter & 1) * 8 + ew], (iter >> 1) & 1);
            tin = stg_in + (iter & 1) * 8 * STG;
          } else {
            LG_PI_WAIT(&in_bar[ew], iter & 1);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 t = *reinterpret_cast<const uint4*>(tin + my_row_off + ((j ^ sw) << 4));
            in[4 * j] = t.x; in[4 * j + 1] = t.y; in[4 * j + 2] = t.z; in[4 * j + 3] = t.w;
          }
          __syncwarp();  // every lane has read the input tile before anyone overwrites it (aliasing)
        }
        if constexpr (MODE == MODE_ROW) {
          if (use_in) {  // residual add (branch hoisted out of the unrolled loop)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const __nv_bfloat162 r = *reinterpret_cast<const __nv_bfloat162*>(&in[j]);
              const float a = fmaf(__uint_as_float(v[2 * j]) + s_par[c_warp + 2 * j], sc, __low2float(r));
              const float b = fmaf(__uint_as_float(v[2 * j + 1]) + s_par[c_warp + 2 * j + 1], sc, __high2float(r));
              pk[j] = tc::pack_bf16(a, b);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              pk[j] = tc::pack_bf16((__uint_as_float(v[2 * j]) + s_par[c_warp + 2 * j]) * sc,
                                    (__uint_as_float(v[2 * j + 1]) + s_par[c_warp + 2 * j + 1]) * sc);
          }
        } else {  // MODE_HEADS: pair j = head-dim (2j, 2j+1) rotates by frequency j
          if (use_in) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = __uint_as_float(v[2 * j]) + s_par[c_warp + 2 * j];
              const float b = __uint_as_float(v[2 * j + 1]) + s_par[c_warp + 2 * j + 1];
              const __half2 cs = *reinterpret_cast<const __half2*>(&in[j]);
              const float c = __low2float(cs) * sc, s = __high2float(cs) * sc;  // scale folded into cos/sin
              pk[j] = tc::pack_bf16(a * c - b * s, b * c + a * s);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              pk[j] = tc::pack_bf16((__uint_as_float(v[2 * j]) + s_par[c_warp + 2 * j]) * sc,
                                    (__uint_as_float(v[2 * j + 1]) + s_par[c_warp + 2 * j + 1]) * sc);
          }
        }
      }
      // stage the 32 x 64 bf16 tile (128-byte swizzle) and hand it to the TMA store engine
      if (!(KBIG && use_in)) {
        if (lane == 0) bulk_wait_read0();  // previous store has finished reading stg_out
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg_out + my_row_off + ((j ^ sw) << 4)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (MODE == MODE_HEADS) {
          const int s = (mt * BM) / g.Lp, l0 = mt * BM - s * g.Lp;
          tma_store_2d(out_map, stg_out, 0, (s * LG_HEADS + head) * g.Lp + l0 + quarter * 32);
        } else {
          tma_store_2d(out_map, stg_out, col0, row0);
        }
        bulk_commit();
      }
      ++iter;
    }
    if (lane == 0) bulk_wait0();  // all stores of this warp have landed before the CTA retires
#ifdef LG_GEMM_DEBUG
    if (rec) { g_gemm_times[3] = e_wait; g_gemm_times[4] = clock64() - e_begin; g_gemm_times[5] = iter; g_gemm_times[6] = e_stats; }
#endif
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody exits while peers may still multicast / st.async here
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

// =====================================================================================================
// CTA-pair variant (tcgen05 cta_group::2) for the ROW epilogue at K = 512, N = 256 (second FFN layer).
//
// Why: the 1-CTA kernel above issues 128x128x16 SS MMAs, which read 8 KB of operands from shared memory per 64
// tensor cycles = 128 B/clk, the whole shared-memory bandwidth of the SM -- and the TMA fill of the A ring and the
// epilogue staging share it.  Measured (LG_GEMM_DEBUG): the issuer needs ~4000 cycles for the 32 MMAs of a K = 512
// tile against 2048 tensor cycles.  A CTA pair computes a 256 x 256 tile with M = 256 MMAs: each CTA supplies its
// own 128 rows of A and HALF of B (its resident 128-column block of W), so every operand byte read from shared
// memory feeds twice the flops (64 B/clk/SM), and an A stage covers twice the tensor time (the 3-stage ring hides
// twice the load latency).
//   cluster = 2 CTAs = one pair; rank 0 (leader) issues all MMAs and commits, rank 1 relays "my A stage / my W is
//   loaded" to the leader with remote mbarrier arrives; both run the producer and the epilogue for their own
//   128 rows x 256 columns (accumulators: 2 x 256 TMEM columns per CTA).
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // one full warp, in each CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  const uint32_t ra = map_to_rank(tc::smem_u32(bar), rank);
  // relaxed: the relay publishes nothing of its own (the data were written by TMA / read through tcgen05);
  // a release at cluster scope costs a MEMBAR + L1 invalidate (~1300 cycles per arrive, measured)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait with cluster-scope acquire (the arrival may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE_C;\n\t"
      "bra.uni WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}

// both 128-row halves of super-tile `mt2` (256 rows) lie beyond their sequence's valid rows
__device__ __forceinline__ bool pair_skipped(const Args& g, int mt2) {
  return tile_skipped(g, 2 * mt2) && tile_skipped(g, 2 * mt2 + 1);
}

template <int MODE, bool KBIG>
__global__ void __launch_bounds__(MODE == MODE_HEADS ? 576 : 320, 1)
tc_pair_row_kernel(const __grid_constant__ Maps maps, const Args g) {
  using L = Lay<KBIG>;
  constexpr int NSTAGE = L::NSTAGE;
  constexpr int BN2 = 256;  // columns of the pair's tile = accumulator columns per CTA
  constexpr int NEW = MODE == MODE_HEADS ? 16 : 8;  // epilogue warps (HEADS was epilogue-bound with 8: issuer waited 56 % for TMEM)

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::OFF_A;
  float* s_par = reinterpret_cast<float*>(smem + L::OFF_PAR);  // bias of the pair's 256 columns
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + NSTAGE;
  uint64_t* peer_full = empty + NSTAGE;  // leader only: the peer's stage is loaded
  uint64_t* tfull = peer_full + NSTAGE;  // [2]
  uint64_t* tempty = tfull + 2;          // [2] leader only: 16 epilogue warps of the pair
  uint64_t* w_full = tempty + 2;
  uint64_t* peer_w = w_full + 1;
  uint64_t* in_bar = peer_w + 1;         // [2 * NEW] two per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bar + 2 * NEW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_rank();  // 0 = leader
  const int cid = (int)cluster_id_x(), ncl = (int)n_clusters_x();
  const int group = cid % g.n_groups;     // which 256-column block
  const int m_first = g.reverse ? g.m_tiles - 1 - cid / g.n_groups : cid / g.n_groups;
  const int m_step = g.reverse ? -(ncl / g.n_groups) : ncl / g.n_groups;
  const int n0 = (group * 2 + (int)crank) * BN;  // this CTA's resident 128-column block of W
  const int pair_col0 = group * BN2;

  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&maps.a0);
    tc::prefetch_tmap(&maps.a1);
    tc::prefetch_tmap(&maps.w);
    tc::prefetch_tmap(&maps.out0);
    tc::prefetch_tmap(&maps.out1);
    tc::prefetch_tmap(&maps.out2);
    tc::prefetch_tmap(&maps.in);
    for (int i = 0; i < NSTAGE; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 1);
      tc::mbar_init(&peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 2 * NEW);
    }
    for (int i = 0; i < 2 * NEW; ++i) tc::mbar_init(&in_bar[i], 1);
    tc::mbar_init(w_full, 1);
    tc::mbar_init(peer_w, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  for (int i = threadIdx.x; i < BN2; i += blockDim.x) s_par[i] = g.bias[pair_col0 + i];
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before any cross-CTA traffic
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the next kernel in the stream may start its own prologue (barriers, TMEM, weight block) while this one runs;
  // this kernel has read nothing of its predecessor's yet: weights, bias and LayerNorm parameters never change
  tc::pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (own 128 rows, own W block)
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(w_full, g.kb_total * WB_BYTES);
      for (int kb = 0; kb < g.kb_total; ++kb) tc::tma_load_2d(sW + kb * WB_BYTES, &maps.w, w_full, kb * BK, n0);
    }
    __syncwarp();
    tc::pdl_wait();  // activations (and lens) come from the predecessor; the weight block above does not
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (pair_skipped(g, mt)) continue;
      const int row = mt * 2 * BM + (int)crank * BM;
      for (int kb = 0; kb < g.kb_total; ++kb) {
        LG_PI_WAIT(&empty[stage], phase ^ 1);  // the pair's MMAs have read this stage (in both CTAs)
        const CUtensorMap* tm = kb < g.kb_a0 ? &maps.a0 : &maps.a1;
        const int kc = (kb < g.kb_a0 ? kb : kb - g.kb_a0) * BK;
        if (tc::elect_one()) {
#if LG_PAIR_2SM_TMA
          if (crank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * A_STAGE);  // this stage here and in the peer
          tma_load_2d_2sm(sA + stage * A_STAGE, tm, &full[stage], kc, row);
#else
          tc::mbar_arrive_expect_tx(&full[stage], A_STAGE);
          tc::tma_load_2d(sA + stage * A_STAGE, tm, &full[stage], kc, row);
#endif
          if (g.prefetch_tiles > 0) {
            const int mp = mt + g.prefetch_tiles * m_step;
            if (MT_IN(mp)) tma_prefetch_l2(tm, kc, mp * 2 * BM + (int)crank * BM);
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    tc::pdl_wait();  // (reads lens)
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    LG_PI_WAIT(w_full, 0);
    if (crank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA)
      constexpr uint32_t idesc = tc::idesc_bf16(256, BN2, 0);
      mbar_wait_cluster(peer_w, 0);
      const uint64_t dW0 = tc::smem_desc_sw128(tc::smem_u32(sW), 0, 1024);
      const uint64_t dA0 = tc::smem_desc_sw128(tc::smem_u32(sA), 0, 1024);
#ifdef LG_GEMM_DEBUG
      long long w_acc = 0, w_full_c = 0, w_peer = 0, t_begin = clock64(), tt;
      int n_t = 0;
      const bool rec = blockIdx.x == 0 && lane == 0;
#endif
      for (int mt = m_first; MT_IN(mt); mt += m_step) {
        if (pair_skipped(g, mt)) continue;
        GT0();
        mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);
        GT1(w_acc);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          GT0();
          LG_PI_WAIT(&full[stage], phase);
          GT1(w_full_c);
#if !LG_PAIR_2SM_TMA
          GT0();
          mbar_wait_cluster(&peer_full[stage], phase);
          GT1(w_peer);
#endif
          tc::fence_after_sync();
          const uint64_t dA = dA0 + (uint64_t)(stage * (A_STAGE >> 4));
          const uint64_t dW = dW0 + (uint64_t)(kb * (WB_BYTES >> 4));
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss2(d_tmem, dA + 2 * k, dW + 2 * k, idesc, (kb | k) != 0);
            umma_commit2_mc(&empty[stage], 3);
            if (kb == g.kb_total - 1) umma_commit2_mc(&tfull[acc], 3);
          }
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef LG_GEMM_DEBUG
        ++n_t;
#endif
      }
#ifdef LG_GEMM_DEBUG
      if (rec) { g_gemm_times[0] = w_acc; g_gemm_times[1] = w_full_c; g_gemm_times[7] = w_peer; g_gemm_times[2] = clock64() - t_begin; g_gemm_times[5] = n_t; g_gemm_times[3] = 0; g_gemm_times[4] = 0; g_gemm_times[6] = 0; }
#endif
    } else {
      // ---------------------------------------------------------------- relay (peer CTA): tell the leader what has landed here
      if (lane == 0) mbar_arrive_remote(peer_w, 0);
#if !LG_PAIR_2SM_TMA
      for (int mt = m_first; MT_IN(mt); mt += m_step) {
        if (pair_skipped(g, mt)) continue;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          LG_PI_WAIT(&full[stage], phase);
          if (lane == 0) mbar_arrive_remote(&peer_full[stage], 0);
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
#endif
    }
  } else {
    tc::pdl_wait();  // residual / rotary rows, lens and the output buffers belong to the predecessor until here
    // ------------------------------------------------------------------ epilogue (8 warps): own 128 rows x 256 columns
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;       // ROW: 64-column half; HEADS: 32-column slice cq (0..3) of a 128-column block
    const int c_warp = half * (MODE == MODE_HEADS ? 32 : 64);
    constexpr int STGW = MODE == MODE_HEADS ? 2048 : STG;  // staging tile of a warp: 32 rows x 64 B or x 128 B
    uint8_t* stg_out = smem + L::OFF_SOUT + ew * STGW;
    uint8_t* stg_in = smem + L::OFF_SIN + ew * STGW;
    const uint32_t my_row_off = (uint32_t)lane * (MODE == MODE_HEADS ? 64u : 128u);
    const uint32_t sw = MODE == MODE_HEADS ? (uint32_t)((lane >> 1) & 3) : (uint32_t)(lane & 7);
    if constexpr (MODE == MODE_HEADS) {
      // head-major outputs [S,4,Lp,64]; a warp owns 32 tokens x 32 columns (half a head) of each 128-column block:
      // 2 KB staging tiles with a 64-byte swizzle, 32 x 32 TMA boxes.  Optional rotary on the first n_rot parts; the
      // (cos, sin) pairs of the warp's 16 frequencies for the NEXT super-tile are fetched while this one is processed.
      const bool use_rot = g.has_in != 0 && (pair_col0 >> 8) < g.n_rot;  // (a pair's 256 columns are one part)
      const int part = pair_col0 >> 8;
      const CUtensorMap* out_map = part == 0 ? &maps.out0 : (part == 1 ? &maps.out1 : &maps.out2);
      const float sc = g.scale[part];
      const int col_in_head = (half & 1) * 32;  // column of this slice inside its 64-wide head row
      auto next_tile = [&](int mt) {
        for (mt += m_step; MT_IN(mt) && pair_skipped(g, mt); mt += m_step) {}
        return mt;
      };
      int acc = 0, iter = 0;
      uint32_t acc_phase = 0;
      if (use_rot && lane == 0) {
        int mt0 = m_first;
        if (MT_IN(mt0) && pair_skipped(g, mt0)) mt0 = next_tile(mt0);
        if (MT_IN(mt0)) {
          tc::mbar_arrive_expect_tx(&in_bar[ew], STGW);
          tc::tma_load_2d(stg_in, &maps.in, &in_bar[ew], col_in_head, mt0 * 2 * BM + (int)crank * BM + quarter * 32);
        }
      }
      for (int mt = m_first; MT_IN(mt); mt += m_step) {
        if (pair_skipped(g, mt)) continue;
        const int r0 = mt * 2 * BM + (int)crank * BM;       // first row of this CTA's 128-row half (inside one sequence)
        const bool store_rows = !tile_skipped(g, 2 * mt + (int)crank);
        const int seq = r0 / g.Lp, l0 = r0 - seq * g.Lp;
        if (use_rot && lane == 0) {
          const int mn = next_tile(mt);
          if (MT_IN(mn)) {  // buffer (iter+1)&1 was read during the previous super-tile
            const int nb_ = (iter + 1) & 1;
            tc::mbar_arrive_expect_tx(&in_bar[nb_ * NEW + ew], STGW);
            tc::tma_load_2d(stg_in + nb_ * NEW * STGW, &maps.in, &in_bar[nb_ * NEW + ew], col_in_head,
                            mn * 2 * BM + (int)crank * BM + quarter * 32);
          }
        }
        LG_PI_WAIT(&tfull[acc], acc_phase);
        tc::fence_after_sync();
        uint32_t in[16];
        if (use_rot) {
          LG_PI_WAIT(&in_bar[(iter & 1) * NEW + ew], (iter >> 1) & 1);
          const uint8_t* tin = stg_in + (iter & 1) * NEW * STGW;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 t = *reinterpret_cast<const uint4*>(tin + my_row_off + ((j ^ sw) << 4));
            in[4 * j] = t.x; in[4 * j + 1] = t.y; in[4 * j + 2] = t.z; in[4 * j + 3] = t.w;
          }
        }
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int cw = cb * BN + c_warp;               // first column of this warp within the pair's 256
          const int head = ((pair_col0 + cw) >> 6) & 3;
          uint32_t v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2 + cw, v);
          tc::tmem_ld_wait();
          if (cb == 1) {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
              if (crank == 0) tc::mbar_arrive(&tempty[acc]);
              else mbar_arrive_remote(&tempty[acc], 0);
            }
          }
          uint32_t pk[16];
          if (use_rot) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = __uint_as_float(v[2 * j]) + s_par[cw + 2 * j];
              const float b = __uint_as_float(v[2 * j + 1]) + s_par[cw + 2 * j + 1];
              const __half2 cs = *reinterpret_cast<const __half2*>(&in[j]);
              const float c = __low2float(cs) * sc, s_ = __high2float(cs) * sc;
              pk[j] = tc::pack_bf16(a * c - b * s_, b * c + a * s_);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = tc::pack_bf16((__uint_as_float(v[2 * j]) + s_par[cw + 2 * j]) * sc,
                                    (__uint_as_float(v[2 * j + 1]) + s_par[cw + 2 * j + 1]) * sc);
          }
          if (lane == 0) bulk_wait_read0();  // previous store has finished reading stg_out
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(stg_out + my_row_off + ((j ^ sw) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && store_rows) {
            tma_store_2d(out_map, stg_out, col_in_head, (seq * LG_HEADS + head) * g.Lp + l0 + quarter * 32);
            bulk_commit();
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        ++iter;
      }
    } else {
    const float sc = g.scale[0];
    const bool use_in = g.has_in != 0;
    int acc = 0;
    uint32_t acc_phase = 0, n_in = 0;
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (pair_skipped(g, mt)) continue;
      const int row0 = mt * 2 * BM + (int)crank * BM + quarter * 32;
      const bool store_rows = !tile_skipped(g, 2 * mt + (int)crank);  // rows of a fully padded half stay untouched
#pragma unroll 1
      for (int cb = 0; cb < 2; ++cb) {
        const int cw = cb * BN + c_warp;        // first column of this warp within the pair's 256
        const int colg = pair_col0 + cw;        // global output column
        if (use_in && lane == 0) {
          if (KBIG) bulk_wait_read0();          // the input tile aliases the previous output tile
          tc::mbar_arrive_expect_tx(&in_bar[ew], STG);
          tc::tma_load_2d(stg_in, &maps.in, &in_bar[ew], colg, row0);  // residual rows
        }
        if (cb == 0) {
          LG_PI_WAIT(&tfull[acc], acc_phase);
          tc::fence_after_sync();
        }
        uint32_t v[64];
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2 + cw;
        tc::tmem_ld32(t_addr, v);
        tc::tmem_ld32(t_addr + 32, v + 32);
        tc::tmem_ld_wait();
        if (cb == 1) {  // the whole accumulator is in registers / stored: release it to the leader's issuer
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (crank == 0) tc::mbar_arrive(&tempty[acc]);
            else mbar_arrive_remote(&tempty[acc], 0);
          }
        }
        uint32_t pk[32];
        uint32_t in[32];
        if (use_in) {
          LG_PI_WAIT(&in_bar[ew], n_in & 1);
          ++n_in;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 t = *reinterpret_cast<const uint4*>(stg_in + my_row_off + ((j ^ sw) << 4));
            in[4 * j] = t.x; in[4 * j + 1] = t.y; in[4 * j + 2] = t.z; in[4 * j + 3] = t.w;
          }
          __syncwarp();  // every lane has read the input tile before anyone overwrites it (aliasing)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const __nv_bfloat162 r = *reinterpret_cast<const __nv_bfloat162*>(&in[j]);
            const float a = fmaf(__uint_as_float(v[2 * j]) + s_par[cw + 2 * j], sc, __low2float(r));
            const float b = fmaf(__uint_as_float(v[2 * j + 1]) + s_par[cw + 2 * j + 1], sc, __high2float(r));
            pk[j] = tc::pack_bf16(a, b);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            pk[j] = tc::pack_bf16((__uint_as_float(v[2 * j]) + s_par[cw + 2 * j]) * sc,
                                  (__uint_as_float(v[2 * j + 1]) + s_par[cw + 2 * j + 1]) * sc);
        }
        if (!(KBIG && use_in)) {
          if (lane == 0) bulk_wait_read0();  // previous store has finished reading stg_out
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg_out + my_row_off + ((j ^ sw) << 4)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && store_rows) {
          tma_store_2d(&maps.out0, stg_out, colg, row0);
          bulk_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
    if (lane == 0) bulk_wait0();
  }
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the pair's MMAs / remote arrives may still touch it
  if (warp == 1) {
    tc::fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

// -----------------------------------------------------------------------------------------------------
// CTA-pair kernel for the first FFN layer: K = 512 (two 256-wide sources), N = 512, LayerNorm + GELU epilogue.
// Cluster of 4 = two pairs: pair p = ranks {2p, 2p+1} computes rows [0,256) x columns [256p, 256p+256) of the
// cluster's 256-row super-tile.  Rank r holds W columns [128r, 128r+128) (resident) and A rows of parity r&1;
// the two CTAs of one parity (r, r^2) each load half of that 128-row A stage and multicast it to both.
// LayerNorm: a row's 512 columns live in CTAs r and r^2 (256 each); per-row (sum, sum^2) partials are exchanged
// with st.async.  The 256 accumulator columns of a CTA do not fit in registers, so the epilogue reads TMEM twice:
// pass 1 statistics, pass 2 normalise + GELU + store.
#ifndef LG_FFN1_DIRECT_STORE
// 1 (experiment, measured and not adopted): the epilogue writes its bf16 rows straight from registers with 256-bit global
// stores (STG.256: a lane owns 32 consecutive columns of one row = two full 32-byte sectors) instead of staging 16 x 2 KB
// in shared memory for TMA stores, and the 32 KB go to the A ring (3 -> 5 stages of 16 KB; the issuer spends 2 655 of its
// 7 227 cycles per 256-row super-tile waiting for A stages).  B200, T = 262 144: 190 us against 141 us with the TMA stores
// (whole step 21.6 vs 21.2 ms) -- a warp store that touches 32 different rows costs the LSU more than the deeper ring
// gains (profiles/r2_ffn1_direct_store_ab.txt).
#define LG_FFN1_DIRECT_STORE 0
#endif
#ifndef LG_FFN1_HALF_STAGING
// 1 (experiment, measured and not adopted): the 16 epilogue warps stage 16 columns at a time (32 rows x 32 B, 32-byte
// swizzle, two TMA stores per 32-column chunk) instead of 32: 16 KB instead of 32 KB of staging, which buys a fourth A
// stage.  B200: 156 us against 141 us under ncu, whole step 20.79 / 20.85 vs 20.78 / 20.85 ms -- the second
// wait-for-read + store per chunk costs what the deeper ring gains.
#define LG_FFN1_HALF_STAGING 0
#endif
struct LayPL {
  static constexpr int NSTAGE = LG_FFN1_DIRECT_STORE ? 5 : LG_FFN1_HALF_STAGING ? 4 : 3;
  static constexpr int STG_W = LG_FFN1_HALF_STAGING ? 1024 : 2048;  // staging bytes per epilogue warp
  static constexpr int W_BYTES = 128 * 1024;
  static constexpr int OFF_A = W_BYTES;
  static constexpr int OFF_SOUT = OFF_A + NSTAGE * A_STAGE;
  static constexpr int OFF_PAR = OFF_SOUT + (LG_FFN1_DIRECT_STORE ? 0 : 16 * STG_W);  // 16 epilogue warps x (32 rows x 64 B); then bias | gamma | beta, 3 x 256 floats
  static constexpr int OFF_STATS = OFF_PAR + 3 * 256 * 4;   // [2 bufs][2 CTAs][128 rows] float2 | local scratch [4][128] float2
  static constexpr int OFF_BAR = OFF_STATS + 2 * 2 * 128 * 8 + 2 * 4 * 128 * 8;  // exchanged partials + 2 local scratch buffers
  static constexpr int SMEM = OFF_BAR + 256;
};
static_assert(LayPL::SMEM <= 227 * 1024, "pair-LN kernel: shared memory");

__device__ __forceinline__ void st_global_256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__global__ void __launch_bounds__(576, 1)
tc_pair_ln_kernel(const __grid_constant__ Maps maps, const Args g) {
  using L = LayPL;
  constexpr int NSTAGE = L::NSTAGE;
  constexpr int BN2 = 256;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L::OFF_A;
  float* s_par = reinterpret_cast<float*>(smem + L::OFF_PAR);
  float2* s_stats = reinterpret_cast<float2*>(smem + L::OFF_STATS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + NSTAGE;
  uint64_t* peer_full = empty + NSTAGE;
  uint64_t* tfull = peer_full + NSTAGE;  // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* w_full = tempty + 2;
  uint64_t* peer_w = w_full + 1;
  uint64_t* stats_bar = peer_w + 1;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stats_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_rank();        // 0..3
  const uint32_t parity = crank & 1;            // A row half; 0 = leader of its pair
  const uint32_t pair = crank >> 1;             // column half of the 512-wide output
  const uint32_t partner = crank ^ 2;           // same rows, other 256 columns
  const int cid = (int)cluster_id_x(), ncl = (int)n_clusters_x();
  const int m_first = g.reverse ? g.m_tiles - 1 - cid : cid;  // (see Args::reverse)
  const int m_step = g.reverse ? -ncl : ncl;
  const int n0 = (int)crank * BN;               // resident W block
  const int pair_col0 = (int)pair * BN2;
  const uint16_t mc_mask = (uint16_t)((1u << crank) | (1u << partner));

  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&maps.a0);
    tc::prefetch_tmap(&maps.a1);
    tc::prefetch_tmap(&maps.w);
    tc::prefetch_tmap(&maps.out0);
    for (int i = 0; i < NSTAGE; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&empty[i], 2);  // both pairs of the cluster have consumed the stage
      tc::mbar_init(&peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 32);  // 16 epilogue warps in each CTA of the pair
      tc::mbar_init(&stats_bar[i], 1);
    }
    tc::mbar_init(w_full, 1);
    tc::mbar_init(peer_w, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  for (int i = threadIdx.x; i < BN2; i += blockDim.x) {
    s_par[i] = g.bias[pair_col0 + i];
    s_par[BN2 + i] = g.gamma[pair_col0 + i];
    s_par[2 * BN2 + i] = g.beta[pair_col0 + i];
  }
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the next kernel in the stream may start its own prologue (barriers, TMEM, weight block) while this one runs;
  // this kernel has read nothing of its predecessor's yet: weights, bias and LayerNorm parameters never change
  tc::pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: half of this parity's A stage, multicast
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(w_full, g.kb_total * WB_BYTES);
      for (int kb = 0; kb < g.kb_total; ++kb) tc::tma_load_2d(sW + kb * WB_BYTES, &maps.w, w_full, kb * BK, n0);
    }
    __syncwarp();
    tc::pdl_wait();  // activations (and lens) come from the predecessor; the weight block above does not
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (pair_skipped(g, mt)) continue;
      const int row = mt * 2 * BM + (int)parity * BM + (int)pair * 64;  // this CTA loads 64 of the 128 rows
      for (int kb = 0; kb < g.kb_total; ++kb) {
        LG_PI_WAIT(&empty[stage], phase ^ 1);
        const CUtensorMap* tm = kb < g.kb_a0 ? &maps.a0 : &maps.a1;
        const int kc = (kb < g.kb_a0 ? kb : kb - g.kb_a0) * BK;
        if (tc::elect_one()) {
#if LG_PAIR_2SM_TMA
          // each destination's copy completes on the barrier of that destination's pair leader: a leader collects its own
          // stage (two 8 KB halves, from ranks r and r^2) and its peer's
          if (parity == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * A_STAGE);
          tma_load_2d_mc_2sm(sA + stage * A_STAGE + (int)pair * (64 * 128), tm, &full[stage], kc, row, mc_mask);
#else
          tc::mbar_arrive_expect_tx(&full[stage], A_STAGE);
          tma_load_2d_mc(sA + stage * A_STAGE + (int)pair * (64 * 128), tm, &full[stage], kc, row, mc_mask);
#endif
          if (g.prefetch_tiles > 0) {
            const int mp = mt + g.prefetch_tiles * m_step;
            if (MT_IN(mp)) tma_prefetch_l2(tm, kc, mp * 2 * BM + (int)parity * BM + (int)pair * 64);
          }
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    tc::pdl_wait();  // (reads lens)
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    LG_PI_WAIT(w_full, 0);
    if (parity == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader of the pair)
      constexpr uint32_t idesc = tc::idesc_bf16(256, BN2, 0);
      const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
      mbar_wait_cluster(peer_w, 0);
      const uint64_t dW0 = tc::smem_desc_sw128(tc::smem_u32(sW), 0, 1024);
      const uint64_t dA0 = tc::smem_desc_sw128(tc::smem_u32(sA), 0, 1024);
#ifdef LG_GEMM_DEBUG
      long long w_acc = 0, w_full_c = 0, w_peer = 0, t_begin = clock64(), tt;
      int n_t = 0;
      const bool rec = blockIdx.x == 0 && lane == 0;
#endif
      for (int mt = m_first; MT_IN(mt); mt += m_step) {
        if (pair_skipped(g, mt)) continue;
        GT0();
        mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);
        GT1(w_acc);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          GT0();
          LG_PI_WAIT(&full[stage], phase);
          GT1(w_full_c);
#if !LG_PAIR_2SM_TMA
          GT0();
          mbar_wait_cluster(&peer_full[stage], phase);
          GT1(w_peer);
#endif
          tc::fence_after_sync();
          const uint64_t dA = dA0 + (uint64_t)(stage * (A_STAGE >> 4));
          const uint64_t dW = dW0 + (uint64_t)(kb * (WB_BYTES >> 4));
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss2(d_tmem, dA + 2 * k, dW + 2 * k, idesc, (kb | k) != 0);
            umma_commit2_mc(&empty[stage], 0xF);  // every CTA of the cluster fills stages of this pair's CTAs
            if (kb == g.kb_total - 1) umma_commit2_mc(&tfull[acc], pair_mask);
          }
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef LG_GEMM_DEBUG
        ++n_t;
#endif
      }
#ifdef LG_GEMM_DEBUG
      if (rec) { g_gemm_times[0] = w_acc; g_gemm_times[1] = w_full_c; g_gemm_times[7] = w_peer; g_gemm_times[2] = clock64() - t_begin; g_gemm_times[5] = n_t; }
#endif
    } else {
      // ---------------------------------------------------------------- relay (odd rank): tell the pair's leader what has landed here
      if (lane == 0) mbar_arrive_remote(peer_w, crank - 1);
#if !LG_PAIR_2SM_TMA
      for (int mt = m_first; MT_IN(mt); mt += m_step) {
        if (pair_skipped(g, mt)) continue;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          LG_PI_WAIT(&full[stage], phase);
          if (lane == 0) mbar_arrive_remote(&peer_full[stage], crank - 1);
          __syncwarp();
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
#endif
    }
  } else {
    tc::pdl_wait();  // residual / rotary rows, lens and the output buffers belong to the predecessor until here
    // ------------------------------------------------------------------ epilogue (16 warps): own 128 rows x 256 columns
    // warp = (TMEM lane quarter, 32-column slice cq of each 128-column block).  With 8 warps (2 per SM sub-partition)
    // the epilogue was busy 6600 cycles per 256-row super-tile against ~3500 issue cycles of work.
    const int ew = warp - 2;        // 0..15
    const int quarter = warp & 3;
    const int cq = ew >> 2;         // 0..3
#if !LG_FFN1_DIRECT_STORE
    uint8_t* stg_out = smem + L::OFF_SOUT + ew * L::STG_W;  // 32 rows x 64 B, 64-byte swizzle (half staging: 32 B rows)
#if LG_FFN1_HALF_STAGING
    const uint32_t my_row_off = (uint32_t)lane * 32u;
    const uint32_t sw = (uint32_t)((lane >> 2) & 1);
#else
    const uint32_t my_row_off = (uint32_t)lane * 64u;
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#endif
#endif
    const int r_in_tile = quarter * 32 + lane;
    float2* s_local0 = s_stats + 2 * 2 * 128;  // [2 bufs][4 cq][128 rows] partials of this CTA
    int acc = 0, iter = 0;
    uint32_t acc_phase = 0;
#ifdef LG_GEMM_DEBUG
    long long e_wait = 0, e_stats = 0, e_begin = clock64(), tt;
    const bool rec = blockIdx.x == 0 && ew == 0 && lane == 0;
#endif
    for (int mt = m_first; MT_IN(mt); mt += m_step) {
      if (pair_skipped(g, mt)) continue;
      const int row0 = mt * 2 * BM + (int)parity * BM + quarter * 32;
      const bool store_rows = !tile_skipped(g, 2 * mt + (int)parity);  // rows of a fully padded half stay untouched
      const int buf = iter & 1;
      if (ew == 0 && lane == 0) tc::mbar_arrive_expect_tx(&stats_bar[buf], 2 * 128 * 8);  // one partial per row from each CTA
      GT0();
      LG_PI_WAIT(&tfull[acc], acc_phase);
      GT1(e_wait);
      tc::fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2 + cq * 32;
      // ---- pass 1: statistics over this thread's 64 columns (2 blocks x 32)
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t v[32];
        tc::tmem_ld32(t_row + cb * BN, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(v[j]) + s_par[cb * BN + cq * 32 + j];
          sum += x;
          sq = fmaf(x, x, sq);
        }
      }
      // the 4 column slices of a row are summed inside the CTA first, then ONE partial per row goes to both CTAs
      float2* s_local = s_local0 + buf * 4 * 128;
      s_local[cq * 128 + r_in_tile] = make_float2(sum, sq);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      if (cq == 0) {
#pragma unroll
        for (int p = 1; p < 4; ++p) {
          const float2 o = s_local[p * 128 + r_in_tile];
          sum += o.x;
          sq += o.y;
        }
        const uint32_t slot = tc::smem_u32(&s_stats[(buf * 2 + (int)pair) * 128 + r_in_tile]);
        const uint32_t bar = tc::smem_u32(&stats_bar[buf]);
        st_async_f2(map_to_rank(slot, crank), sum, sq, map_to_rank(bar, crank));
        st_async_f2(map_to_rank(slot, partner), sum, sq, map_to_rank(bar, partner));
      }
      GT0();
      LG_PI_WAIT(&stats_bar[buf], (iter >> 1) & 1);
      GT1(e_stats);
      const float2 st0 = s_stats[(buf * 2 + 0) * 128 + r_in_tile], st1 = s_stats[(buf * 2 + 1) * 128 + r_in_tile];
      const float inv_n = 1.f / 512.f;
      const float mean = (st0.x + st1.x) * inv_n;
      const float rstd = rsqrtf(fmaxf((st0.y + st1.y) * inv_n - mean * mean, 0.f) + 1e-5f);
      const float nmr = -mean * rstd;
      // ---- pass 2: normalise, GELU, store
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        uint32_t v[32];
        tc::tmem_ld32(t_row + cb * BN, v);
        tc::tmem_ld_wait();
        if (cb == 1) {  // accumulator fully consumed: release it to the pair's issuer
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (parity == 0) tc::mbar_arrive(&tempty[acc]);
            else mbar_arrive_remote(&tempty[acc], crank - 1);
          }
        }
        const int cw = cb * BN + cq * 32;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = cw + 2 * j;
          const float x0 = __uint_as_float(v[2 * j]) + s_par[c], x1 = __uint_as_float(v[2 * j + 1]) + s_par[c + 1];
          const float a = gelu_act(fmaf(fmaf(x0, rstd, nmr), s_par[BN2 + c], s_par[2 * BN2 + c]));
          const float b = gelu_act(fmaf(fmaf(x1, rstd, nmr), s_par[BN2 + c + 1], s_par[2 * BN2 + c + 1]));
          pk[j] = tc::pack_bf16(a, b);
        }
#if LG_FFN1_DIRECT_STORE
        if (store_rows) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.out16) + (size_t)(row0 + lane) * (2 * BN2) + pair_col0 + cw;
          st_global_256(dst, pk);
          st_global_256(dst + 16, pk + 8);
        }
#elif LG_FFN1_HALF_STAGING
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (lane == 0) bulk_wait_read0();  // previous store has finished reading stg_out
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(stg_out + my_row_off + ((j ^ sw) << 4)) =
                make_uint4(pk[8 * hh + 4 * j], pk[8 * hh + 4 * j + 1], pk[8 * hh + 4 * j + 2], pk[8 * hh + 4 * j + 3]);
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && store_rows) {
            tma_store_2d(&maps.out1, stg_out, pair_col0 + cw + 16 * hh, row0);  // out1: 16-column boxes, 32-byte swizzle
            bulk_commit();
          }
        }
#else
        if (lane == 0) bulk_wait_read0();  // previous store has finished reading stg_out
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(stg_out + my_row_off + ((j ^ sw) << 4)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && store_rows) {
          tma_store_2d(&maps.out1, stg_out, pair_col0 + cw, row0);  // out1: 32-column boxes, 64-byte swizzle
          bulk_commit();
        }
#endif
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      ++iter;
    }
    if (lane == 0) bulk_wait0();
#ifdef LG_GEMM_DEBUG
    if (rec) { g_gemm_times[3] = e_wait; g_gemm_times[4] = clock64() - e_begin; g_gemm_times[6] = e_stats; }
#endif
  }
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc::fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

int make_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  const uint64_t d[2] = {inner, rows}, s[1] = {inner * 2};
  const uint32_t b[2] = {64, box_rows};
  return lg_make_tmap_bf16(m, base, 2, d, s, b);
}

template <int MODE, int CL, bool KBIG>
int launch(const Maps& maps, Args g, int n_blocks, cudaStream_t st) {
  using L = Lay<KBIG>;
  auto kern = tc_ws_linear_kernel<MODE, CL, KBIG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g.n_groups = n_blocks / CL;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = L::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  tc::lg_pdl_attr(&attr[1]);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  // persistent clusters with static striding: never launch more clusters than can be co-resident
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    cfg.gridDim = dim3(sms / CL * CL);
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = sms / CL;
    max_clusters = n;
  }
  int clusters = max_clusters / g.n_groups * g.n_groups;  // whole groups only
  const int need = g.m_tiles * g.n_groups;
  if (clusters > need) clusters = need;
  if (clusters < g.n_groups) clusters = g.n_groups;
  cfg.gridDim = dim3(clusters * CL);
  e = cudaLaunchKernelEx(&cfg, kern, maps, g);
  if (e != cudaSuccess) return (int)e;
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}


template <int MODE, bool KBIG>
int launch_pair_row(const Maps& maps, Args g, int N, cudaStream_t st) {
  using L = Lay<KBIG>;
  auto kern = tc_pair_row_kernel<MODE, KBIG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g.n_groups = N / 256;
  g.m_tiles = g.m_tiles / 2;  // 256-row super-tiles
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(MODE == MODE_HEADS ? 576 : 320);
  cfg.dynamicSmemBytes = L::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  tc::lg_pdl_attr(&attr[1]);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    cfg.gridDim = dim3(sms / 2 * 2);
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = sms / 2;
    max_clusters = n;
  }
  int clusters = max_clusters / g.n_groups * g.n_groups;
  const int need = g.m_tiles * g.n_groups;
  if (clusters > need) clusters = need;
  if (clusters < g.n_groups) clusters = g.n_groups;
  cfg.gridDim = dim3(clusters * 2);
  e = cudaLaunchKernelEx(&cfg, kern, maps, g);
  if (e != cudaSuccess) return (int)e;
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

int launch_pair_ln(const Maps& maps, Args g, cudaStream_t st) {
  auto kern = tc_pair_ln_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LayPL::SMEM);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g.n_groups = 1;
  g.m_tiles = g.m_tiles / 2;  // 256-row super-tiles
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(576);
  cfg.dynamicSmemBytes = LayPL::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  tc::lg_pdl_attr(&attr[1]);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    cfg.gridDim = dim3(sms / 4 * 4);
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = sms / 4;
    max_clusters = n;
  }
  int clusters = max_clusters;
  if (clusters > g.m_tiles) clusters = g.m_tiles;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3(clusters * 4);
  e = cudaLaunchKernelEx(&cfg, kern, maps, g);
  if (e != cudaSuccess) return (int)e;
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

}  // namespace

// Returns LGB200_ERR_SHAPE for combinations this kernel does not cover (the caller falls back to v1).
int lg_tc_linear_v2(int epilogue, const __nv_bfloat16* A0, const __nv_bfloat16* A1, int K0,
                    const __nv_bfloat16* W, int T, int N, int K, const int32_t* lens, LgEpi epi,
                    const void* rot16, const __nv_bfloat16* resid16, cudaStream_t st) {
  if (T % BM || K % BK || K0 % BK || K > 512 || (N != 256 && N != 512 && N != 768)) return LGB200_ERR_SHAPE;
  if (epi.out32 || epi.resid32) return LGB200_ERR_SHAPE;  // bf16 I/O only
  const bool ln = epilogue == LGB200_EPI_LN_GELU;
  const bool kbig = K > 256;
  if (ln && (N != 512 || K != 512 || !epi.out16)) return LGB200_ERR_SHAPE;
  if (kbig && !ln && !(epilogue == LGB200_EPI_ROWMAJOR && N == 256)) return LGB200_ERR_SHAPE;
  if (epilogue == LGB200_EPI_HEADS && epi.n_rot > 0 && !rot16) return LGB200_ERR_SHAPE;
  const int CL = ln ? 4 : 2;
  Maps maps;
  int rc;
  if ((rc = make_map(&maps.a0, A0, K0, T, BM / CL))) return rc;
  if (K0 < K) {
    if ((rc = make_map(&maps.a1, A1, K - K0, T, BM / CL))) return rc;
  } else {
    maps.a1 = maps.a0;
  }
  if ((rc = make_map(&maps.w, W, K, N, BN))) return rc;
  Args g = {};
  g.kb_total = K / BK;
  g.kb_a0 = K0 / BK;
  g.m_tiles = T / BM;
  g.Lp = epi.Lp;
  g.lens = lens;
  g.bias = epi.bias;
  g.gamma = epi.gamma;
  g.beta = epi.beta;
  g.out16 = epi.out16;
  g.scale[0] = epi.scale[0]; g.scale[1] = epi.scale[1]; g.scale[2] = epi.scale[2];
  g.n_rot = epi.n_rot;
  g.has_in = 0;
  static const int pf = getenv("LGB200_GEMM_PREFETCH") ? atoi(getenv("LGB200_GEMM_PREFETCH")) : 1;
  g.prefetch_tiles = pf;
  // walking direction of the M tiles (Args::reverse), bit 0: LayerNorm layer (FFN1), bit 1: HEADS projections, bit 2: the
  // K = 512 ROW layer (FFN2).  Attention writes ctx ascending, so FFN1 walks down (the last-written ctx rows are still in
  // L2), FFN2 walks up (FFN1 wrote the low rows of `hid` last and has just read the low rows of x), the projection of the
  // next block walks down again (FFN2 wrote the high rows of x last).
  static const int rev_mask = getenv("LGB200_GEMM_REVERSE") ? atoi(getenv("LGB200_GEMM_REVERSE")) : 3;
  g.reverse = ln ? (rev_mask & 1) : epilogue == LGB200_EPI_HEADS ? ((rev_mask >> 1) & 1) : kbig ? ((rev_mask >> 2) & 1) : 0;
  maps.in = maps.a0;
  if (epilogue == LGB200_EPI_HEADS) {
    const uint64_t rows = (uint64_t)T * LG_HEADS;  // [S*4*Lp, 64]
    for (int p = 0; p < N / 256; ++p) {
      CUtensorMap* m = p == 0 ? &maps.out0 : (p == 1 ? &maps.out1 : &maps.out2);
      if ((rc = make_map(m, epi.outp[p], 64, rows, 32))) return rc;
    }
    if (N / 256 < 2) maps.out1 = maps.out0;
    if (N / 256 < 3) maps.out2 = maps.out0;
    if (epi.n_rot > 0) {
      if ((rc = make_map(&maps.in, rot16, 64, T, 32))) return rc;  // 32 x (cos, sin) fp16 pairs = 64 halves / row
      g.has_in = 1;
    }
  } else {
    if (!epi.out16) return LGB200_ERR_SHAPE;
    if ((rc = make_map(&maps.out0, epi.out16, N, T, 32))) return rc;
    maps.out1 = maps.out0;
    maps.out2 = maps.out0;
    if (resid16) {
      if ((rc = make_map(&maps.in, resid16, N, T, 32))) return rc;
      g.has_in = 1;
    }
  }
  const int n_blocks = N / BN;
  // bit 1: CTA-pair kernel for the LayerNorm layer (default on: 150 vs 225 us at T = 262144);
  // bit 0: CTA-pair kernel for the K = 512 ROW layer (default off: that layer already runs at 88 % of HBM peak, 93 us both ways)
  // bit 2: CTA-pair kernel for the HEADS layers (QKV, cross projections)
  static const int pair_mode = getenv("LGB200_GEMM_PAIR") ? atoi(getenv("LGB200_GEMM_PAIR")) : 6;
  if (ln && (pair_mode & 2) && T % 256 == 0) {  // CTA-pair kernel: each CTA loads 64-row boxes of A
    Maps pm = maps;
    if ((rc = make_map(&pm.a0, A0, K0, T, 64))) return rc;
    if ((rc = make_map(&pm.a1, A1, K - K0, T, 64))) return rc;
    {  // output boxes of 32 rows x 32 columns, 64-byte swizzle (one per epilogue warp and column block)
      const uint64_t d[2] = {(uint64_t)N, (uint64_t)T}, sb[1] = {(uint64_t)N * 2};
#if LG_FFN1_HALF_STAGING
      const uint32_t bx[2] = {16, 32};
      if ((rc = lg_make_tmap_bf16_sw(&pm.out1, epi.out16, 2, d, sb, bx, 32))) return rc;
#else
      const uint32_t bx[2] = {32, 32};
      if ((rc = lg_make_tmap_bf16_sw(&pm.out1, epi.out16, 2, d, sb, bx, 64))) return rc;
#endif
    }
    return launch_pair_ln(pm, g, st);
  }
  if (ln) return launch<MODE_LN, 4, true>(maps, g, n_blocks, st);
  if (epilogue == LGB200_EPI_HEADS && (pair_mode & 4) && T % 256 == 0 && epi.Lp % 128 == 0) {
    Maps pm = maps;  // CTA-pair kernel: 128-row A boxes, one pair per 256-column group (cluster of 2, no multicast);
                     // outputs and rotary rows in 32-row x 32-column boxes with a 64-byte swizzle (16 epilogue warps)
    if ((rc = make_map(&pm.a0, A0, K0, T, BM))) return rc;
    pm.a1 = pm.a0;
    {
      const uint64_t rows = (uint64_t)T * LG_HEADS;
      const uint64_t d[2] = {64, rows}, sb[1] = {128};
      const uint32_t bx[2] = {32, 32};
      for (int p = 0; p < N / 256; ++p) {
        CUtensorMap* m = p == 0 ? &pm.out0 : (p == 1 ? &pm.out1 : &pm.out2);
        if ((rc = lg_make_tmap_bf16_sw(m, epi.outp[p], 2, d, sb, bx, 64))) return rc;
      }
      if (N / 256 < 2) pm.out1 = pm.out0;
      if (N / 256 < 3) pm.out2 = pm.out0;
      if (epi.n_rot > 0) {
        const uint64_t dr[2] = {64, (uint64_t)T};
        if ((rc = lg_make_tmap_bf16_sw(&pm.in, rot16, 2, dr, sb, bx, 64))) return rc;
      }
    }
    return launch_pair_row<MODE_HEADS, false>(pm, g, N, st);
  }
  if (epilogue == LGB200_EPI_HEADS) return launch<MODE_HEADS, 2, false>(maps, g, n_blocks, st);
  if (kbig && (pair_mode & 1) && T % 256 == 0) {  // CTA-pair kernel: the A boxes are 128 rows (no multicast inside a pair)
    Maps pm = maps;
    if ((rc = make_map(&pm.a0, A0, K0, T, BM))) return rc;
    if (K0 < K) {
      if ((rc = make_map(&pm.a1, A1, K - K0, T, BM))) return rc;
    } else {
      pm.a1 = pm.a0;
    }
    return launch_pair_row<MODE_ROW, true>(pm, g, N, st);
  }
  if (kbig) return launch<MODE_ROW, 2, true>(maps, g, n_blocks, st);
  return launch<MODE_ROW, 2, false>(maps, g, n_blocks, st);
}

extern "C" int lgb200_debug_gemm_times(long long* host_out, int n) {
  if (n > 16) n = 16;
  return (int)cudaMemcpyFromSymbol(host_out, g_gemm_times, sizeof(long long) * n);
}
