// Flash attention forward on tcgen05 / TMEM / TMA (sm_100a), head_dim 64, bf16 in, fp32 accumulate.
//
// PERSISTENT kernel, one CTA per SM, all 512 TMEM columns.  A work item is 256 query rows of one (sequence, head):
// two 128-row sub-tiles A and B that SHARE every K/V tile in shared memory.  Why: with 128 queries per K/V tile
// the kernel moved 32 KB from L2 per 128x128 score tile (4.3 GB per launch at 64 pairs x 2048) and sat at
// ~6 TB/s of L2->SM traffic whatever the softmax or the MMAs did (removing the exponentials: -8 %; removing
// the MMAs: -9 %).  Sharing a K/V tile between two query tiles halves that traffic.
//   warp 0     TMA producer + work scheduler: takes item indices from a global counter (co-resident / neighbouring
//              CTAs do not run at the same speed, a static split was 12 % slower), publishes them in a 4-deep
//              shared-memory ring; Q (2 tiles) per item, K and V tiles of 128 keys into two rings
//   warp 1     MMA issuer: S_X = Q_X.K^T (SS, N=128) into TMEM, O_X += P_X.V (TS: P from TMEM, V MN-major), X in {A,B}
//   warps 2-5  softmax group A, warps 6-9 softmax group B: ONE thread per query row (warp w owns TMEM lanes
//              32(w%4)..), 128 score columns in two 64-column register passes; online max with lazy rescale (O is
//              only touched when the max grows by > 8 in log2 units), ex2 (one in four by polynomial on the FMA
//              pipe), row sum, bf16 P back to TMEM with tcgen05.st; at the end of an item O / l -> ctx
// Issue order per tile g: QK_A(g+1), QK_B(g+1), PV_A(g), PV_B(g): the next score tiles are produced while the
// softmax warps still exponentiate tile g, across item boundaries too.
// The waiting warps (producer, issuer) sleep between polls: a bare try_wait loop took 13 % of the SM's issue slots.
// TMEM columns: S_A [0,128) S_B [128,256) fp32 | P_A [256,320) P_B [320,384) bf16x2 | O_A [384,448) O_B [448,512) fp32.
// Scores arrive in the log2 domain (the Q projection epilogue folds log2(e)/sqrt(64)).
// Keys >= lens[kv sequence] are masked to -inf; query rows >= lens[q sequence] are not stored.
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"
#include <stdlib.h>

// debug timeline (clock64 stamps of CTA 0; read back with lgb200_debug_attn_times)
__device__ long long g_attn_times[2 * 16 * 16 + 256 + 4];

#ifndef LG_ATTN_POLY
#define LG_ATTN_POLY 2  // one exponential in (2 * LG_ATTN_POLY) is evaluated by polynomial on the FMA pipe
#endif

namespace {

constexpr int AT_BM = 128;   // query rows per sub-tile (TMEM lanes)
constexpr int AT_ITEM = 256; // query rows per work item
constexpr int AT_BN = 128;   // keys per step
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int KST = 4, VST = 4;            // K / V ring depth
constexpr int RING = 4;                    // work item ring depth
constexpr int AT_SMEM = TILE_BYTES * (2 + KST + VST) + 512;
constexpr int AT_THREADS = 320;
constexpr int W_PROD = 0, W_MMA = 1, W_SOFT0 = 2;  // issuing warps at the LOW warp ids (lowest arbiter priority)

constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384, TM_COLS = 512;

__device__ __forceinline__ float ex2(float x) {
#ifdef LG_ATTN_X_NOEXP  // experiment: no MUFU at all (results are wrong)
  return x * x;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// 2^x for x <= 8 on the FMA/ALU pipes (the MUFU unit delivers only 16 ex2/clk/SM): round-to-nearest split
// x = n + f, |f| <= 0.5, degree-3 minimax polynomial (max rel. err 1.0e-4, far below the bf16 rounding of P),
// exponent patched in with integer ops.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05500889f, 0.24221097f);
  p = fmaf(p, f, 0.69328294f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

struct Geometry {
  int n_items, QB, Lp, kv_xor;
  const int32_t* lens;
};
struct Item {
  int s, h, q0, nq, nk, n_tiles;
  // false if the item has no valid query row (nothing to do for anybody)
  __device__ __forceinline__ bool decode(const Geometry& G, int idx) {
    const int qb = idx % G.QB, sh = idx / G.QB;
    h = sh % LG_HEADS; s = sh / LG_HEADS; q0 = qb * AT_ITEM;
    nq = G.lens ? G.lens[s] : G.Lp;
    if (q0 >= nq) return false;
    nk = G.lens ? G.lens[s ^ G.kv_xor] : G.Lp;
    n_tiles = (nk + AT_BN - 1) / AT_BN;
    return true;
  }
};
// Reader of the work item ring (whole warp calls it; lane 0 releases the slot when `release` is set).
struct Walker {
  uint32_t n;  // ring entries consumed so far
  Item it;
  __device__ __forceinline__ void init() { n = 0; }
  // moves to the next item with a valid query row (and, for `only_mma`, at least one key tile)
  __device__ __forceinline__ bool next(const Geometry& G, const int* ring, uint64_t* item_full, uint64_t* item_empty,
                                       bool only_mma, bool release, int lane) {
    for (;;) {
      const uint32_t slot = n % RING;
      if (only_mma) tc::mbar_wait_relaxed(&item_full[slot], (n / RING) & 1);
      else tc::mbar_wait(&item_full[slot], (n / RING) & 1);
      const int idx = ring[slot];
      if (release) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&item_empty[slot]);
      }
      ++n;
      if (idx >= G.n_items) return false;
      if (!it.decode(G, idx)) continue;
      if (only_mma && it.n_tiles == 0) continue;
      return true;
    }
  }
};

__global__ void __launch_bounds__(AT_THREADS, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, int S, int Lp, const int32_t* __restrict__ lens,
                    int kv_xor, __nv_bfloat16* __restrict__ ctx, unsigned* __restrict__ counter, int dbg_arg) {
  // The clock64 timeline exists only when the file is compiled with -DLG_ATTN_DEBUG; in the product build every
  // debug branch folds away (run-time debug branches cost ~50 BRA per 64 exponentials in the unrolled loop).
#ifdef LG_ATTN_DEBUG
  const bool tl = (dbg_arg & 16) && blockIdx.x == 0;
#else
  constexpr bool tl = false;
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef LG_ATTN_DEBUG
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // SM clock check: clock64 vs globaltimer over the CTA's life
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_attn_times[768] = clock64();
    g_attn_times[769] = (long long)gt;
  }
#endif

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-byte aligned base
  uint8_t* sQ = smem;                            // 2 tiles: Q_A, Q_B
  uint8_t* sK = smem + 2 * TILE_BYTES;           // KST stages
  uint8_t* sV = smem + (2 + KST) * TILE_BYTES;   // VST stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + KST + VST) * TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;             // [KST]
  uint64_t* k_empty = k_full + KST;        // [KST]
  uint64_t* v_full = k_empty + KST;        // [VST]
  uint64_t* v_empty = v_full + VST;        // [VST]
  uint64_t* s_full = v_empty + VST;        // [2]  (A, B)
  uint64_t* s_free = s_full + 2;           // [2]
  uint64_t* p_ready = s_free + 2;          // [2]
  uint64_t* pv_done = p_ready + 2;         // [2]
  uint64_t* item_full = pv_done + 2;       // [RING]
  uint64_t* item_empty = item_full + RING; // [RING]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(item_empty + RING);
  int* ring = reinterpret_cast<int*>(tmem_slot + 1);  // [RING]
  Geometry G;
  G.QB = (Lp + AT_ITEM - 1) / AT_ITEM; G.n_items = S * LG_HEADS * G.QB; G.Lp = Lp; G.kv_xor = kv_xor; G.lens = lens;

  if (warp == W_PROD && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    tc::mbar_init(q_full, 1);
    tc::mbar_init(q_empty, 1);
    for (int i = 0; i < KST; ++i) { tc::mbar_init(&k_full[i], 1); tc::mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VST; ++i) { tc::mbar_init(&v_full[i], 1); tc::mbar_init(&v_empty[i], 1); }
    for (int x = 0; x < 2; ++x) {
      tc::mbar_init(&s_full[x], 1);
      tc::mbar_init(&s_free[x], 4);
      tc::mbar_init(&p_ready[x], 4);
      tc::mbar_init(&pv_done[x], 1);
    }
    for (int i = 0; i < RING; ++i) { tc::mbar_init(&item_full[i], 1); tc::mbar_init(&item_empty[i], 9); }
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, TM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_PROD) {
    if (lane == 0) {
      Item w;
      uint32_t kc = 0, vc = 0, ic = 0;  // K tiles, V tiles, items issued so far by this CTA
      for (uint32_t n = 0;; ++n) {
        const uint32_t slot = n % RING;
        tc::mbar_wait_relaxed(&item_empty[slot], ((n / RING) & 1) ^ 1);
        const int idx = (int)atomicAdd(counter, 1u);
        ring[slot] = idx;
        tc::mbar_arrive(&item_full[slot]);
        if (idx >= G.n_items) break;
        if (!w.decode(G, idx) || w.n_tiles == 0) continue;
        const int qrow = (w.s * LG_HEADS + w.h) * Lp + w.q0;
        const int kvrow = ((w.s ^ kv_xor) * LG_HEADS + w.h) * Lp;
        tc::mbar_wait_relaxed(q_empty, (ic & 1) ^ 1);  // the previous item's last QK^T retired
        tc::mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
        tc::tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
        // (sub-tile B of the last, odd item of a sequence reads the next head's rows or the zero fill past the
        //  end of the tensor; its rows are >= nq and are never stored)
        tc::tma_load_2d(sQ + TILE_BYTES, &tmQ, q_full, 0, qrow + AT_BM);
        ++ic;
        for (int j = 0; j < w.n_tiles; ++j) {
          const uint32_t ks = kc % KST, vs = vc % VST;
          tc::mbar_wait_relaxed(&k_empty[ks], ((kc / KST) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
          tc::tma_load_2d(sK + ks * TILE_BYTES, &tmK, &k_full[ks], 0, kvrow + j * AT_BN);
          ++kc;
          tc::mbar_wait_relaxed(&v_empty[vs], ((vc / VST) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
          tc::tma_load_2d(sV + vs * TILE_BYTES, &tmV, &v_full[vs], 0, kvrow + j * AT_BN);
          ++vc;
        }
      }
    }
  } else if (warp == W_MMA) {
    // MMA issuer.  The WHOLE warp runs this loop with warp-uniform control flow and one elected lane
    // issues: descriptors and barrier addresses then live in uniform registers.  (Running the loop on lane 0
    // only needed R2UR moves for every tcgen05.mma and cost ~100 issue cycles each.)
    constexpr uint32_t idesc_qk = tc::idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = tc::idesc_bf16(128, 64, 1);
    const uint64_t dQ0 = tc::smem_desc_sw128(tc::smem_u32(sQ), 0, 1024);
    const uint64_t dK0 = tc::smem_desc_sw128(tc::smem_u32(sK), 0, 1024);
    const uint64_t dV0 = tc::smem_desc_sw128(tc::smem_u32(sV), TILE_BYTES, 1024);
    // two walkers over the same tile sequence: `a` (QK^T) runs one tile ahead of `b` (P.V)
    // (`a` only peeks at the ring; `b`, the later of the two, releases the slots)
    Walker a, b;
    a.init();
    b.init();
    int aj = 0, bj = 0;
    bool a_ok = a.next(G, ring, item_full, item_empty, true, false, lane);
    bool b_ok = b.next(G, ring, item_full, item_empty, true, true, lane);
    uint32_t gq = 0, gp = 0, ia = 0;  // QK^T / P.V tiles issued, items started by the QK walker
    uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
    auto issue_qk = [&]() {  // S_A, S_B = Q_A, Q_B . K[ks]^T for tile (a, aj)
      if (aj == 0) { tc::mbar_wait_relaxed(q_full, ia & 1); ++ia; }
      tc::mbar_wait_relaxed(&k_full[ks], kph);
      const uint64_t dK = dK0 + (uint64_t)(ks * (TILE_BYTES >> 4));
      const bool last = aj + 1 == a.it.n_tiles;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (gq > 0) tc::mbar_wait_relaxed(&s_free[x], (gq - 1) & 1);  // group x holds S_x(gq-1) in registers
        tc::fence_after_sync();
        const uint64_t dQ = dQ0 + (uint64_t)(x * (TILE_BYTES >> 4));
        if (tc::elect_one()) {
#ifndef LG_ATTN_X_NOMMA  // experiment: barriers only, no tensor work
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_ss(tmem + TM_S + x * 128, dQ + 2 * k, dK + 2 * k, idesc_qk, k != 0);
#endif
          tc::umma_commit(&s_full[x]);
          if (x == 1) {
            tc::umma_commit(&k_empty[ks]);  // K stage free once both QK^T retire
            if (last) tc::umma_commit(q_empty);
          }
        }
        __syncwarp();
      }
      if (++ks == KST) { ks = 0; kph ^= 1; }
      ++gq;
      if (last) { aj = 0; a_ok = a.next(G, ring, item_full, item_empty, true, false, lane); } else ++aj;
    };
    if (a_ok) issue_qk();
#define MSTAMP(k) do { if (tl && lane == 0 && gp < 16) g_attn_times[256 + (gp * 16) + (k)] = clock64(); } while (0)
    while (b_ok) {
      MSTAMP(0);
      if (a_ok) issue_qk();
      MSTAMP(3);
      tc::mbar_wait_relaxed(&v_full[vs], vph);
      const uint64_t dV = dV0 + (uint64_t)(vs * (TILE_BYTES >> 4));
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        tc::mbar_wait_relaxed(&p_ready[x], gp & 1);
        tc::fence_after_sync();
        MSTAMP(4 + x);
        if (tc::elect_one()) {
#ifndef LG_ATTN_X_NOMMA
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k)  // 16 keys per MMA: P columns k*8.., V rows k*16.. (2048 B)
            tc::umma_ts(tmem + TM_O + x * 64, tmem + TM_P + x * 64 + k * 8, dV + k * (2048 >> 4), idesc_pv, (bj | k) != 0);
#endif
          if (x == 1) tc::umma_commit(&v_empty[vs]);
          tc::umma_commit(&pv_done[x]);
        }
        __syncwarp();
      }
      MSTAMP(6);
      if (++vs == VST) { vs = 0; vph ^= 1; }
      ++gp;
      if (bj + 1 == b.it.n_tiles) { bj = 0; b_ok = b.next(G, ring, item_full, item_empty, true, true, lane); } else ++bj;
    }
  } else {
    // softmax: group x = 0 (A, warps 2-5) / 1 (B, warps 6-9); ONE thread per query row, warp w owns TMEM lanes
    // 32(w%4)..  The kernel is bound by instruction issue, not MUFU alone (1 warp-instruction per clock per SM
    // sub-partition; tools/micro/pipe_rate.cu); thin warps (two threads per row) needed a max exchange through
    // shared memory and twice the barrier traffic for the same arithmetic.
    //   pass 1: columns 0-63 -> max;  columns 64-127 -> max, 2^x, pack (they stay in registers)
    //   pass 2: columns 0-63 loaded again from TMEM (cheaper than 64 more live registers) -> 2^x, pack
    const int x = (warp - W_SOFT0) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem + lane_base + TM_S + x * 128, tP = tmem + lane_base + TM_P + x * 64,
                   tO = tmem + lane_base + TM_O + x * 64;
    uint64_t* my_s_full = &s_full[x];
    uint64_t* my_s_free = &s_free[x];
    uint64_t* my_p_ready = &p_ready[x];
    uint64_t* my_pv_done = &pv_done[x];
    const bool rec = tl && warp == W_SOFT0 && lane == 0;
#define STAMP(k) do { if (rec && g < 16) g_attn_times[(g * 16) + (k)] = clock64(); } while (0)
    Walker wk;
    wk.init();
    uint32_t g = 0;  // tiles consumed so far by this CTA (all items)
    while (wk.next(G, ring, item_full, item_empty, false, true, lane)) {
      const Item& w = wk.it;
      const int row = w.q0 + x * AT_BM + r;  // query row within the sequence
      __nv_bfloat16* out_row = ctx + ((size_t)w.s * Lp + row) * LG_D + w.h * LG_DH;
      if (w.n_tiles == 0) {  // no keys: attention output is defined as zero (nan_to_num)
        if (row < w.nq) {
          uint4* dst = reinterpret_cast<uint4*>(out_row);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < w.n_tiles; ++j, ++g) {
        STAMP(0);
        if (rec && g < 128) g_attn_times[512 + 2 * g] = clock64();
        tc::mbar_wait(my_s_full, g & 1);
        tc::fence_after_sync();
        STAMP(1);
        const int valid = w.nk - j * AT_BN;  // valid keys among the 128 columns of this tile
        uint32_t sv[64];
        float mxs[4];
        // ---- pass 1a: columns 0..63, maximum only
        tc::tmem_ld32(tS, sv);
        tc::tmem_ld32(tS + 32, sv + 32);
        tc::tmem_ld_wait();
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (i >= valid) sv[i] = 0xff800000u;  // -inf
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = __uint_as_float(sv[i]);
#pragma unroll
        for (int i = 4; i < 64; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sv[i]));
        // ---- pass 1b: columns 64..127
        tc::tmem_ld32(tS + 64, sv);
        tc::tmem_ld32(tS + 96, sv + 32);
        tc::tmem_ld_wait();
        STAMP(2);
        if (valid < 128) {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (i + 64 >= valid) sv[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sv[i]));
        const float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
        STAMP(4);
        // lazy rescale: keep the reference max unless it grows by more than 8 (factor 256)
        float m_new = m_ref;
        if (mx > m_ref + 8.f) m_new = mx;
        const float alpha = ex2(m_ref - m_new);  // 1 when unchanged, 0 on the first tile
        float rsum[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pkb[32], pka[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float p0 = __uint_as_float(sv[2 * i]) - m_new, p1 = __uint_as_float(sv[2 * i + 1]) - m_new;
          p0 = ex2(p0);
#if LG_ATTN_POLY > 0
          p1 = (i % LG_ATTN_POLY == 0) ? ex2_poly(p1) : ex2(p1);
#else
          p1 = ex2(p1);
#endif
          rsum[i & 3] += p0 + p1;
          pkb[i] = tc::pack_bf16(p0, p1);
        }
        // ---- pass 2: columns 0..63 again
        tc::tmem_ld32(tS, sv);
        tc::tmem_ld32(tS + 32, sv + 32);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(my_s_free);  // S is consumed: the next QK^T may overwrite it
        STAMP(3);
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (i >= valid) sv[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float p0 = __uint_as_float(sv[2 * i]) - m_new, p1 = __uint_as_float(sv[2 * i + 1]) - m_new;
          p0 = ex2(p0);
#if LG_ATTN_POLY > 0
          p1 = (i % LG_ATTN_POLY == 0) ? ex2_poly(p1) : ex2(p1);
#else
          p1 = ex2(p1);
#endif
          rsum[i & 3] += p0 + p1;
          pka[i] = tc::pack_bf16(p0, p1);
        }
        l_run = l_run * alpha + ((rsum[0] + rsum[1]) + (rsum[2] + rsum[3]));
        STAMP(5);
        if (g > 0) {
          tc::mbar_wait(my_pv_done, (g - 1) & 1);  // PV(g-1) retired: P is free, O is up to date
          tc::fence_after_sync();
        }
        STAMP(6);
        if (j > 0) {
          const bool need = m_new != m_ref;
          if (__any_sync(0xffffffffu, need)) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[32];
              tc::tmem_ld32(tO + hh * 32, o);
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tc::tmem_st32(tO + hh * 32, o);
            }
          }
        }
        m_ref = m_new;
        tc::tmem_st32(tP, pka);
        tc::tmem_st32(tP + 32, pkb);
        tc::tmem_st_wait();
        STAMP(7);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(my_p_ready);
        STAMP(8);
        if (rec && g < 128) g_attn_times[512 + 2 * g + 1] = clock64();
      }
      // end of item: normalise the 64 output columns of this row
      tc::mbar_wait(my_pv_done, (g - 1) & 1);
      tc::fence_after_sync();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      // (the next item's first P.V overwrites O only after every warp of the group arrived on p_ready again,
      //  i.e. after these reads)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tc::tmem_ld32(tO + hh * 32, o);
        tc::tmem_ld_wait();
        if (row < w.nq) {
          uint4* dst = reinterpret_cast<uint4*>(out_row + hh * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 q;
            q.x = tc::pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            q.y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            q.z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            q.w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
            dst[i] = q;
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
#ifdef LG_ATTN_DEBUG
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_attn_times[770] = clock64();
    g_attn_times[771] = (long long)gt;
  }
#endif
  // the last CTA to finish re-arms the work counter for the next launch that uses this slot
  if (threadIdx.x == 0 && atomicAdd(counter + 1, 1u) == gridDim.x - 1) {
    counter[0] = 0;
    counter[1] = 0;
  }
}

}  // namespace

int lg_tc_attention(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                    const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, cudaStream_t st) {
  static const int dbg = getenv("LGB200_ATTN_DBG") ? atoi(getenv("LGB200_ATTN_DBG")) : 0;
  cudaError_t e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 16) return LGB200_ERR_SHAPE;
  static int n_sm[16] = {0};
  if (n_sm[dev] == 0) {
    if ((e = cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
  }
  if (Lp % AT_BM != 0) return LGB200_ERR_SHAPE;
  CUtensorMap tq, tk, tv;
  const uint64_t d[2] = {64, (uint64_t)S * LG_HEADS * Lp}, sb[1] = {128};
  const uint32_t box[2] = {64, 128};
  int rc;
  if ((rc = lg_make_tmap_bf16(&tq, Q, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tk, K, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tv, V, 2, d, sb, box))) return rc;
  e = cudaFuncSetAttribute(tc_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  if (e != cudaSuccess) return (int)e;
  const int n_items = S * LG_HEADS * ((Lp + AT_ITEM - 1) / AT_ITEM);
  if (n_items <= 0) return LGB200_OK;
  const int grid = n_items < n_sm[dev] ? n_items : n_sm[dev];
  // work counters {next item, finished CTAs}: a small pool per device so that launches in flight on different
  // streams do not share one; the kernel re-arms its slot when it finishes.  One-time allocation per device.
  static unsigned* pool[16] = {nullptr};
  static unsigned seq = 0;
  if (!pool[dev]) {
    if ((e = cudaMalloc(&pool[dev], 64 * 2 * sizeof(unsigned))) != cudaSuccess) return (int)e;
    if ((e = cudaMemset(pool[dev], 0, 64 * 2 * sizeof(unsigned))) != cudaSuccess) return (int)e;
  }
  unsigned* counter = pool[dev] + 2 * (seq++ % 64);
  tc_attention_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(tq, tk, tv, S, Lp, lens, kv_xor, ctx, counter, dbg);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

extern "C" int lgb200_debug_attn_times(long long* host_out, int n) {
  if (n > 772) n = 772;
  return (int)cudaMemcpyFromSymbol(host_out, g_attn_times, sizeof(long long) * n);
}
