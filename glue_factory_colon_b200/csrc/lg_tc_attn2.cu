// Flash attention forward, two-query-tile variant (sm_100a), OPT-IN (LGB200_ATTN2=1): one CTA per SM works on 256
// query rows of one (sequence, head) -- two 128-row tiles A and B that share every K/V tile in shared memory (half the
// L2->SM traffic of lg_tc_attn.cu) -- with all 512 TMEM columns and two softmax groups of 8 warps.
//
//   warp 0      TMA producer: Q (2 tiles) once, then K and V tiles of 128 keys into two rings
//   warp 1      MMA issuer, per step i:  PV_A(i-1), QK_A(i+1), PV_B(i-1), QK_B(i+1)
//   warps 2-9   softmax group A, warps 10-17 group B: quad TMEM layout (16x256b loads: a warp owns 16 lanes and all
//               128 key columns, a row's columns live in one quad, row maximum = two shuffles; 16x128b stores for P;
//               no shared-memory exchange, no named barriers)
// Scores arrive relative to the row's reference maximum: a fifth K=16 MMA slice multiplies Q_ext (-m_ref as hi + lo
// bf16, rewritten by the softmax threads when the reference moves) with a constant K_ext of ones.
// TMEM columns: S_A [0,128) S_B [128,256) fp32 | P_A [256,320) P_B [320,384) bf16x2 | O_A [384,448) O_B [448,512) fp32.
//
// Measured at S=128, Lp=2048 in one run: shipped kernel 0.671 ms; this one free-running 0.667, with group B delayed once
// by 400 ns 0.661, with STRICTLY ordered exponential phases (group B exponentiates only after group A and vice versa,
// LG_ATTN2_ORDERED=1) 0.710 ms.  Parity-green incl. ragged batches.  Every structure tried this round lands at
// 0.66-0.68 ms with MUFU ~60 % busy; this file is the base for the next attempt (registers for a software-pipelined
// softmax are available here: 113 per thread at 1 CTA/SM).
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"
#include <stdlib.h>

#ifndef LG_ATTN2_POLY16
#define LG_ATTN2_POLY16 3  // of every 16 exponentials, this many are evaluated by polynomial on the FMA pipe
#endif
#define LG2_POLY_HERE(i) ((((i) * LG_ATTN2_POLY16) % 8) < LG_ATTN2_POLY16)
#ifndef LG_ATTN2_ORDERED
#define LG_ATTN2_ORDERED 0  // 1: strict alternation of the two groups' exponential phases (measured slower: 0.710 vs 0.670 ms)
#endif
#ifndef LG_ATTN2_STAGGER_NS
#define LG_ATTN2_STAGGER_NS 400
#endif

namespace {

constexpr int BM = 128, BN = 128;
constexpr int TILE_BYTES = 128 * 64 * 2;
constexpr int XT_BYTES = 128 * 16 * 2;
constexpr int KST = 4, VST = 3;
constexpr int OFF_K = 2 * TILE_BYTES;
constexpr int OFF_V = OFF_K + KST * TILE_BYTES;
constexpr int OFF_X = OFF_V + VST * TILE_BYTES;  // Q_ext A | Q_ext B | K_ext
constexpr int OFF_BAR = OFF_X + 3 * XT_BYTES;
constexpr int A2_NBAR = 32;                 // mbarriers per pass (25 used)
constexpr int A2_SMEM = OFF_BAR + 2 * A2_NBAR * 8 + 64;  // two barrier sets (pass 0 / exact-mode restart) + tmem slot + panic flag
constexpr int A2_THREADS = 18 * 32;
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_poly(float x) {  // valid for x < 127.5 (the deferred mode checks the inputs' maximum)
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05500889f, 0.24221097f);
  p = fmaf(p, f, 0.69328294f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__global__ void __launch_bounds__(A2_THREADS, 1)
tc_attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, int Lp, const int32_t* __restrict__ lens, int kv_xor,
                     __nv_bfloat16* __restrict__ ctx, int start_mode) {
  const int s = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 2 * BM;
  const int nq = lens ? lens[s] : Lp;
  if (q0 >= nq) return;
  const int skv = s ^ kv_xor;
  const int nk = lens ? lens[skv] : Lp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (nk + BN - 1) / BN;
  if (n_tiles == 0) {  // no keys: the output is defined as zero
    if (warp >= 2 && warp < 10) {
      const int r = (warp - 2) * 32 + lane;  // 0..255
      if (q0 + r < nq) {
        uint4* dst = reinterpret_cast<uint4*>(ctx + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;  // 2 tiles
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sQx = smem + OFF_X;             // [2][128][16] bf16
  uint8_t* sKx = sQx + 2 * XT_BYTES;       // [128][16] bf16
  uint64_t* bars_base = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_base + 2 * A2_NBAR);
  volatile int* panic = reinterpret_cast<volatile int*>(tmem_slot + 1);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    for (int set = 0; set < 2; ++set) {  // second set: the exact-mode restart (see `panic`) starts on fresh barriers
      uint64_t* b = bars_base + set * A2_NBAR;
      tc::mbar_init(b, 1);
      for (int i = 0; i < KST; ++i) { tc::mbar_init(b + 1 + i, 1); tc::mbar_init(b + 1 + KST + i, 1); }
      for (int i = 0; i < VST; ++i) { tc::mbar_init(b + 1 + 2 * KST + i, 1); tc::mbar_init(b + 1 + 2 * KST + VST + i, 1); }
      uint64_t* g = b + 1 + 2 * KST + 2 * VST;
      for (int x = 0; x < 2; ++x) {
        tc::mbar_init(g + x, 1);      // s_full
        tc::mbar_init(g + 2 + x, 8);  // s_free
        tc::mbar_init(g + 4 + x, 8);  // p_ready
        tc::mbar_init(g + 6 + x, 1);  // pv_done
        tc::mbar_init(g + 8 + x, 8);  // turn
      }
    }
    *panic = 0;
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint4*>(sQx)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x < 256) reinterpret_cast<uint4*>(sKx)[threadIdx.x] = make_uint4(0x3f803f80u, 0u, 0u, 0u);
  tc::fence_proxy_async();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // Pass 0: deferred row maximum (mode 0, see lg_tc_attn.cu): after the first tile the exponentials run against the
  // reference the tile was produced with; the quad-reduced row sum of a tile, formed in the tile's tail, moves the
  // reference at the next tile when it exceeds 2^24.  Above 2^70 / inf / NaN (or a polynomial-lane input above 126)
  // the CTA sets `panic`, finishes its schedule and repeats the work item in the exact mode (mode 1) on fresh barriers.
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
  const int mode = start_mode | pass;
  uint64_t* bars = bars_base + pass * A2_NBAR;
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [KST]
  uint64_t* k_empty = k_full + KST;        // [KST]
  uint64_t* v_full = k_empty + KST;        // [VST]
  uint64_t* v_empty = v_full + VST;        // [VST]
  uint64_t* s_full = v_empty + VST;        // [2]
  uint64_t* s_free = s_full + 2;           // [2]
  uint64_t* p_ready = s_free + 2;          // [2]
  uint64_t* pv_done = p_ready + 2;         // [2]
  uint64_t* turn = pv_done + 2;            // [2] turn[x]: group x may exponentiate
  (void)turn;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = (s * LG_HEADS + h) * Lp + q0;
      const int kvrow = (skv * LG_HEADS + h) * Lp;
      tc::mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
      tc::tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
      // (tile B of the last, odd 256-row block of a sequence reads the next head's rows or the zero fill past the
      //  tensor; its rows are >= nq and never stored)
      tc::tma_load_2d(sQ + TILE_BYTES, &tmQ, q_full, 0, qrow + BM);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % KST, vs = j % VST;
        tc::mbar_wait_relaxed(&k_empty[ks], ((j / KST) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
        tc::tma_load_2d(sK + ks * TILE_BYTES, &tmK, &k_full[ks], 0, kvrow + j * BN);
        tc::mbar_wait_relaxed(&v_empty[vs], ((j / VST) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
        tc::tma_load_2d(sV + vs * TILE_BYTES, &tmV, &v_full[vs], 0, kvrow + j * BN);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = tc::idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = tc::idesc_bf16(128, 64, 1);
    const uint64_t dQ0 = tc::smem_desc_sw128(tc::smem_u32(sQ), 0, 1024);
    const uint64_t dK0 = tc::smem_desc_sw128(tc::smem_u32(sK), 0, 1024);
    const uint64_t dV0 = tc::smem_desc_sw128(tc::smem_u32(sV), TILE_BYTES, 1024);
    const uint64_t dQx0 = tc::smem_desc_sw32(tc::smem_u32(sQx), 256);
    const uint64_t dKx = tc::smem_desc_sw32(tc::smem_u32(sKx), 256);
    // S_x(t) = Q_x . K(t)^T (+ the reference slice); the K stage is released after group B's product
    auto issue_qk = [&](int x, int t) {
      const int ks = t % KST;
      if (x == 0) tc::mbar_wait_relaxed(&k_full[ks], (t / KST) & 1);
      if (t > 0) tc::mbar_wait_relaxed(&s_free[x], (t - 1) & 1);  // group x holds S_x(t-1) in registers
      tc::fence_after_sync();
      const uint64_t dQ = dQ0 + (uint64_t)(x * (TILE_BYTES >> 4));
      const uint64_t dK = dK0 + (uint64_t)(ks * (TILE_BYTES >> 4));
      const uint64_t dQx = dQx0 + (uint64_t)(x * (XT_BYTES >> 4));
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_ss(tmem + TM_S + x * 128, dQ + 2 * k, dK + 2 * k, idesc_qk, k != 0);
        tc::umma_ss(tmem + TM_S + x * 128, dQx, dKx, idesc_qk, 1);
        tc::umma_commit(&s_full[x]);
        if (x == 1) tc::umma_commit(&k_empty[ks]);
      }
      __syncwarp();
    };
    // O_x += P_x(t) . V(t); the V stage is released after group B's product
    auto issue_pv = [&](int x, int t) {
      const int vs = t % VST;
      if (x == 0) tc::mbar_wait_relaxed(&v_full[vs], (t / VST) & 1);
      tc::mbar_wait_relaxed(&p_ready[x], t & 1);
      tc::fence_after_sync();
      const uint64_t dV = dV0 + (uint64_t)(vs * (TILE_BYTES >> 4));
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          tc::umma_ts(tmem + TM_O + x * 64, tmem + TM_P + x * 64 + k * 8, dV + k * (2048 >> 4), idesc_pv, (t | k) != 0);
        if (x == 1) tc::umma_commit(&v_empty[vs]);
        tc::umma_commit(&pv_done[x]);
      }
      __syncwarp();
    };
    tc::mbar_wait_relaxed(q_full, 0);
    issue_qk(0, 0);
    issue_qk(1, 0);
    for (int i = 0; i <= n_tiles; ++i) {
      if (i >= 1) issue_pv(0, i - 1);
      if (i + 1 < n_tiles) issue_qk(0, i + 1);
      if (i >= 1) issue_pv(1, i - 1);
      if (i + 1 < n_tiles) issue_qk(1, i + 1);
    }
  } else {
    const int x = (warp - 2) >> 3;          // softmax group: 0 = tile A, 1 = tile B
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int sub = ((warp - 2) >> 2) & 1;  // 16-lane half of the quarter
    const int c4 = lane & 3;
    const int r_lo = quarter * 32 + sub * 16 + (lane >> 2), r_hi = r_lo + 8;  // rows inside the 128-row tile
    const uint32_t lane_base = (uint32_t)(quarter * 32 + sub * 16) << 16;
    const uint32_t tS = tmem + lane_base + TM_S + x * 128, tP = tmem + lane_base + TM_P + x * 64,
                   tO = tmem + lane_base + TM_O + x * 64;
    uint8_t* myQx = sQx + x * XT_BYTES;
    uint64_t* my_s_full = &s_full[x];
    uint64_t* my_s_free = &s_free[x];
    uint64_t* my_p_ready = &p_ready[x];
    uint64_t* my_pv_done = &pv_done[x];
    float m_lo = 0.f, m_hi = 0.f, l_lo = 0.f, l_hi = 0.f;
    float jn_lo = 0.f, jn_hi = 0.f;  // deferred mode: the rows' sums over the previous tile (whole quad)
    bool bad = false, mvn_lo = false, mvn_hi = false, any_next = false;
    for (int j = 0; j < n_tiles; ++j) {
      bool mv_lo = mvn_lo, mv_hi = mvn_hi, any_move = any_next;  // deferred mode: decided in the previous tile's tail
      tc::mbar_wait(my_s_full, j & 1);
      tc::fence_after_sync();
      uint32_t sv[64];  // [32h + 4k + e]: key column 64h + 8k + 2*c4 + (e & 1), row (e < 2 ? lo : hi)
      tc::tmem_ld_16x256b_x8(tS, sv);
      tc::tmem_ld_16x256b_x8(tS + 64, sv + 32);
      tc::tmem_ld_wait();
      const int valid = nk - j * BN;
      if (valid < BN) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const int col = (i >> 5) * 64 + ((i >> 2) & 7) * 8 + 2 * c4 + (i & 1);
          if (col >= valid) sv[i] = 0xff800000u;  // -inf
        }
      }
      const bool exact_tile = mode != 0 || j == 0;  // uniform over the CTA
      float up_lo = 0.f, up_hi = 0.f;
      if (exact_tile) {
        float a0 = __uint_as_float(sv[0]), a1 = __uint_as_float(sv[1]), b0 = __uint_as_float(sv[2]), b1 = __uint_as_float(sv[3]);
#pragma unroll
        for (int q = 1; q < 16; ++q) {
          a0 = fmaxf(a0, __uint_as_float(sv[4 * q]));
          a1 = fmaxf(a1, __uint_as_float(sv[4 * q + 1]));
          b0 = fmaxf(b0, __uint_as_float(sv[4 * q + 2]));
          b1 = fmaxf(b1, __uint_as_float(sv[4 * q + 3]));
        }
        float mx_lo = fmaxf(a0, a1), mx_hi = fmaxf(b0, b1);
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        mv_lo = j == 0 || mx_lo > 8.f;
        mv_hi = j == 0 || mx_hi > 8.f;
        up_lo = mx_lo;
        up_hi = mx_hi;
        any_move = __any_sync(0xffffffffu, mv_lo || mv_hi);
      }
      float d_lo = 0.f, d_hi = 0.f, al_lo = 1.f, al_hi = 1.f;
      auto new_ref = [&](float& m, float up, int row, bool writer) -> float {
        const float base = m;  // (0 before the first tile)
        const float want = base + up;
        const __nv_bfloat16 hi = __float2bfloat16_rn(want);
        const __nv_bfloat16 lo = __float2bfloat16_rn(want - __bfloat162float(hi));
        const float m_abs = __bfloat162float(hi) + __bfloat162float(lo);
        m = m_abs;
        if (writer) {
          const uint32_t bits = ((uint32_t)(*reinterpret_cast<const unsigned short*>(&hi)) |
                                 ((uint32_t)(*reinterpret_cast<const unsigned short*>(&lo)) << 16)) ^ 0x80008000u;
          *reinterpret_cast<uint32_t*>(myQx + row * 32) = bits;
        }
        return m_abs - base;
      };
      if (any_move) {  // rare after the first tile
        if (!exact_tile) {  // the previous tile's row sum, taken against the current reference, says by how much
          if (mv_lo) { if (!(jn_lo <= 0x1p70f)) { bad = true; mv_lo = false; } else up_lo = (float)((int)(__float_as_uint(jn_lo) >> 23) - 127); }
          if (mv_hi) { if (!(jn_hi <= 0x1p70f)) { bad = true; mv_hi = false; } else up_hi = (float)((int)(__float_as_uint(jn_hi) >> 23) - 127); }
        }
        if (mv_lo) d_lo = new_ref(m_lo, up_lo, r_lo, c4 == 0);
        if (mv_hi) d_hi = new_ref(m_hi, up_hi, r_hi, c4 == 1);
        tc::fence_proxy_async();
        al_lo = j == 0 ? 0.f : ex2(-d_lo);
        al_hi = j == 0 ? 0.f : ex2(-d_hi);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(my_s_free);  // S is in registers, Q_ext is up to date: QK^T(j+1) may go
#if LG_ATTN2_ORDERED
      // my turn to exponentiate?  A(j) follows B(j-1), B(j) follows A(j)
      if (x == 0) { if (j > 0) tc::mbar_wait(&turn[0], (j - 1) & 1); }
      else tc::mbar_wait(&turn[1], j & 1);
#else
      // free-running groups, started half a step apart: group B delays its first exponentials once
      if (LG_ATTN2_STAGGER_NS > 0 && x == 1 && j == 0) __nanosleep(LG_ATTN2_STAGGER_NS);
#endif
      uint32_t pk[32];
      float2 s_lo = make_float2(0.f, 0.f), s_hi = make_float2(0.f, 0.f);
      float pmax = -INFINITY;  // largest input of a polynomial lane
      if (any_move) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float p0 = ex2(__uint_as_float(sv[4 * q]) - d_lo), p1 = __uint_as_float(sv[4 * q + 1]) - d_lo;
          float p2 = ex2(__uint_as_float(sv[4 * q + 2]) - d_hi), p3 = __uint_as_float(sv[4 * q + 3]) - d_hi;
          if (LG2_POLY_HERE(2 * q)) pmax = fmaxf(pmax, p1);
          if (LG2_POLY_HERE(2 * q + 1)) pmax = fmaxf(pmax, p3);
          p1 = LG2_POLY_HERE(2 * q) ? ex2_poly(p1) : ex2(p1);
          p3 = LG2_POLY_HERE(2 * q + 1) ? ex2_poly(p3) : ex2(p3);
          s_lo = __fadd2_rn(s_lo, make_float2(p0, p1));
          s_hi = __fadd2_rn(s_hi, make_float2(p2, p3));
          pk[2 * q] = tc::pack_bf16(p0, p1);
          pk[2 * q + 1] = tc::pack_bf16(p2, p3);
        }
      } else {
        float pend = -INFINITY;
        bool have_pend = false;  // (compile-time after unrolling)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float p0 = ex2(__uint_as_float(sv[4 * q])), p1 = __uint_as_float(sv[4 * q + 1]);
          float p2 = ex2(__uint_as_float(sv[4 * q + 2])), p3 = __uint_as_float(sv[4 * q + 3]);
          if (LG2_POLY_HERE(2 * q)) {
            if (have_pend) { pmax = max3(pmax, pend, p1); have_pend = false; } else { pend = p1; have_pend = true; }
          }
          if (LG2_POLY_HERE(2 * q + 1)) {
            if (have_pend) { pmax = max3(pmax, pend, p3); have_pend = false; } else { pend = p3; have_pend = true; }
          }
          p1 = LG2_POLY_HERE(2 * q) ? ex2_poly(p1) : ex2(p1);
          p3 = LG2_POLY_HERE(2 * q + 1) ? ex2_poly(p3) : ex2(p3);
          s_lo = __fadd2_rn(s_lo, make_float2(p0, p1));
          s_hi = __fadd2_rn(s_hi, make_float2(p2, p3));
          pk[2 * q] = tc::pack_bf16(p0, p1);
          pk[2 * q + 1] = tc::pack_bf16(p2, p3);
        }
        if (have_pend) pmax = fmaxf(pmax, pend);
      }
      bad = bad || !(pmax <= 126.f);
#if LG_ATTN2_ORDERED
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&turn[x ^ 1]);  // the other group may exponentiate now
#endif
      const float t_lo = s_lo.x + s_lo.y, t_hi = s_hi.x + s_hi.y;
      l_lo = l_lo * al_lo + t_lo;
      l_hi = l_hi * al_hi + t_hi;
      if (j > 0) {
        tc::mbar_wait(my_pv_done, (j - 1) & 1);  // PV(j-1) retired: P is free, O is up to date
        tc::fence_after_sync();
        if (any_move) {
          uint32_t o[32];
          tc::tmem_ld_16x256b_x8(tO, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * ((i & 2) ? al_hi : al_lo));
          tc::tmem_st_16x256b_x8(tO, o);
        }
      }
      tc::tmem_st_16x128b_x16(tP, pk);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(my_p_ready);
      // tail (overlaps the wait for the next score tile): the rows' sums over this tile, identical in the four
      // threads of a quad (butterfly of commutative adds), decide whether the next tile moves the reference
      jn_lo = t_lo + __shfl_xor_sync(0xffffffffu, t_lo, 1);
      jn_hi = t_hi + __shfl_xor_sync(0xffffffffu, t_hi, 1);
      jn_lo += __shfl_xor_sync(0xffffffffu, jn_lo, 2);
      jn_hi += __shfl_xor_sync(0xffffffffu, jn_hi, 2);
      mvn_lo = !(jn_lo <= 0x1p24f);
      mvn_hi = !(jn_hi <= 0x1p24f);
      any_next = mode == 0 && __any_sync(0xffffffffu, mvn_lo || mvn_hi);
    }
    if (mode == 0 && (bad || !(jn_lo <= 0x1p70f) || !(jn_hi <= 0x1p70f))) *panic = 1;  // (last tile's sums included)
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    tc::mbar_wait(my_pv_done, (n_tiles - 1) & 1);
    tc::fence_after_sync();
    const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
    uint32_t o[32];
    tc::tmem_ld_16x256b_x8(tO, o);
    tc::tmem_ld_wait();
    const int row_lo = q0 + x * BM + r_lo, row_hi = q0 + x * BM + r_hi;
    __nv_bfloat16* out_lo = ctx + ((size_t)s * Lp + row_lo) * LG_D + h * LG_DH + 2 * c4;
    __nv_bfloat16* out_hi = ctx + ((size_t)s * Lp + row_hi) * LG_D + h * LG_DH + 2 * c4;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (row_lo < nq)
        *reinterpret_cast<uint32_t*>(out_lo + 8 * k) =
            tc::pack_bf16(__uint_as_float(o[4 * k]) * inv_lo, __uint_as_float(o[4 * k + 1]) * inv_lo);
      if (row_hi < nq)
        *reinterpret_cast<uint32_t*>(out_hi + 8 * k) =
            tc::pack_bf16(__uint_as_float(o[4 * k + 2]) * inv_hi, __uint_as_float(o[4 * k + 3]) * inv_hi);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (mode != 0 || *panic == 0) break;
  // restart in the exact mode: everything issued in pass 0 has been consumed (both groups waited for their last P.V)
  tc::fence_after_sync();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint4*>(sQx)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  }  // pass
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int lg_tc_attention2(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                     const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, cudaStream_t st) {
  if (Lp % BM != 0) return LGB200_ERR_SHAPE;
  CUtensorMap tq, tk, tv;
  const uint64_t d[2] = {64, (uint64_t)S * LG_HEADS * Lp}, sb[1] = {128};
  const uint32_t box[2] = {64, 128};
  int rc;
  if ((rc = lg_make_tmap_bf16(&tq, Q, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tk, K, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tv, V, 2, d, sb, box))) return rc;
  cudaError_t e = cudaFuncSetAttribute(tc_attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((Lp + 2 * BM - 1) / (2 * BM), LG_HEADS, S);
  static const int start_mode = getenv("LGB200_ATTN_EXACT_MAX") ? atoi(getenv("LGB200_ATTN_EXACT_MAX")) : 0;
  tc_attention2_kernel<<<grid, A2_THREADS, A2_SMEM, st>>>(tq, tk, tv, Lp, lens, kv_xor, ctx, start_mode);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}
