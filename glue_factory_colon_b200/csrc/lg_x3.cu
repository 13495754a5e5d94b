// fp32-accurate linear layers and similarity GEMM on the 5th-gen tensor cores (tcgen05, sm_100a):
// "x3" = every operand is carried as TWO fp16 planes, x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits),
// and every product is three MMAs with fp32 accumulation in TMEM,
//
//     A.W^T  ~=  A_hi.W_hi^T + A_hi.W_lo^T + A_lo.W_hi^T          (the dropped lo.lo term is 2^-22 relative)
//
// so that `precision = fp32` (the reference's default numerics, 1e-3 on log_assignment) runs on tcgen05 instead of
// the CUDA cores (lg_simt.cu: 134 pairs/s at 64 x 2048 keypoints; cuBLAS fp32 gives the reference 136 on the same
// GPU).  Activations live in HBM as split planes [2][rows][K] fp16 (the same bytes as fp32), written by the producing
// epilogue; weights are split once at pack time.  The residual stream x is ALSO kept in fp32 (exact residual adds,
// token heads, returned descriptors).
//
// Kernel: persistent, warp-specialised, one CTA per SM (the v1 streaming skeleton of lg_tc_gemm.cu):
//   warp 0      TMA producer: per K block of 64 -- A_hi, A_lo (128 rows) and W_hi, W_lo (256 rows) = 96 KB, 2 stages
//   warp 1      MMA issuer:   3 x 4 tcgen05.mma (M=128, N=256, K=16, kind::f16 with fp16 operands) per stage
//   warps 2-5   epilogue:     tcgen05.ld (thread = row); bias / scale / rotary / residual, or LayerNorm + GELU(erf), or the
//                             raw fp32 similarity tile; results leave as fp32 and / or as split planes
// Reference call sites: lightglue.py:157-164 (Wqkv, out_proj), :144-149 (ffn), :193-222 (cross block), :281-284
// (final_proj and the similarity einsum of MatchAssignment).
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int XM = 128, XK = 64, XN = 256;
constexpr int PA = XM * XK * 2;            // one plane of an A stage (16 KB)
constexpr int PW = XN * XK * 2;            // one plane of a W stage (32 KB)
constexpr int XSTAGE = 2 * PA + 2 * PW;    // 96 KB
constexpr int XSTAGES = 2;
constexpr int X_SMEM = XSTAGES * XSTAGE + 1024 /*align*/ + 256 /*barriers*/ + 3 * 512 * 4 /*LN parameters*/;

enum { X_ROW = 0, X_HEADS = 1, X_LN = 2, X_SIM = 3 };

struct X3Args {
  int kb_total, kb_a0;     // K / 64, K0 / 64
  int n_tiles;             // column blocks of 256 per row tile (LN: 1 logical tile made of 2 blocks)
  int m_tiles;             // T / 128 (X_SIM: B * Lp / 128)
  const int32_t* lens;
  int Lp;
  // split output planes (nullable): [2][rows][N] fp16
  __half* outs;
  size_t outs_plane;       // elements between the hi and the lo plane
  size_t outp_plane;       // head-major part outputs: elements between the planes
  float* sim;              // X_SIM: [B][Lp][Lp] fp32
};

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn_major) {
  // kind::f16, A and B fp16 (format code 0), D fp32; same field layout as tc::idesc_bf16
  return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {  // planes of (a, b) * LG_X3_EA
  a *= LG_X3_EA;
  b *= LG_X3_EA;
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 32 consecutive values of one row -> 64 bytes in each plane
__device__ __forceinline__ void store_split32(__half* hi_ptr, size_t plane, const float (&v)[32]) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  uint4* dh = reinterpret_cast<uint4*>(hi_ptr);
  uint4* dl = reinterpret_cast<uint4*>(hi_ptr + plane);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dh[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    dl[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
  }
}

// tile t -> rows of A, rows of W, skipped?
struct TileRef {
  int a_row, w_row, seq_row0, m_tile, n_tile;
  bool skip;
};
template <int MODE>
__device__ __forceinline__ TileRef tile_ref(const X3Args& g, int t) {
  TileRef r;
  r.m_tile = t / g.n_tiles;
  r.n_tile = t - r.m_tile * g.n_tiles;
  r.skip = false;
  if constexpr (MODE == X_SIM) {
    const int tps = g.Lp / XM;
    const int b = r.m_tile / tps, mt = r.m_tile - b * tps;
    r.a_row = (2 * b) * g.Lp + mt * XM;
    r.w_row = (2 * b + 1) * g.Lp + r.n_tile * XN;
    r.seq_row0 = mt * XM;
    if (g.lens) r.skip = mt * XM >= g.lens[2 * b] || r.n_tile * XN >= g.lens[2 * b + 1];
  } else {
    r.a_row = r.m_tile * XM;
    r.w_row = r.n_tile * XN;
    const int s = r.a_row / g.Lp;
    r.seq_row0 = r.a_row - s * g.Lp;
    if (g.lens) r.skip = r.seq_row0 >= g.lens[s];
  }
  return r;
}

// Work items of one CTA.  CL == 1: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  CL > 1: the CTAs of a cluster walk
// GROUPS of CL consecutive row tiles x one column block together (rank r takes row tile G*CL + r), so that every W
// stage is needed by all of them at the same time: each fetches 1/CL of it and multicasts.  A group is skipped only if
// all of its row tiles are padding; a padded tile inside a live group is computed like any other (finite junk in rows
// nobody reads) because its CTA has to take part in the W stream anyway.
template <int MODE, int CL>
struct TileWalk {
  const X3Args& g;
  int it, step, rank, n_items;
  __device__ TileWalk(const X3Args& g_, int rank_) : g(g_), rank(rank_) {
    if (CL == 1) {
      it = blockIdx.x; step = gridDim.x; n_items = g.m_tiles * g.n_tiles;
    } else {
      it = blockIdx.x / CL; step = gridDim.x / CL; n_items = (g.m_tiles / CL) * g.n_tiles;
    }
  }
  __device__ bool next(TileRef& tr) {
    while (it < n_items) {
      const int cur = it;
      it += step;
      if (CL == 1) {
        tr = tile_ref<MODE>(g, cur);
        if (!tr.skip) return true;
      } else {
        const int G = cur / g.n_tiles, nt = cur - G * g.n_tiles;
        bool all_skip = true;
        for (int r = 0; r < CL; ++r) all_skip = all_skip && tile_ref<MODE>(g, (G * CL + r) * g.n_tiles + nt).skip;
        if (all_skip) continue;
        tr = tile_ref<MODE>(g, (G * CL + rank) * g.n_tiles + nt);
        tr.skip = false;
        return true;
      }
    }
    return false;
  }
};

template <int MODE, int CL>
__global__ void __launch_bounds__(192, 1)
x3_linear_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, X3Args g, LgEpi epi) {
  constexpr bool LN = MODE == X_LN;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  constexpr int WSL = XN / CL;           // W rows this CTA fetches for the whole cluster
  const uint32_t crank = CL > 1 ? tc::cluster_ctarank() : 0;
  // The tensor core rounds its fp32 accumulator TOWARD ZERO after every MMA (tools/x3_micro.py: K = 512 loses 1.5e-6
  // relative, one-sidedly; the CUDA-core kernel 3e-10).  Where the result feeds the residual stream or the score matrix
  // directly (X_ROW: FFN layer 2, final_proj; X_SIM) the two small products go to a SECOND accumulator, so that the large
  // one sees one rounding per K step instead of three; the epilogue adds the two in fp32.  (X_LN needs all 512 columns
  // for one row tile and its uniform shrink is removed by the LayerNorm itself; X_HEADS with a split accumulator gave
  // q / k / v at 4.3e-7 instead of 1.2e-6 rms but changed nothing in log_assignment and cost 8 % of the step, so both
  // keep two alternating accumulators.)
  constexpr bool SPLIT_ACC = MODE == X_ROW || MODE == X_SIM;
  constexpr int NSUB = LN ? 2 : 1;          // column blocks accumulated into one TMEM accumulator
  constexpr int ACC = (LN || SPLIT_ACC) ? 1 : 2;  // TMEM accumulators in flight (512 columns in total)
  constexpr int ACC_COLS = XN * NSUB;
  constexpr uint32_t LO_OFF = SPLIT_ACC ? XN : 0;  // column offset of the accumulator of the small products
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + XSTAGES * XSTAGE);
  uint64_t* empty = full + XSTAGES;
  uint64_t* tfull = empty + XSTAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_par = reinterpret_cast<float*>(smem + XSTAGES * XSTAGE + 256);  // bias | gamma | beta (LN)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA0);
    tc::prefetch_tmap(&tmA1);
    tc::prefetch_tmap(&tmW);
    // (cluster: a stage is refilled by every CTA's slice, so each CTA's MMAs release it in all of them)
    for (int i = 0; i < XSTAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], CL); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  if (LN) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      s_par[i] = epi.bias[i];
      s_par[512 + i] = epi.gamma[i];
      s_par[1024 + i] = epi.beta[i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc::cluster_sync();  // the peers' barriers exist before anyone multicasts into them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TileWalk<MODE, CL> walk(g, (int)crank);
      TileRef tr;
      while (walk.next(tr)) {
        for (int sub = 0; sub < NSUB; ++sub) {
          for (int kb = 0; kb < g.kb_total; ++kb) {
            tc::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * XSTAGE;
            uint8_t* sw = sa + 2 * PA;
            tc::mbar_arrive_expect_tx(&full[stage], XSTAGE);
            const CUtensorMap* ta = kb < g.kb_a0 ? &tmA0 : &tmA1;
            const int ka = (kb < g.kb_a0 ? kb : kb - g.kb_a0) * XK;
            tc::tma_load_3d(sa, ta, &full[stage], ka, tr.a_row, 0);
            tc::tma_load_3d(sa + PA, ta, &full[stage], ka, tr.a_row, 1);
            if (CL > 1) {
              const int wr = tr.w_row + sub * XN + (int)crank * WSL;
              tc::tma_load_3d_mc(sw + crank * (WSL * 128), &tmW, &full[stage], kb * XK, wr, 0, MC_MASK);
              tc::tma_load_3d_mc(sw + PW + crank * (WSL * 128), &tmW, &full[stage], kb * XK, wr, 1, MC_MASK);
            } else {
              tc::tma_load_3d(sw, &tmW, &full[stage], kb * XK, tr.w_row + sub * XN, 0);
              tc::tma_load_3d(sw + PW, &tmW, &full[stage], kb * XK, tr.w_row + sub * XN, 1);
            }
            if (++stage == XSTAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    constexpr uint32_t idesc = idesc_f16(XM, XN, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    TileWalk<MODE, CL> walk(g, (int)crank);
    TileRef tr;
    while (walk.next(tr)) {
      tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc::fence_after_sync();
      for (int sub = 0; sub < NSUB; ++sub) {
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS + sub * XN;
        for (int kb = 0; kb < g.kb_total; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + stage * XSTAGE);
          const uint32_t sw = sa + 2 * PA;
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < XK / 16; ++k) {
              const uint64_t a_hi = tc::smem_desc_sw128(sa + k * 32, 0, 1024);
              const uint64_t a_lo = tc::smem_desc_sw128(sa + PA + k * 32, 0, 1024);
              const uint64_t w_hi = tc::smem_desc_sw128(sw + k * 32, 0, 1024);
              const uint64_t w_lo = tc::smem_desc_sw128(sw + PW + k * 32, 0, 1024);
              // small terms first, the hi.hi product last
              tc::umma_ss(d_tmem + LO_OFF, a_lo, w_hi, idesc, (kb | k) != 0);
              tc::umma_ss(d_tmem + LO_OFF, a_hi, w_lo, idesc, 1);
              tc::umma_ss(d_tmem, a_hi, w_hi, idesc, SPLIT_ACC ? (kb | k) != 0 : 1);
            }
            if (CL > 1) tc::umma_commit_mc(&empty[stage], MC_MASK);  // stage reusable once these MMAs retire
            else tc::umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == XSTAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (tc::elect_one()) tc::umma_commit(&tfull[acc]);  // accumulator complete
      __syncwarp();
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    TileWalk<MODE, CL> walk(g, (int)crank);
    TileRef tr;
    while (walk.next(tr)) {
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::fence_after_sync();
      const int rt = quarter * 32 + lane;          // row within the tile
      const int row = tr.a_row + rt;               // global row of A (= output row, except X_SIM)
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_COLS;
      if constexpr (MODE == X_SIM) {
        const int b = tr.a_row / (2 * g.Lp);
        float* out = g.sim + ((size_t)b * g.Lp + tr.seq_row0 + rt) * g.Lp + tr.n_tile * XN;
#pragma unroll 1
        for (int c = 0; c < XN / 32; ++c) {
          uint32_t r[32], rl[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld32(t_row + LO_OFF + c * 32, rl);
          tc::tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(out + c * 32);
          constexpr float us = 1.f / (LG_X3_EA * LG_X3_EA);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4((__uint_as_float(r[4 * q]) + __uint_as_float(rl[4 * q])) * us,
                                 (__uint_as_float(r[4 * q + 1]) + __uint_as_float(rl[4 * q + 1])) * us,
                                 (__uint_as_float(r[4 * q + 2]) + __uint_as_float(rl[4 * q + 2])) * us,
                                 (__uint_as_float(r[4 * q + 3]) + __uint_as_float(rl[4 * q + 3])) * us);
        }
      } else if constexpr (!LN) {
        const int n0 = tr.n_tile * XN;
#pragma unroll 1
        for (int c = 0; c < XN / 32; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          if constexpr (SPLIT_ACC) {
            uint32_t rl[32];
            tc::tmem_ld32(t_row + LO_OFF + c * 32, rl);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(rl[i]));
          } else {
            tc::tmem_ld_wait();
          }
          const int col = n0 + c * 32;
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            constexpr float us = 1.f / (LG_X3_EA * LG_X3_EW);  // undo the plane scaling (exact)
            const float4 b4 = *reinterpret_cast<const float4*>(epi.bias + col + 4 * q);
            v[4 * q] = fmaf(__uint_as_float(r[4 * q]), us, b4.x);
            v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), us, b4.y);
            v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), us, b4.z);
            v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), us, b4.w);
          }
          if constexpr (MODE == X_ROW) {
            const size_t off = (size_t)row * epi.N + col;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= epi.scale[0];
            if (epi.resid32) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 r4 = *reinterpret_cast<const float4*>(epi.resid32 + off + 4 * q);
                v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
              }
            }
            if (epi.out32) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(epi.out32 + off + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (g.outs) store_split32(g.outs + off, g.outs_plane, v);
          } else {  // X_HEADS: the 32 columns lie in one (part, head)
            const int part = col >> 8, h = (col >> 6) & 3, d = col & 63;
            if (part < epi.n_rot) {  // rotary pairs (d, d+1) use table entry d/2 = (cos, sin) at rot[row*64 + d]
              const float* rt_ = epi.rot + (size_t)row * 64 + d;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 cs = *reinterpret_cast<const float4*>(rt_ + 4 * q);
                const float a0 = v[4 * q] * cs.x - v[4 * q + 1] * cs.y, a1 = v[4 * q + 1] * cs.x + v[4 * q] * cs.y;
                const float a2 = v[4 * q + 2] * cs.z - v[4 * q + 3] * cs.w, a3 = v[4 * q + 3] * cs.z + v[4 * q + 2] * cs.w;
                v[4 * q] = a0; v[4 * q + 1] = a1; v[4 * q + 2] = a2; v[4 * q + 3] = a3;
              }
            }
            const float sc = epi.scale[part];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= sc;
            const int s = row / epi.Lp, l = row - s * epi.Lp;
            const size_t off = (((size_t)s * LG_HEADS + h) * epi.Lp + l) * LG_DH + d;
            store_split32(reinterpret_cast<__half*>(epi.outp[part]) + off, g.outp_plane, v);
          }
        }
      } else {
        // LayerNorm(512) + GELU(erf), lightglue.py:144-149.  Pass 1: row statistics (thread = row)
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = fmaf(__uint_as_float(r[j]), 1.f / (LG_X3_EA * LG_X3_EW), s_par[c * 32 + j]);
            sum += v;
            sq = fmaf(v, v, sq);
          }
        }
        const float mean = sum * (1.f / 512.f);
        // second pass over the row for the variance: E[(v - mean)^2] (the one-pass form loses bits at 1e-3 parity)
        float var = 0.f;
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dv = fmaf(__uint_as_float(r[j]), 1.f / (LG_X3_EA * LG_X3_EW), s_par[c * 32 + j]) - mean;
            var = fmaf(dv, dv, var);
          }
        }
        (void)sq;
        const float rstd = rsqrtf(var * (1.f / 512.f) + 1e-5f);
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          uint32_t r[32];
          tc::tmem_ld32(t_row + c * 32, r);
          tc::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c * 32 + j;
            const float a = (fmaf(__uint_as_float(r[j]), 1.f / (LG_X3_EA * LG_X3_EW), s_par[col]) - mean) * rstd * s_par[512 + col] +
                            s_par[1024 + col];
            v[j] = lg_gelu_erf(a);
          }
          const size_t off = (size_t)row * 512 + c * 32;
          if (epi.out32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(epi.out32 + off + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          if (g.outs) store_split32(g.outs + off, g.outs_plane, v);
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc::cluster_sync();  // nobody retires while a peer may still multicast into its shared memory
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// split planes [2][rows][K] fp16 -> 3-D tensor map (K, rows, plane), box 64 x box_rows x 1
int make_split_map(CUtensorMap* tm, const __half* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  const uint64_t d[3] = {K, rows, 2}, s[2] = {K * 2, rows * K * 2};
  const uint32_t b[3] = {64, box_rows, 1};
  return lg_make_tmap_bf16(tm, base, 3, d, s, b);  // (2-byte elements: the bf16 encoder serves fp16 as well)
}

#ifndef LG_X3_CL
#define LG_X3_CL 1
#endif
// Row tiles that share a W stream (cluster size of the linear modes; opt-in with -DLG_X3_CL=2|4).  One CTA per 128-row
// tile streams the whole weight matrix (both planes: 512 KB at K = 512, N = 256) next to 256 KB of activations, 1.6 GB of
// L2 -> SM traffic per FFN2 launch at 64 x 2048 keypoints, so W multicast across a cluster looked like the fix.
// Measured on B200 (whole fp32-mode step, same box): cluster 1 61.0 ms, cluster 2 60.6 ms, cluster 4 76.8 ms -- the
// kernel is not L2-bound but SHARED-MEMORY-bound: the 12 MMAs of a stage read 144 KB of operands and the stage's TMA
// fill writes 96 KB, 240 KB per 1 536 tensor cycles = 156 B/clk against the SM's 128 B/clk; lockstep with only two
// stages then costs more than the halved L2 stream saves.  The fix for that is cta_group::2 (half of B per CTA, as in
// the bf16 pair kernels of lg_tc_gemm2.cu), not multicast.  With a cluster, padded row tiles inside a live group are
// computed (finite junk in rows nobody reads) instead of skipped.
constexpr int X_CL = LG_X3_CL;

template <int MODE, int CL>
int launch_x3_cl(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, X3Args g, LgEpi epi, cudaStream_t st) {
  auto kern = x3_linear_kernel<MODE, CL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = (g.m_tiles / CL) * g.n_tiles;           // groups of CL row tiles x column blocks
  const int clusters = items < sms / CL ? items : sms / CL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = X_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, a0, a1, w, g, epi);
  if (e != cudaSuccess) return (int)e;
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

// cluster size a launch can use: the row tiles must split into whole groups
inline int x3_cluster(int mode, int m_tiles) { return (mode != X_SIM && X_CL > 1 && m_tiles % X_CL == 0) ? X_CL : 1; }

template <int MODE>
int launch_x3(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, X3Args g, LgEpi epi, cudaStream_t st) {
  if (x3_cluster(MODE, g.m_tiles) > 1) return launch_x3_cl<MODE, X_CL>(a0, a1, w, g, epi, st);
  return launch_x3_cl<MODE, 1>(a0, a1, w, g, epi, st);
}

// ---------------------------------------------------------------------------------------------------------------
// small fp32 kernels around the GEMMs

// x [rows][D] fp32 -> split planes [2][rows][D] fp16 (after input staging and after point pruning)
__global__ void x3_split_kernel(const float* __restrict__ x, size_t n4, __half* __restrict__ xs, size_t plane) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  reinterpret_cast<uint2*>(xs)[i] = make_uint2(h0, h1);
  reinterpret_cast<uint2*>(xs + plane)[i] = make_uint2(l0, l1);
}

// Gradients have no fixed range: lgb200_split_dynamic scales by the power of two g that brings max |x| into [256, 512)
// before splitting (a fixed scaling would leave small gradients in the fp16 subnormals).
__global__ void __launch_bounds__(256) x3_absmax_kernel(const float* __restrict__ x, size_t n4, unsigned* __restrict__ slot) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    if (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) m = INFINITY;  // (fmaxf drops NaN)
  }
  for (int ofs = 16; ofs; ofs >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, ofs));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
}
// the same pass with per-block column sums (bias gradients) on the side: x [rows][4 * blockDim.x]; blockDim.y row lanes
// per CTA, four rows in flight per thread; CTA b sums rows b * RY + ty, + gridDim.x * RY, ... into partials[b][cols]; the
// caller adds the gridDim.x partial rows.  (One row lane and one load in flight per thread ran at 1 TB/s: 57 us per
// 33-100 MB tensor, 5 % of the training step.)
__global__ void __launch_bounds__(512) x3_absmax_colsum_kernel(const float* __restrict__ x, int rows, unsigned* __restrict__ slot,
                                                               float* __restrict__ partials) {
  __shared__ float4 red[512];
  const int c4 = threadIdx.x, n4 = blockDim.x, ty = threadIdx.y, ry = blockDim.y;
  const int stride = gridDim.x * ry;
  float m = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto take = [&](const float4 v) {
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    if (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) m = INFINITY;  // (fmaxf drops NaN)
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  };
  const float4* x4 = reinterpret_cast<const float4*>(x);
  int r = blockIdx.x * ry + ty;
  for (; r + 3 * stride < rows; r += 4 * stride) {
    const float4 v0 = x4[(size_t)r * n4 + c4], v1 = x4[(size_t)(r + stride) * n4 + c4];
    const float4 v2 = x4[(size_t)(r + 2 * stride) * n4 + c4], v3 = x4[(size_t)(r + 3 * stride) * n4 + c4];
    take(v0); take(v1); take(v2); take(v3);
  }
  for (; r < rows; r += stride) take(x4[(size_t)r * n4 + c4]);
  red[ty * n4 + c4] = acc;
  __syncthreads();
  if (ty == 0) {
    for (int y = 1; y < ry; ++y) {
      const float4 o = red[y * n4 + c4];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    reinterpret_cast<float4*>(partials)[(size_t)blockIdx.x * n4 + c4] = acc;
  }
  for (int ofs = 16; ofs; ofs >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, ofs));
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
}
// split planes [2][n] (value * scale_in = hi + lo) -> fp32
__global__ void x3_merge_kernel(const __half* __restrict__ xs, size_t plane, size_t n4, float inv_scale, float* __restrict__ x) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uint2 h = reinterpret_cast<const uint2*>(xs)[i], l = reinterpret_cast<const uint2*>(xs + plane)[i];
  const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
  const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
  reinterpret_cast<float4*>(x)[i] = make_float4((h0.x + l0.x) * inv_scale, (h0.y + l0.y) * inv_scale,
                                                (h1.x + l1.x) * inv_scale, (h1.y + l1.y) * inv_scale);
}
__device__ __forceinline__ float x3_dynamic_scale(unsigned bits) {  // 1 for an all-zero or non-finite tensor
  const float m = __uint_as_float(bits);
  if (!(m > 0.f && m < INFINITY)) return 1.f;
  int e;
  frexpf(m, &e);  // m = f 2^e, f in [0.5, 1)
  e = 9 - e;
  return ldexpf(1.f, e < -100 ? -100 : (e > 100 ? 100 : e));
}
__global__ void x3_split_dynamic_kernel(const float* __restrict__ x, size_t n4, const unsigned* __restrict__ slot,
                                        __half* __restrict__ xs, size_t plane, float* __restrict__ inv_scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float g = x3_dynamic_scale(*slot);
  if (i == 0) *inv_scale = 1.f / g;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const __half2 h0 = __floats2half2_rn(v.x * g, v.y * g), h1 = __floats2half2_rn(v.z * g, v.w * g);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v.x * g - f0.x, v.y * g - f0.y), l1 = __floats2half2_rn(v.z * g - f1.x, v.w * g - f1.y);
  reinterpret_cast<uint2*>(xs)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  reinterpret_cast<uint2*>(xs + plane)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// MatchAssignment normalisers from the similarity sim [B][Lp][Lp] (lightglue.py:262-263):
// lse[2b, i] = logsumexp_j sim[b,i,j] (one warp per row) and lse[2b+1, j] = logsumexp_i sim[b,i,j] (one thread per
// column, 8 warps striding the rows, combined through shared memory).
__global__ void __launch_bounds__(256) x3_lse_rows_kernel(const float* __restrict__ sim, int Lp,
                                                          const int32_t* __restrict__ lens, int n0_all, int n1_all,
                                                          float* __restrict__ lse) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = lens ? lens[2 * b] : n0_all, n1 = lens ? lens[2 * b + 1] : n1_all;
  const int i = blockIdx.x * 8 + warp;
  if (i >= n0) return;
  const float* row = sim + ((size_t)b * Lp + i) * Lp;
  float m = -INFINITY;
  for (int j = lane; j < n1; j += 32) m = fmaxf(m, row[j]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
  for (int j = lane; j < n1; j += 32) l += expf(row[j] - m);
  for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) lse[(size_t)(2 * b) * Lp + i] = n1 > 0 ? m + logf(l) : 0.f;
}

__global__ void __launch_bounds__(256) x3_lse_cols_kernel(const float* __restrict__ sim, int Lp,
                                                          const int32_t* __restrict__ lens, int n0_all, int n1_all,
                                                          float* __restrict__ lse) {
  __shared__ float sm[8][32], sl[8][32];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = lens ? lens[2 * b] : n0_all, n1 = lens ? lens[2 * b + 1] : n1_all;
  const int j = blockIdx.x * 32 + lane;
  if (blockIdx.x * 32 >= n1) return;
  const float* col = sim + (size_t)b * Lp * Lp + j;
  float m = -INFINITY, l = 0.f;
  if (j < n1) {
    for (int i = warp; i < n0; i += 8) {  // online (max, sum): one pass over the column
      const float v = col[(size_t)i * Lp];
      const float mn = fmaxf(m, v);
      l = l * expf(m - mn) + expf(v - mn);
      m = mn;
    }
  }
  sm[warp][lane] = m;
  sl[warp][lane] = l;
  __syncthreads();
  if (warp == 0 && j < n1) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) mm = fmaxf(mm, sm[w][lane]);
    float ll = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) ll += sm[w][lane] == -INFINITY ? 0.f : sl[w][lane] * expf(sm[w][lane] - mm);
    lse[(size_t)(2 * b + 1) * Lp + j] = n0 > 0 ? mm + logf(ll) : 0.f;
  }
}

// scores [B][R][C] from sim (lightglue.py:257-269): inner block, dustbin row / column, zero padding
__global__ void __launch_bounds__(256) x3_scores_kernel(const float* __restrict__ sim, const float* __restrict__ z,
                                                        const float* __restrict__ lse, int Lp,
                                                        const int32_t* __restrict__ lens, int R, int C,
                                                        float* __restrict__ scores) {
  const int b = blockIdx.z, r = blockIdx.y;
  const int na = lens ? lens[2 * b] : R - 1, nb = lens ? lens[2 * b + 1] : C - 1;
  const float* zr = z + (size_t)(2 * b) * Lp;
  const float* zc = z + (size_t)(2 * b + 1) * Lp;
  const float* lc = lse + (size_t)(2 * b + 1) * Lp;
  float* out = scores + ((size_t)b * R + r) * C;
  float rz = 0.f, rl = 0.f, rneg = 0.f;
  if (r < na) {
    const float zz = zr[r];
    rz = lg_logsigmoid(zz);
    rneg = lg_logsigmoid(-zz);
    rl = lse[(size_t)(2 * b) * Lp + r];
  }
  const float* srow = sim + ((size_t)b * Lp + (r < Lp ? r : 0)) * Lp;
  for (int c = blockIdx.x * 1024 + threadIdx.x; c < min(C, (int)(blockIdx.x + 1) * 1024); c += 256) {
    float v = 0.f;
    if (r < na && c < nb) {
      const float a = srow[c];
      v = (a - rl) + (a - lc[c]) + (rz + lg_logsigmoid(zc[c]));  // same association as the CUDA-core kernel
    } else if (r < na && c == C - 1) v = rneg;
    else if (r == R - 1 && c < nb) v = lg_logsigmoid(-zc[c]);
    out[c] = v;
  }
}

}  // namespace

int lg_x3_linear(int epilogue, const void* A0, const void* A1, int K0, const void* W, int T, int N, int K,
                 const int32_t* lens, LgEpi epi, void* outs, cudaStream_t st) {
  if (T % XM || K % XK || K0 % XK || N % XN) return LGB200_ERR_SHAPE;
  const bool ln = epilogue == LGB200_EPI_LN_GELU;
  if (ln && N != 512) return LGB200_ERR_SHAPE;
  if (epilogue == LGB200_EPI_ROWMAJOR && !epi.out32 && !outs) return LGB200_ERR_NULL;
  if (epilogue == LGB200_EPI_HEADS && epi.n_rot > 0 && !epi.rot) return LGB200_ERR_NULL;
  CUtensorMap tA0, tA1, tW;
  int rc;
  if ((rc = make_split_map(&tA0, (const __half*)A0, T, K0, XM))) return rc;
  if (K0 < K) {
    if ((rc = make_split_map(&tA1, (const __half*)A1, T, K - K0, XM))) return rc;
  } else {
    tA1 = tA0;
  }
  const int mode = ln ? X_LN : (epilogue == LGB200_EPI_HEADS ? X_HEADS : X_ROW);
  if ((rc = make_split_map(&tW, (const __half*)W, N, K, XN / x3_cluster(mode, T / XM)))) return rc;  // box = one CTA's slice
  X3Args g{};
  g.kb_total = K / XK;
  g.kb_a0 = K0 / XK;
  g.n_tiles = ln ? 1 : N / XN;
  g.m_tiles = T / XM;
  g.lens = lens;
  g.Lp = epi.Lp;
  g.outs = (__half*)outs;
  g.outs_plane = (size_t)T * N;
  g.outp_plane = (size_t)T * LG_D;
  if (ln) return launch_x3<X_LN>(tA0, tA1, tW, g, epi, st);
  if (epilogue == LGB200_EPI_HEADS) return launch_x3<X_HEADS>(tA0, tA1, tW, g, epi, st);
  return launch_x3<X_ROW>(tA0, tA1, tW, g, epi, st);
}

extern "C" int lgb200_split_rows(const float* x, long long n, void* xs, void* stream) {
  if (!x || !xs) return LGB200_ERR_NULL;
  if (n <= 0 || n % 4) return LGB200_ERR_SHAPE;
  const size_t n4 = (size_t)n / 4;
  x3_split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, lg_stream(stream)>>>(x, n4, (__half*)xs, (size_t)n);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

extern "C" int lgb200_split_dynamic(const float* x, long long n, void* xs, float* inv_scale, int cols,
                                    float* colsum_partials, int n_partials, void* stream) {
  if (!x || !xs || !inv_scale) return LGB200_ERR_NULL;
  if (n <= 0 || n % 4) return LGB200_ERR_SHAPE;
  if (colsum_partials && (cols <= 0 || cols % 4 || cols > 2048 || n % cols || n_partials <= 0)) return LGB200_ERR_SHAPE;
  const size_t n4 = (size_t)n / 4;
  cudaStream_t st = lg_stream(stream);
  unsigned* slot = reinterpret_cast<unsigned*>(inv_scale) + 1;  // inv_scale[1] holds the bit pattern of max |x|
  cudaError_t e;
  if ((e = cudaMemsetAsync(slot, 0, sizeof(unsigned), st)) != cudaSuccess) return (int)e;
  const unsigned nb = (unsigned)((n4 + 255) / 256);
  if (colsum_partials)
    x3_absmax_colsum_kernel<<<n_partials, dim3(cols / 4, 512 / (cols / 4) ? 512 / (cols / 4) : 1), 0, st>>>(x, (int)(n / cols), slot,
                                                                                                            colsum_partials);
  else
    x3_absmax_kernel<<<nb < 1184u ? nb : 1184u, 256, 0, st>>>(x, n4, slot);
  LG_LAUNCH_CHECK();
  x3_split_dynamic_kernel<<<nb, 256, 0, st>>>(x, n4, slot, (__half*)xs, (size_t)n, inv_scale);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

extern "C" int lgb200_merge_rows(const void* xs, long long n, float inv_scale, float* x, void* stream) {
  if (!x || !xs) return LGB200_ERR_NULL;
  if (n <= 0 || n % 4) return LGB200_ERR_SHAPE;
  const size_t n4 = (size_t)n / 4;
  x3_merge_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, lg_stream(stream)>>>((const __half*)xs, (size_t)n, n4, inv_scale, x);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

extern "C" int lgb200_x3_similarity(const void* mds, int B, int Lp, const int32_t* lens, float* sim, void* stream) {
  if (!mds || !sim) return LGB200_ERR_NULL;
  if (B <= 0 || Lp <= 0 || Lp % XN) return LGB200_ERR_SHAPE;
  const int T = 2 * B * Lp;
  CUtensorMap tA, tW;
  int rc;
  if ((rc = make_split_map(&tA, (const __half*)mds, T, LG_D, XM))) return rc;
  if ((rc = make_split_map(&tW, (const __half*)mds, T, LG_D, XN))) return rc;
  X3Args g{};
  g.kb_total = g.kb_a0 = LG_D / XK;
  g.n_tiles = Lp / XN;
  g.m_tiles = B * (Lp / XM);
  g.lens = lens;
  g.Lp = Lp;
  g.sim = sim;
  LgEpi epi{};
  epi.Lp = Lp;
  return launch_x3<X_SIM>(tA, tA, tW, g, epi, lg_stream(stream));
}

extern "C" int lgb200_x3_assign_lse(const float* sim, int B, int Lp, const int32_t* lens, int n0, int n1, float* lse,
                                    void* stream) {
  if (!sim || !lse) return LGB200_ERR_NULL;
  if (B <= 0 || Lp <= 0 || n0 < 0 || n1 < 0 || n0 > Lp || n1 > Lp) return LGB200_ERR_SHAPE;
  cudaStream_t st = lg_stream(stream);
  if (n0 > 0) {
    x3_lse_rows_kernel<<<dim3((n0 + 7) / 8, B), 256, 0, st>>>(sim, Lp, lens, n0, n1, lse);
    LG_LAUNCH_CHECK();
  }
  if (n1 > 0) {
    x3_lse_cols_kernel<<<dim3((n1 + 31) / 32, B), 256, 0, st>>>(sim, Lp, lens, n0, n1, lse);
    LG_LAUNCH_CHECK();
  }
  return LGB200_OK;
}

extern "C" int lgb200_x3_assign_scores(const float* sim, const float* z, const float* lse, int B, int Lp,
                                       const int32_t* lens, int R, int C, float* scores, void* stream) {
  if (!sim || !z || !lse || !scores) return LGB200_ERR_NULL;
  if (B <= 0 || Lp <= 0 || R < 1 || C < 1 || R - 1 > Lp || C - 1 > Lp) return LGB200_ERR_SHAPE;
  x3_scores_kernel<<<dim3((C + 1023) / 1024, R, B), 256, 0, lg_stream(stream)>>>(sim, z, lse, Lp, lens, R, C, scores);
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}
