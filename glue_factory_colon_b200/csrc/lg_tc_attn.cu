// Flash attention forward on tcgen05 / TMEM / TMA (sm_100a), head_dim 64, bf16 in, fp32 accumulate.
//
// One CTA = 128 query rows of one (sequence, head).  Two CTAs are resident per SM (80 KB smem,
// 256 TMEM columns each) so that one CTA's softmax overlaps the other's MMAs.
//   warp 0     TMA producer: Q once, then K and V tiles of 128 keys into two independent 3-stage rings
//              (a K stage is released as soon as its QK^T retires, so K runs ~3 tiles ahead; with the
//              first 2-stage K/V ring the ~2 us TMA latency was exposed on every step: 2830 cycles/step)
//   warp 1     MMA issuer:   S = Q.K^T (SS, N=128) into TMEM, O += P.V (TS: P read from TMEM, V MN-major)
//   warps 2-9  softmax:      two threads per query row (64 key columns each).  tcgen05.ld S -> registers, online max with lazy
//                            rescale (O is only touched when the max grows by > 8 in log2 units),
//                            ex2, row sum, bf16 P written back to TMEM with tcgen05.st
// Issue order QK(j+1) before PV(j): the next score tile is produced while the softmax warps are
// still exponentiating tile j, and PV(j) runs while they work on tile j+1.
// TMEM columns: [0,128) S fp32 | [128,192) P bf16x2 | [192,256) O fp32.
// Scores arrive in the log2 domain (the Q projection epilogue folds log2(e)/sqrt(64)).
// Keys >= lens[kv sequence] are masked to -inf; query rows >= lens[q sequence] are not stored.
//
// r1 instruction diet (the loop is issue-bound): max subtraction folded into the QK^T MMA (LG_ATTN_MSUB, 0.721 ->
// 0.697 ms), polynomial share retuned to 3/16 (0.681), packed f32x2 row sum (0.675 ms = 814 TFLOP/s).
// Structures tried against this one at S=128, Lp=2048 (0.72 ms), all parity-green, none faster (git history,
// DESIGN.md section 3.2): persistent CTAs with a dynamic work ring (0.74-0.75), issuer/producer at the highest
// warp ids (0.75), sleeping waits for issuer/producer (no change), four "fat" softmax warps with one thread per
// row (0.76), 256 queries per CTA sharing each K/V tile with all 512 TMEM columns (0.82).  Removing all MUFU work
// (-8 %) or all MMAs (-9 %) moves the time as little: the softmax warps issue only 60-66 % of the cycles and the
// rest is fixed-latency dependency stalls spread over the whole unrolled loop (ncu source page) -- the lever that
// is left is the instruction count per score element (tools/micro/pipe_rate.cu: 1 warp-instruction/clk/SMSP).
#include "lg_internal.cuh"
#include "lg_tc_common.cuh"
#include <stdlib.h>

// debug timeline (clock64 stamps of CTA (0,0,0); read back with lgb200_debug_attn_times)
__device__ long long g_attn_times[2 * 16 * 16];
// Arrival counter per SM: the two CTAs that share an SM take alternating slots, and the odd one starts half
// a step late.  Without this the co-resident CTAs run in lockstep: all 16 softmax warps exponentiate at the
// same time (MUFU 100 % busy for ~2000 cycles) and then all leave it idle for ~1000 (measured timeline).
__device__ unsigned int g_attn_sm_slot[1024];

#ifndef LG_DBG_Z
#define LG_DBG_Z 0  // sequence index of the CTA whose clock64 timeline is recorded (debug builds)
#endif
#ifndef LG_ATTN_MSUB
#define LG_ATTN_MSUB 1  // 1: the row-max subtraction s - m_ref is folded into the QK^T MMA (a fifth K=16 slice)
#endif
#ifndef LG_ATTN_POLY16
#define LG_ATTN_POLY16 5  // LG_ATTN_POLY16 of every 16 exponentials are evaluated by polynomial on the FMA pipe (0..8)
#endif
#ifndef LG_ATTN_POLY_DEG
#define LG_ATTN_POLY_DEG 3
#endif
#ifndef LG_ATTN_SPLIT_P
#define LG_ATTN_SPLIT_P 0  // 1: the fast path stores P in two halves, with the P.V(j-1) wait between them
#endif
#ifndef LG_ATTN_EARLY_POLL
// 1: the softmax warps poll pv_done(j-1) / s_full(j+1) with a non-blocking test one phase of work BEFORE they need the
// answer (a successful mbarrier try_wait still costs ~100-130 cycles of latency: ncu source page of r2, 15 % + 6 % of a
// softmax warp's tile time sat on the two polls), and fall back to the blocking wait only if the early test failed
#define LG_ATTN_EARLY_POLL 1
#endif
#ifndef LG_ATTN_PRELOAD
// 1: the score tile of step j+1 is loaded (tcgen05.ld) right behind the P store of step j, so that the TMEM load
// latency overlaps the store's completion wait and the p_ready hand-shake instead of opening the next step
#define LG_ATTN_PRELOAD 1
#endif
#ifndef LG_ATTN_PVTEST_AT
#define LG_ATTN_PVTEST_AT 20  // pair index (of 32) in the exponential loop at which pv_done(j-1) is polled (PRELOAD)
#endif
#ifndef LG_ATTN_POLY_PACKED
#define LG_ATTN_POLY_PACKED 1  // 1: the polynomial exponentials are evaluated two at a time with packed f32x2 FMA-pipe ops
#endif
#if LG_ATTN_POLY_PACKED
// pair i of a thread's 32 pairs: BOTH of its exponentials go through the polynomial (LG_ATTN_POLY16 of every 16 pairs)
#define LG_POLY_PAIR(i) ((((i) * LG_ATTN_POLY16) % 16) < LG_ATTN_POLY16)
#define LG_POLY_HERE(i) 0
#else
#define LG_POLY_PAIR(i) 0
#define LG_POLY_HERE(i) ((((i) * LG_ATTN_POLY16) % 8) < LG_ATTN_POLY16)
#endif

// waits of the TMA-producer and MMA-issuer warps: polling with a 64 ns sleep in between (a bare try_wait loop takes
// issue slots from the softmax warps of the same sub-partitions: 0.661 -> 0.650 ms with the deferred-maximum loop, which
// is short enough to feel it; with the r1 loop it made no difference).  -DLG_ATTN_SPIN restores the spinning waits.
#ifdef LG_ATTN_SPIN
#define LG_PI_WAIT(...) tc::mbar_wait(__VA_ARGS__)
#else
#define LG_PI_WAIT(...) tc::mbar_wait_relaxed(__VA_ARGS__)
#endif

namespace {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per step
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int KST = 3, VST = 2;            // K / V ring depth
constexpr int XT_BYTES = 128 * 16 * 2;     // 4 KB: 128 rows x 16 bf16 (one K=16 MMA slice), 32-byte swizzle
constexpr int AT_XOFF = TILE_BYTES * (1 + KST + VST);             // Q_ext | K_ext (LG_ATTN_MSUB)
constexpr int AT_BAROFF = AT_XOFF + (LG_ATTN_MSUB ? 2 * XT_BYTES : 0);
constexpr int AT_NBAR = 16;                // mbarriers per pass (15 used)
constexpr int AT_BARBYTES = 2 * AT_NBAR * 8 + 64;  // two barrier sets (pass 0 / restart) + tmem slot + panic flag
// + barriers + exchange ring [4 slots][128 rows][NP parts] fp32 (tile sums; maxima in the exact mode; final row sums)
constexpr int at_smem(int np) { return AT_BAROFF + AT_BARBYTES + 4 * 128 * np * 4; }

constexpr uint32_t TM_S = 0, TM_P = 128, TM_O = 192, TM_COLS = 256;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 8 on the FMA/ALU pipes (the MUFU unit delivers only 16 ex2/clk/SM, which is what bounds
// this kernel at d = 64): round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax polynomial
// (max rel. err 1.0e-4, far below the bf16 rounding of P), exponent patched in with integer ops.
// tools/micro/softmax_rate.cu: 13.7 -> 15.3 elements/clk/SM with one exponential in four done this way;
// In this kernel at S=128, Lp=2048, with the subtraction folded into the MMA (LG_ATTN_MSUB), polynomial share of
// 0 / 1 / 2 / 3 / 4 / 5 / 6 / 8 sixteenths: 0.728 / 0.705 / 0.693 / 0.681 / 0.690 / 0.703 / 0.731 / 0.764 ms -- the
// polynomial costs 8 issue slots against 1 for MUFU, and the loop is bound by issue slots as much as by MUFU.
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ float ex2_poly(float x) {
  // (valid for x < 127.5: the exponent patch below wraps silently above; the deferred-maximum mode, where a score may
  // exceed the reference by any amount, keeps the maximum of the polynomial lanes' inputs and checks it per tile)
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
#if LG_ATTN_POLY_DEG == 2  // max rel. err 1.7e-3 (a bf16 half-ulp is 1.95e-3)
  float p = fmaf(f, 0.23842894f, 0.70344801f);
  p = fmaf(p, f, 1.00044314f);
#else
  float p = fmaf(f, 0.05500889f, 0.24221097f);
  p = fmaf(p, f, 0.69328294f);
  p = fmaf(p, f, 1.0f);
#endif
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Two polynomial exponentials per call on packed f32x2 operations (add.f32x2 / fma.rn.f32x2: one issue slot for both
// lanes): 2 FMNMX + 3 FADD2 + 3 FFMA2 + 2 integer ops per pair = 5 issue slots per exponential instead of 8.5 -- the
// loop is bound by issue slots as much as by the MUFU unit, so a cheaper polynomial moves the optimum share up.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 u = __fadd2_rn(t, nmagic);
  const float2 f = __fadd2_rn(x, make_float2(-u.x, -u.y));
#if LG_ATTN_POLY_DEG == 2
  float2 p = __ffma2_rn(f, make_float2(0.23842894f, 0.23842894f), make_float2(0.70344801f, 0.70344801f));
  p = __ffma2_rn(p, f, make_float2(1.00044314f, 1.00044314f));
#else
  float2 p = __ffma2_rn(f, make_float2(0.05500889f, 0.05500889f), make_float2(0.24221097f, 0.24221097f));
  p = __ffma2_rn(p, f, make_float2(0.69328294f, 0.69328294f));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
#endif
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// NP = softmax threads per query row (2: 8 softmax warps, 64 key columns each; 4: 16 warps, 32 columns each)
template <int CL, int NP>
__global__ void __launch_bounds__(64 + 128 * NP, 2)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, int Lp, const int32_t* __restrict__ lens,
                    int kv_xor, __nv_bfloat16* __restrict__ ctx, int dbg_arg, unsigned stagger_ns, int start_mode,
                    const int32_t* __restrict__ order) {
  // Debug modes (skeleton runs, no-MUFU run, clock64 timeline) exist only when the file is compiled with
  // -DLG_ATTN_DEBUG; in the product build `dbg` is the constant 0 and every debug branch folds away
  // (leaving them as run-time branches cost ~50 BRA per 64 exponentials in the unrolled loop).
#ifdef LG_ATTN_DEBUG
  const int dbg_in = dbg_arg;
#else
  constexpr int dbg_in = 0;
#endif
  const int dbg = dbg_in & 15;  // bit 4 of dbg_in enables the clock64 timeline
#ifdef LG_ATTN_DEBUG
#define XSTAMP(k) do { if ((dbg_in & 16) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == LG_DBG_Z && threadIdx.x == 64) g_attn_times[506 + (k)] = clock64(); } while (0)
#else
#define XSTAMP(k) do {} while (0)
#endif
  XSTAMP(0);
  // CL CTAs with consecutive query tiles of the same (sequence, head) form a cluster and share every
  // K/V tile: each loads 1/CL of it and TMA-multicasts it to the others.  (Measured: with one CTA per
  // K/V tile the kernel sat at ~5 TB/s of L2->SM traffic regardless of MUFU / pipelining changes.)
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  constexpr int SLICE = AT_BN / CL;  // K/V rows this CTA loads per tile
  const uint32_t crank = CL > 1 ? tc::cluster_ctarank() : 0;
  const int h = blockIdx.y, q0 = blockIdx.x * AT_BM;
  // PDL: everything this kernel reads (lens, Q, K, V) comes from its predecessor, so the wait is the first thing it
  // does; what overlaps is the launch itself.  Its own dependents (the FFN GEMM: barriers, TMEM, weight block) may
  // start at once.
  tc::pdl_wait();
  tc::pdl_launch_dependents();
  // ragged batches: CTAs are handed out in blockIdx order, so the caller may pass the sequences sorted by key count,
  // longest first (a CTA's duration is proportional to it) -- the kernel's tail is then made of the shortest items
  const int s = order ? order[blockIdx.z] : (int)blockIdx.z;
  const int nq = lens ? lens[s] : Lp;
  if ((int)(blockIdx.x - crank) * AT_BM >= nq) return;  // whole cluster is past the valid rows
  const int skv = s ^ kv_xor;
  const int nk = lens ? lens[skv] : Lp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (nk + AT_BN - 1) / AT_BN;

  if (n_tiles == 0) {  // no keys: attention output is defined as zero (nan_to_num)
    if (warp >= 2 && warp < 6) {  // (uniform across the cluster: no barrier was touched yet)
      const int r = (warp & 3) * 32 + lane;
      if (q0 + r < nq) {
        uint4* dst = reinterpret_cast<uint4*>(ctx + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-byte aligned base
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;              // KST stages
  uint8_t* sV = smem + (1 + KST) * TILE_BYTES;  // VST stages
  uint8_t* sQx = smem + AT_XOFF;             // [128 queries][16] bf16: columns 0,1 = -m_ref(row) as hi, lo; rest 0
  uint8_t* sKx = sQx + XT_BYTES;             // [128 keys][16] bf16: 1, 1 at the start of each 16-byte chunk
  uint64_t* bars_base = reinterpret_cast<uint64_t*>(smem + AT_BAROFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_base + 2 * AT_NBAR);
  volatile int* panic = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_sum = reinterpret_cast<float*>(smem + AT_BAROFF + AT_BARBYTES);  // [4 slots][128 rows][NP parts]

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ);
    tc::prefetch_tmap(&tmK);
    tc::prefetch_tmap(&tmV);
    for (int set = 0; set < 2; ++set) {  // second set: the restart in exact mode (see `panic`) starts on fresh barriers
      uint64_t* b = bars_base + set * AT_NBAR;
      tc::mbar_init(b + 0, 1);                                                                   // q_full
      for (int i = 0; i < KST; ++i) { tc::mbar_init(b + 1 + i, 1); tc::mbar_init(b + 1 + KST + i, CL); }
      for (int i = 0; i < VST; ++i) { tc::mbar_init(b + 1 + 2 * KST + i, 1); tc::mbar_init(b + 1 + 2 * KST + VST + i, CL); }
      tc::mbar_init(b + 1 + 2 * KST + 2 * VST + 0, 1);  // s_full
      tc::mbar_init(b + 1 + 2 * KST + 2 * VST + 1, 4 * NP);  // s_free
      tc::mbar_init(b + 1 + 2 * KST + 2 * VST + 2, 4 * NP);  // p_ready
      tc::mbar_init(b + 1 + 2 * KST + 2 * VST + 3, 1);  // pv_done
    }
    *panic = 0;
    tc::fence_barrier_init();
  }
#if LG_ATTN_MSUB
  // S' = Q.K^T + Q_ext.K_ext^T = s - m_ref: the softmax loop then needs no subtraction (64 of its ~480 instructions
  // per 32x64 block; the loop is issue-bound).  Q_ext starts at 0 and is rewritten by the softmax threads whenever the
  // reference maximum of a row moves (first tile, then only when the maximum grows by > 8).  m_ref is kept exactly
  // representable as hi + lo of two bf16 (16 mantissa bits: < 1 unit of error up to |m| = 65 536 log2 units), carried
  // in Q_ext columns 0 and 1.  K_ext has ones in the first two elements of BOTH 16-byte chunks of a row, so the
  // product is -(hi + lo) whichever chunk the 32-byte swizzle maps Q_ext's non-zero pair to.
  if (threadIdx.x < 256) {
    reinterpret_cast<uint4*>(sQx)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(sKx)[threadIdx.x] = make_uint4(0x3f803f80u, 0u, 0u, 0u);  // two leading ones
  }
#endif
  for (int i = threadIdx.x; i < 4 * 128 * NP / 4; i += blockDim.x) reinterpret_cast<uint4*>(s_sum)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async();
  if (warp == 1) tc::tmem_alloc(tmem_slot, TM_COLS);
  if (threadIdx.x == 64 && stagger_ns > 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (atomicAdd(&g_attn_sm_slot[smid & 1023], 1u) & 1u) __nanosleep(stagger_ns);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc::cluster_sync();  // peers' barriers exist before anyone multicasts into them
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  XSTAMP(1);

  // Pass 0 runs with the DEFERRED row maximum (mode 0): after the first tile the softmax threads exponentiate a score
  // tile as soon as it is in registers, against the reference the tile was produced with, and look at the tile's row
  // sum afterwards; the reference moves (one tile later) when that sum exceeds 2^24.  A sum above 2^70 (or inf / NaN)
  // means the scores jumped by more than the fp32 range can absorb: the CTA sets `panic`, finishes the schedule and
  // runs the whole work item again in the exact mode (mode 1: maximum before the exponentials, every tile).
  // (Clusters with K/V multicast always run in mode 1: a restart would have to be agreed across the cluster.)
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
  const int mode = (CL > 1) ? 1 : (start_mode | pass);
  uint64_t* bars = bars_base + pass * AT_NBAR;
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;             // [KST]
  uint64_t* k_empty = k_full + KST;        // [KST]
  uint64_t* v_full = k_empty + KST;        // [VST]
  uint64_t* v_empty = v_full + VST;        // [VST]
  uint64_t* s_full = v_empty + VST;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_ready = s_free + 1;
  uint64_t* pv_done = p_ready + 1;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = (s * LG_HEADS + h) * Lp + q0;
      const int kvrow = (skv * LG_HEADS + h) * Lp;
      tc::mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tc::tma_load_2d(sQ, &tmQ, q_full, 0, qrow);
      for (int j = 0; j < n_tiles; ++j) {
        const int ks = j % KST, vs = j % VST;
        const int row = kvrow + j * AT_BN + (int)crank * SLICE;
        const int off = (int)crank * SLICE * 128;
        LG_PI_WAIT(&k_empty[ks], ((j / KST) & 1) ^ 1);  // every CTA of the cluster released the stage
        tc::mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
        if (CL > 1) tc::tma_load_2d_mc(sK + ks * TILE_BYTES + off, &tmK, &k_full[ks], 0, row, MC_MASK);
        else tc::tma_load_2d(sK + ks * TILE_BYTES, &tmK, &k_full[ks], 0, row);
        LG_PI_WAIT(&v_empty[vs], ((j / VST) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&v_full[vs], TILE_BYTES);
        if (CL > 1) tc::tma_load_2d_mc(sV + vs * TILE_BYTES + off, &tmV, &v_full[vs], 0, row, MC_MASK);
        else tc::tma_load_2d(sV + vs * TILE_BYTES, &tmV, &v_full[vs], 0, row);
      }
    }
  } else if (warp == 1) {
    // MMA issuer.  The WHOLE warp runs this loop with warp-uniform control flow and one elected lane
    // issues: descriptors and barrier addresses then live in uniform registers.  (The first version
    // ran the loop on lane 0 only; every tcgen05.mma needed R2UR moves and cost ~100 issue cycles,
    // 1190 of the 1730 cycles of a step.)
    constexpr uint32_t idesc_qk = tc::idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_pv = tc::idesc_bf16(128, 64, 1);
    const uint64_t dQ = tc::smem_desc_sw128(tc::smem_u32(sQ), 0, 1024);
    const uint64_t dK0 = tc::smem_desc_sw128(tc::smem_u32(sK), 0, 1024);
    const uint64_t dV0 = tc::smem_desc_sw128(tc::smem_u32(sV), TILE_BYTES, 1024);
    const uint32_t tS = tmem + TM_S, tP = tmem + TM_P, tO = tmem + TM_O;
#if LG_ATTN_MSUB
    const uint64_t dQx = tc::smem_desc_sw32(tc::smem_u32(sQx), 256);
    const uint64_t dKx = tc::smem_desc_sw32(tc::smem_u32(sKx), 256);
#endif
    int ks = 0, vs = 0;            // ring positions of the next K tile to multiply / V tile to consume
    uint32_t kph = 0, vph = 0;
    auto issue_qk = [&]() {        // S = Q . K[ks]^T
      const uint64_t dK = dK0 + (uint64_t)(ks * (TILE_BYTES >> 4));
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (dbg != 6) tc::umma_ss(tS, dQ + 2 * k, dK + 2 * k, idesc_qk, k != 0);
#if LG_ATTN_MSUB
        tc::umma_ss(tS, dQx, dKx, idesc_qk, 1);
#endif
        tc::umma_commit(s_full);
        if (CL > 1) tc::umma_commit_mc(&k_empty[ks], MC_MASK);  // K stage free once this QK^T retires
        else tc::umma_commit(&k_empty[ks]);
      }
      __syncwarp();
      if (++ks == KST) { ks = 0; kph ^= 1; }
    };
    LG_PI_WAIT(q_full, 0);
    LG_PI_WAIT(&k_full[0], 0);
    tc::fence_after_sync();
    issue_qk();
    const bool recm = (dbg_in & 16) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == LG_DBG_Z && lane == 0;
#define MSTAMP(k) do { if (recm && j < 16) g_attn_times[256 + (j * 16) + (k)] = clock64(); } while (0)
    for (int j = 0; j < n_tiles; ++j) {
      MSTAMP(0);
      if (j + 1 < n_tiles) {
        LG_PI_WAIT(&k_full[ks], kph);
        MSTAMP(1);
        LG_PI_WAIT(s_free, j & 1);  // softmax holds S(j) in registers
        tc::fence_after_sync();
        MSTAMP(2);
        issue_qk();
        MSTAMP(3);
      }
      LG_PI_WAIT(&v_full[vs], vph);
      LG_PI_WAIT(p_ready, j & 1);
      tc::fence_after_sync();
      MSTAMP(4);
      const uint64_t dV = dV0 + (uint64_t)(vs * (TILE_BYTES >> 4));
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)  // 16 keys per MMA: P columns k*8.., V rows k*16.. (2048 B)
          if (dbg != 5) tc::umma_ts(tO, tP + k * 8, dV + k * (2048 >> 4), idesc_pv, (j | k) != 0);
        if (CL > 1) tc::umma_commit_mc(&v_empty[vs], MC_MASK);
        else tc::umma_commit(&v_empty[vs]);
        tc::umma_commit(pv_done);
      }
      __syncwarp();
      if (++vs == VST) { vs = 0; vph ^= 1; }
      MSTAMP(5);
    }
  } else {
    // softmax: 4 * NP warps, NP threads per query row.  Warp (quarter, part) owns TMEM lanes quarter*32.. and key
    // columns part*COLS..+COLS of the score tile; the threads of a row talk through the shared-memory ring.
    constexpr int COLS = AT_BN / NP;   // score columns per thread
    constexpr int OC = LG_DH / NP;     // output columns per thread
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    // m_ref: the reference the running sum, O and (through Q_ext) the arriving score tiles are expressed in; both
    // threads of a row hold identical copies.
    // Deferred mode: the softmax warps are a serial chain per tile (score load -> exponentials -> P store) and every
    // instruction in it costs ~0.3 % of the kernel, so the bookkeeping is a dozen instructions inside the unrolled loop's
    // basic block: this thread's row sum of tile j-1 is published to shared memory (no barrier; before the p_ready(j)
    // arrive), both halves of tile j-3 are read back with one 8-byte load (the reader observed pv_done(j-2) in tile j-1,
    // the publisher's p_ready(j-2) arrive follows its store), and one compare + one warp vote after the loop decide
    // whether tile j+1 takes the rare path, where the reference moves up by floor(log2(sum)) -- unless it moved less than
    // four tiles ago (the sums still on their way were taken against the old reference) -- or, above 2^70 / inf / NaN,
    // the `bad` bit is set.  The polynomial lanes, whose exponent patch would wrap silently for inputs above 127, feed
    // a 3-input maximum that is checked once per tile.
    float m_ref = 0.f, l_part = 0.f, sum_m1 = 0.f, joint_next = 0.f;
    bool bad = false, move_next = false, any_next = false;
    int last_move = -4;
    float* const sum_row = s_sum + r * NP;  // [slot][row][part]
    auto row_total = [&](int slot) -> float {  // all parts of this row in a slot, summed in a fixed order
      if constexpr (NP == 2) {
        const float2 v = *reinterpret_cast<const float2*>(sum_row + slot * (128 * NP));
        return v.x + v.y;
      } else {
        const float4 v = *reinterpret_cast<const float4*>(sum_row + slot * (128 * NP));
        return (v.x + v.y) + (v.z + v.w);
      }
    };
#define LG_DEFERRED_BOOKKEEPING()                                                                                  \
    do { /* unconditional and branch-free: stays inside the unrolled loop's basic block; tiles < 0 read zeros */   \
      sum_row[((j + 3) & 3) * (128 * NP) + part] = sum_m1;                                           /* tile j-1 */ \
      joint_next = row_total((j + 1) & 3);                                                           /* tile j-3 */ \
      move_next = !(joint_next <= 0x1p24f);                                                                        \
    } while (0)
    const bool rec = (dbg_in & 16) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == LG_DBG_Z && warp == 2 && lane == 0;
#define STAMP(k) do { if (rec && j < 16) g_attn_times[(j * 16) + (k)] = clock64(); } while (0)
    bool sf_ok = false;
    for (int j = 0; j < n_tiles; ++j) {
      STAMP(0);
      bool move = move_next, any_move = any_next;  // deferred mode: decided in the previous tile
      const float joint = joint_next;
#if LG_ATTN_EARLY_POLL
      if (!sf_ok) tc::mbar_wait(s_full, j & 1);
#else
      tc::mbar_wait(s_full, j & 1);
#endif
      tc::fence_after_sync();
      STAMP(1);
      uint32_t sv[COLS];
#pragma unroll
      for (int c = 0; c < COLS; c += 32) tc::tmem_ld32(tmem + lane_base + TM_S + part * COLS + c, sv + c);
#if LG_ATTN_PRELOAD
      if (j > 0) {  // the previous step's P store: its completion wait and the p_ready hand-shake sit behind this load
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(p_ready);
      }
#endif
#if LG_ATTN_EARLY_POLL
      bool pv_ok = false;
#if !LG_ATTN_PRELOAD
      pv_ok = j > 0 && tc::mbar_test(pv_done, (j - 1) & 1);  // consumed after the exponentials
#endif
#endif
      tc::tmem_ld_wait();
      STAMP(2);
      const int valid = nk - j * AT_BN - part * COLS;  // valid keys among this thread's columns
      if (valid < COLS) {
#pragma unroll
        for (int i = 0; i < COLS; ++i) {
          if (i >= valid) sv[i] = 0xff800000u;  // -inf
        }
      }
      // Exact mode and first tile: this tile's row maximum (relative to m_ref; 0 on the first tile) decides, the
      // reference moves when it exceeds 8 (factor 256).
      float up = 0.f;
      if (mode != 0 || j == 0) {  // uniform over the CTA
        float mxs[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = __uint_as_float(sv[i]);
#pragma unroll
        for (int i = 4; i < COLS; ++i) mxs[i & 3] = fmaxf(mxs[i & 3], __uint_as_float(sv[i]));
        float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
        // exchange through ring slot j & 1: in the deferred mode (first tile only) slot 0 is next written by the
        // publishers of tile 0's sums, each into its own part, after everyone has read the maxima
        sum_row[(j & 1) * (128 * NP) + part] = mx;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NP) : "memory");
        if constexpr (NP == 2) {
          const float2 v = *reinterpret_cast<const float2*>(sum_row + (j & 1) * (128 * NP));
          mx = fmaxf(v.x, v.y);
        } else {
          const float4 v = *reinterpret_cast<const float4*>(sum_row + (j & 1) * (128 * NP));
          mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        }
#if !LG_ATTN_MSUB
        mx -= m_ref;  // (the scores arrive raw: relative to the reference here)
#endif
        move = j == 0 || mx > 8.f;
        up = mx;
        any_move = __any_sync(0xffffffffu, move);  // rare after the first tile
      }
      STAMP(4);
      float delta = 0.f, alpha = 1.f;
      if (any_move) {
        // The new reference is rounded to hi + lo bf16 so that the MMA subtracts exactly what this thread accounts
        // for; this tile was produced with the old one, so the difference is subtracted below.
        if (!(mode != 0 || j == 0) && move) {
          if (!(joint <= 0x1p70f)) { bad = true; move = false; }  // garbage ahead: the work item will be repeated
          else if (j - last_move < 4) move = false;
          else up = (float)((int)(__float_as_uint(joint) >> 23) - 127);
        }
        if (move) {
          last_move = j;
          const float m_want = m_ref + up;
          const __nv_bfloat16 hi = __float2bfloat16_rn(m_want);
          const __nv_bfloat16 lo = __float2bfloat16_rn(m_want - __bfloat162float(hi));
          const float m_abs = __bfloat162float(hi) + __bfloat162float(lo);
          delta = m_abs - m_ref;
          m_ref = m_abs;
#if LG_ATTN_MSUB
          if (part == 0) {  // row r of Q_ext: 32-byte rows, (-hi, -lo) in the first two elements of the row
            const uint32_t bits = ((uint32_t)(*reinterpret_cast<const unsigned short*>(&hi)) |
                                   ((uint32_t)(*reinterpret_cast<const unsigned short*>(&lo)) << 16)) ^ 0x80008000u;
            *reinterpret_cast<uint32_t*>(sQx + r * 32) = bits;
          }
#endif
        }
        tc::fence_proxy_async();  // generic-proxy write -> visible to the next QK^T (async proxy)
        alpha = j == 0 ? 0.f : ex2(-delta);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(s_free);  // S is in registers and Q_ext is up to date: QK^T(j+1) may go
      STAMP(3);
      float tile_sum, pmax = -INFINITY;  // pmax: largest input of a polynomial lane
      uint32_t pk[COLS / 2];
      if (any_move) {
        LG_DEFERRED_BOOKKEEPING();
        float rsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < COLS / 2; ++i) {
#if LG_ATTN_MSUB
          float p0 = __uint_as_float(sv[2 * i]) - delta, p1 = __uint_as_float(sv[2 * i + 1]) - delta;
#else
          float p0 = __uint_as_float(sv[2 * i]) - m_ref, p1 = __uint_as_float(sv[2 * i + 1]) - m_ref;
#endif
          if (LG_POLY_PAIR(i)) {
            pmax = max3(pmax, p0, p1);
            const float2 pp = ex2_poly2(make_float2(p0, p1));
            p0 = pp.x, p1 = pp.y;
          } else {
            p0 = ex2(p0);
            if (LG_POLY_HERE(i)) pmax = fmaxf(pmax, p1);
            p1 = LG_POLY_HERE(i) ? ex2_poly(p1) : ex2(p1);
          }
          rsum[i & 3] += p0 + p1;
          pk[i] = tc::pack_bf16(p0, p1);
        }
        tile_sum = (rsum[0] + rsum[1]) + (rsum[2] + rsum[3]);
      } else {
        // packed f32x2 row sum: one issue slot per pair instead of two (0.682 -> 0.675 ms)
        LG_DEFERRED_BOOKKEEPING();
        float pend = -INFINITY;
        bool have_pend = false;  // (compile-time after unrolling)
        float2 rs2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#if LG_ATTN_SPLIT_P
        // P leaves in two halves: the wait for P.V(j-1) (P buffer free) sits in the middle of the exponentials instead of
        // after them, and the first half's tcgen05.st is in flight during the second half's exponentials
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = half * (COLS / 4); i < (half + 1) * (COLS / 4); ++i) {
#else
        {
#pragma unroll
          for (int i = 0; i < COLS / 2; ++i) {
#endif
#if LG_ATTN_EARLY_POLL && LG_ATTN_PRELOAD
            if (i == LG_ATTN_PVTEST_AT) pv_ok = tc::mbar_test(pv_done, (j - 1) & 1);  // P.V(j-1) was released at the top of this step (j = 0: not consumed)
#endif
#if LG_ATTN_MSUB
            float p0 = __uint_as_float(sv[2 * i]), p1 = __uint_as_float(sv[2 * i + 1]);
#else  // no subtraction slice in the MMA: one packed add per pair (saves 1/9 of the MMA work, i.e. energy under the
       // power cap, for half an issue slot per element)
            const float2 q2 = __fadd2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])),
                                         make_float2(-m_ref, -m_ref));
            float p0 = q2.x, p1 = q2.y;
#endif
            if (LG_POLY_PAIR(i)) {
              pmax = max3(pmax, p0, p1);
              const float2 pp = ex2_poly2(make_float2(p0, p1));
              p0 = pp.x, p1 = pp.y;
            } else {
              p0 = ex2(p0);
              if (LG_POLY_HERE(i)) {  // two polynomial inputs per FMNMX3
                if (have_pend) { pmax = max3(pmax, pend, p1); have_pend = false; }
                else { pend = p1; have_pend = true; }
              }
              p1 = LG_POLY_HERE(i) ? ex2_poly(p1) : ex2(p1);
            }
            rs2[i & 1] = __fadd2_rn(rs2[i & 1], make_float2(p0, p1));
            pk[i] = tc::pack_bf16(p0, p1);
          }
#if LG_ATTN_SPLIT_P
          if (half == 0 && j > 0) {
            tc::mbar_wait(pv_done, (j - 1) & 1);  // PV(j-1) retired: P is free
            tc::fence_after_sync();
          }
          if constexpr (NP == 2) tc::tmem_st16(tmem + lane_base + TM_P + part * (COLS / 2) + half * (COLS / 4), pk + half * (COLS / 4));
          else tc::tmem_st8(tmem + lane_base + TM_P + part * (COLS / 2) + half * (COLS / 4), pk + half * (COLS / 4));
#endif
        }
        if (have_pend) pmax = fmaxf(pmax, pend);
        tile_sum = (rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y);
      }
      bad = bad || !(pmax <= 126.f);
      l_part = l_part * alpha + tile_sum;
      any_next = __any_sync(0xffffffffu, move_next);  // (exact mode: sums <= 128, never set)
      sum_m1 = tile_sum;
      STAMP(5);
#if LG_ATTN_EARLY_POLL
      sf_ok = tc::mbar_test(s_full, (j + 1) & 1);  // S(j+1) (never completes after the last tile: not consumed)
#endif
      if (!LG_ATTN_SPLIT_P || any_move) {
      if (j > 0) {
#if LG_ATTN_EARLY_POLL
        if (!pv_ok)
#endif
        tc::mbar_wait(pv_done, (j - 1) & 1);  // PV(j-1) retired: P is free, O is up to date
        tc::fence_after_sync();
        STAMP(6);
        if (any_move) {  // same rows in both half-warps -> same decision
          uint32_t o[OC];
          if constexpr (OC == 32) tc::tmem_ld32(tmem + lane_base + TM_O + part * OC, o);
          else tc::tmem_ld16(tmem + lane_base + TM_O + part * OC, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < OC; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          if constexpr (OC == 32) tc::tmem_st32(tmem + lane_base + TM_O + part * OC, o);
          else tc::tmem_st16(tmem + lane_base + TM_O + part * OC, o);
        }
      }
      if constexpr (NP == 2) tc::tmem_st32(tmem + lane_base + TM_P + part * (COLS / 2), pk);
      else tc::tmem_st16(tmem + lane_base + TM_P + part * (COLS / 2), pk);
      }
#if !LG_ATTN_PRELOAD
      tc::tmem_st_wait();
      STAMP(7);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_ready);
      STAMP(8);
#endif
    }
#if LG_ATTN_PRELOAD
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(p_ready);
#endif
    XSTAMP(2);
    // combine the partial row sums (ring slot n_tiles & 3 is the one no pending tile sum lives in; the exact mode, whose
    // maxima use slots 0 and 1, gets slot 2), normalise this thread's output columns
    const int fin_slot = mode == 0 ? (n_tiles & 3) : 2;
    sum_row[fin_slot * (128 * NP) + part] = l_part;
    if (mode == 0) sum_row[((n_tiles - 1) & 3) * (128 * NP) + part] = sum_m1;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NP) : "memory");
    const float l_sum = row_total(fin_slot);
    if (mode == 0) {  // the last three tiles' row sums have not been looked at yet
      bool ok = !bad;
#pragma unroll
      for (int t = 1; t <= 3; ++t) {
        if (n_tiles >= t) ok = ok && (row_total((n_tiles - t) & 3) <= 0x1p70f);
      }
      if (!ok) *panic = 1;
    }
    tc::mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc::fence_after_sync();
    const float inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
    uint32_t o[OC];
    if constexpr (OC == 32) tc::tmem_ld32(tmem + lane_base + TM_O + part * OC, o);
    else tc::tmem_ld16(tmem + lane_base + TM_O + part * OC, o);
    tc::tmem_ld_wait();
    if (q0 + r < nq) {
      uint4* dst = reinterpret_cast<uint4*>(ctx + ((size_t)s * Lp + q0 + r) * LG_D + h * LG_DH + part * OC);
#pragma unroll
      for (int i = 0; i < OC / 8; ++i) {
        uint4 w;
        w.x = tc::pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        w.y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        w.z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        w.w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        dst[i] = w;
      }
    }
  }
  XSTAMP(3);
  tc::fence_before_sync();
  __syncthreads();
  XSTAMP(4);
  if (mode != 0 || *panic == 0) break;
  // restart in the exact mode: every TMA load and MMA of pass 0 has been consumed (the softmax warps waited for the
  // last P.V); fresh barrier set, Q_ext back to zero, O is overwritten by the first P.V (accumulate flag)
  tc::fence_after_sync();
#if LG_ATTN_MSUB
  if (threadIdx.x < 256) reinterpret_cast<uint4*>(sQx)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
#endif
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  }  // pass
  if (CL > 1) tc::cluster_sync();  // nobody retires while a peer may still multicast into its smem
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
}

}  // namespace

template <int CL, int NP>
static int launch_attention(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                            const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, int dbg, const int32_t* order,
                            cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  const uint64_t d[2] = {64, (uint64_t)S * LG_HEADS * Lp}, sb[1] = {128};
  const uint32_t box[2] = {64, 128}, box_kv[2] = {64, 128 / CL};
  int rc;
  if ((rc = lg_make_tmap_bf16(&tq, Q, 2, d, sb, box))) return rc;
  if ((rc = lg_make_tmap_bf16(&tk, K, 2, d, sb, box_kv))) return rc;
  if ((rc = lg_make_tmap_bf16(&tv, V, 2, d, sb, box_kv))) return rc;
  auto kern = tc_attention_kernel<CL, NP>;
  const int smem = (dbg & 15) == 3 ? 120 * 1024 : at_smem(NP);  // dbg 3: one CTA per SM
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(Lp / AT_BM, LG_HEADS, S);
  cfg.blockDim = dim3(64 + 128 * NP);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  tc::lg_pdl_attr(&attr[1]);
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static const unsigned stagger = getenv("LGB200_ATTN_STAGGER_NS") ? (unsigned)atoi(getenv("LGB200_ATTN_STAGGER_NS")) : 0u;  // (mattered before the instruction diet; 0 .. 1800 ns now within 1.5 %)
  static const int start_mode = getenv("LGB200_ATTN_EXACT_MAX") ? atoi(getenv("LGB200_ATTN_EXACT_MAX")) : 0;  // 1: maximum before the exponentials on every tile (r1 behaviour)
  e = cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, Lp, lens, kv_xor, ctx, dbg, stagger, start_mode, order);
  if (e != cudaSuccess) return (int)e;
  LG_LAUNCH_CHECK();
  return LGB200_OK;
}

int lg_tc_attention(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                    const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, const int32_t* order, cudaStream_t st) {
  static const int dbg = getenv("LGB200_ATTN_DBG") ? atoi(getenv("LGB200_ATTN_DBG")) : 0;
  static const int force_cl = getenv("LGB200_ATTN_CL") ? atoi(getenv("LGB200_ATTN_CL")) : 0;
  const int qt = Lp / AT_BM;
  // Measured at S=128, Lp=2048: CL=1 0.84 ms, CL=2 0.87 ms, CL=4 0.93 ms -- the kernel is bound by the
  // softmax/MUFU side, not by L2->SM traffic, so K/V multicast stays opt-in (LGB200_ATTN_CL=2|4).
  int cl = 1;
  if (force_cl == 1 || force_cl == 2 || force_cl == 4) cl = (qt % force_cl == 0) ? force_cl : 1;
  static const int np = getenv("LGB200_ATTN_NP") ? atoi(getenv("LGB200_ATTN_NP")) : 2;  // softmax threads per query row
  if (cl == 4) return launch_attention<4, 2>(Q, K, V, S, Lp, lens, kv_xor, ctx, dbg, nullptr, st);  // (clusters: x only)
  if (cl == 2) return launch_attention<2, 2>(Q, K, V, S, Lp, lens, kv_xor, ctx, dbg, nullptr, st);
  if (np == 4) return launch_attention<1, 4>(Q, K, V, S, Lp, lens, kv_xor, ctx, dbg, order, st);
  return launch_attention<1, 2>(Q, K, V, S, Lp, lens, kv_xor, ctx, dbg, order, st);
}

extern "C" int lgb200_debug_attn_times(long long* host_out, int n) {
  if (n > 512) n = 512;
  return (int)cudaMemcpyFromSymbol(host_out, g_attn_times, sizeof(long long) * n);
}
