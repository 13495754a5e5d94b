// extern "C" dispatch for the compute-bound entry points (include/lightglue_b200.h).
#include "lg_internal.cuh"
#include <stdlib.h>

extern "C" int lgb200_linear(int precision, int epilogue, const void* A0, const void* A1, int K0,
                             const void* W, const float* bias, int T, int N, int K,
                             const int32_t* lens, int Lp, float scale0, float scale1, float scale2,
                             const float* resid32, const void* resid16, float* out32, void* out16,
                             const float* rot, const void* rot16, int n_rot, void* outp0, void* outp1,
                             void* outp2, const float* gamma, const float* beta, void* stream) {
  if (!A0 || !W || !bias) return LGB200_ERR_NULL;
  if (T <= 0 || N <= 0 || K <= 0 || K0 <= 0 || K0 > K || Lp <= 0 || Lp % 128 || T % Lp)
    return LGB200_ERR_SHAPE;
  if (K0 < K && !A1) return LGB200_ERR_NULL;
  if (resid32 && resid16) return LGB200_ERR_SHAPE;
  if ((precision == LGB200_F32 || precision == LGB200_F32X3) && (resid16 || (rot16 && !rot))) return LGB200_ERR_PRECISION;
  LgEpi e;
  e.mode = epilogue;
  e.N = N;
  e.Lp = Lp;
  e.bias = bias;
  e.scale[0] = scale0; e.scale[1] = scale1; e.scale[2] = scale2;
  e.resid32 = resid32;
  e.out32 = out32;
  e.out16 = reinterpret_cast<__nv_bfloat16*>(out16);
  e.rot = rot;
  e.n_rot = n_rot;
  e.outp[0] = outp0; e.outp[1] = outp1; e.outp[2] = outp2;
  e.gamma = gamma;
  e.beta = beta;
  switch (epilogue) {
    case LGB200_EPI_ROWMAJOR:
      if (!out32 && !out16) return LGB200_ERR_NULL;
      break;
    case LGB200_EPI_HEADS: {
      if (N % 256 || N / 256 > 3) return LGB200_ERR_SHAPE;
      for (int p = 0; p < N / 256; ++p)
        if (!e.outp[p]) return LGB200_ERR_NULL;
      if (n_rot > 0 && !rot && !rot16) return LGB200_ERR_NULL;
      break;
    }
    case LGB200_EPI_LN_GELU:
      if (N != 512) return LGB200_ERR_SHAPE;
      if (!gamma || !beta) return LGB200_ERR_NULL;
      break;
    default:
      return LGB200_ERR_PRECISION;
  }
  cudaStream_t st = lg_stream(stream);
  if (precision == LGB200_F32)
    return lg_simt_linear(epilogue, (const float*)A0, (const float*)A1, K0, (const float*)W, T, N, K,
                          lens, e, st);
  if (precision == LGB200_F32X3) {  // operands and out16 / outp are split-fp16 planes
    e.out16 = nullptr;
    return lg_x3_linear(epilogue, A0, A1, K0, W, T, N, K, lens, e, out16, st);
  }
  if (precision == LGB200_BF16) {
    // v2 = weight-stationary / cluster-multicast kernel (lg_tc_gemm2.cu); LGB200_GEMM_V1=1 selects the
    // first-generation streaming kernel (lg_tc_gemm.cu) for A/B measurements.
    static const bool use_v1 = getenv("LGB200_GEMM_V1") != nullptr;
    if (!use_v1) {
      const int rc = lg_tc_linear_v2(epilogue, (const __nv_bfloat16*)A0, (const __nv_bfloat16*)A1, K0,
                                     (const __nv_bfloat16*)W, T, N, K, lens, e, rot16,
                                     (const __nv_bfloat16*)resid16, st);
      if (rc != LGB200_ERR_SHAPE) return rc;  // v2 covers bf16-in/bf16-out shapes; the rest goes to v1
    }
    if (resid16 || (n_rot > 0 && !rot)) return LGB200_ERR_SHAPE;  // v1 takes fp32 side inputs only
    return lg_tc_linear(epilogue, (const __nv_bfloat16*)A0, (const __nv_bfloat16*)A1, K0,
                        (const __nv_bfloat16*)W, T, N, K, lens, e, st);
  }
  return LGB200_ERR_PRECISION;
}

extern "C" int lgb200_attention_ordered(int precision, const void* Q, const void* K, const void* V, int S, int Lp,
                                        const int32_t* lens, const int32_t* order, int kv_xor, void* ctx,
                                        void* stream);
extern "C" int lgb200_attention(int precision, const void* Q, const void* K, const void* V, int S,
                                int Lp, const int32_t* lens, int kv_xor, void* ctx, void* stream) {
  return lgb200_attention_ordered(precision, Q, K, V, S, Lp, lens, nullptr, kv_xor, ctx, stream);
}

extern "C" int lgb200_attention_ordered(int precision, const void* Q, const void* K, const void* V, int S, int Lp,
                                        const int32_t* lens, const int32_t* order, int kv_xor, void* ctx,
                                        void* stream) {
  if (!Q || !K || !V || !ctx) return LGB200_ERR_NULL;
  if (S <= 0 || Lp <= 0 || Lp % 128 || (kv_xor != 0 && kv_xor != 1) || (kv_xor && (S & 1)))
    return LGB200_ERR_SHAPE;
  cudaStream_t st = lg_stream(stream);
  if (precision == LGB200_F32)
    return lg_simt_attention((const float*)Q, (const float*)K, (const float*)V, S, Lp, lens, kv_xor,
                             (float*)ctx, st);
  if (precision == LGB200_F32X3) return lg_x3_attention(Q, K, V, S, Lp, lens, kv_xor, ctx, st);
  if (precision == LGB200_BF16) {
    // LGB200_ATTN2=1 selects the two-query-tile kernel with ordered softmax groups (lg_tc_attn2.cu)
    static const int attn2 = getenv("LGB200_ATTN2") ? atoi(getenv("LGB200_ATTN2")) : 0;
    if (attn2)
      return lg_tc_attention2((const __nv_bfloat16*)Q, (const __nv_bfloat16*)K, (const __nv_bfloat16*)V, S, Lp,
                              lens, kv_xor, (__nv_bfloat16*)ctx, st);
    return lg_tc_attention((const __nv_bfloat16*)Q, (const __nv_bfloat16*)K, (const __nv_bfloat16*)V, S,
                           Lp, lens, kv_xor, (__nv_bfloat16*)ctx, order, st);
  }
  return LGB200_ERR_PRECISION;
}

extern "C" int lgb200_assign_lse(int precision, const void* md, int S, int Lp, const int32_t* lens,
                                 float* lse, void* stream) {
  if (!md || !lse) return LGB200_ERR_NULL;
  if (S <= 0 || (S & 1) || Lp <= 0 || Lp % 128) return LGB200_ERR_SHAPE;
  cudaStream_t st = lg_stream(stream);
  if (precision == LGB200_F32) return lg_simt_assign_lse((const float*)md, S, Lp, lens, lse, st);
  if (precision == LGB200_BF16)
    return lg_tc_assign_lse((const __nv_bfloat16*)md, S, Lp, lens, lse, st);
  return LGB200_ERR_PRECISION;
}

extern "C" int lgb200_assign_scores(int precision, const void* md, const float* z, const float* lse,
                                    int B, int Lp, const int32_t* lens, int R, int C, float* scores,
                                    void* best_ws, void* stream) {
  if (!md || !z || !lse || !scores) return LGB200_ERR_NULL;
  if (B <= 0 || Lp <= 0 || Lp % 128 || R < 1 || C < 1 || R - 1 > Lp || C - 1 > Lp)
    return LGB200_ERR_SHAPE;
  cudaStream_t st = lg_stream(stream);
  if (precision == LGB200_F32) {
    if (best_ws) return LGB200_ERR_PRECISION;  // the fused argmax exists in the tcgen05 kernel only
    return lg_simt_assign_scores((const float*)md, z, lse, B, Lp, lens, R, C, scores, nullptr, st);
  }
  if (precision == LGB200_BF16)
    return lg_tc_assign_scores((const __nv_bfloat16*)md, z, lse, B, Lp, lens, R, C, scores, best_ws, st);
  return LGB200_ERR_PRECISION;
}

extern "C" int lgb200_assign_loss(int precision, const void* md, const float* z, const float* lse, int B, int Lp,
                                  const int32_t* lens, int R, int C, const uint8_t* gt_assignment, float* row_pos,
                                  float* row_cnt, float* row_exp, int32_t* row_arg, int32_t* col_arg, void* workspace,
                                  void* stream) {
  if (!md || !z || !lse || !gt_assignment || !row_pos || !row_cnt || !row_exp || !workspace) return LGB200_ERR_NULL;
  if (B <= 0 || Lp <= 0 || Lp % 128 || R < 1 || C < 1 || R - 1 > Lp || C - 1 > Lp) return LGB200_ERR_SHAPE;
  if (precision != LGB200_BF16) return LGB200_ERR_PRECISION;  // fp32: lgb200_assign_scores + lgb200_loss_reduce
  return lg_tc_assign_loss((const __nv_bfloat16*)md, z, lse, B, Lp, lens, R, C, gt_assignment, row_pos, row_cnt,
                           row_exp, row_arg, col_arg, workspace, lg_stream(stream));
}
