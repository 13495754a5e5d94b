// Internal (non-ABI) entry points shared between translation units.
#pragma once
#include "lg_common.cuh"

// fp32 CUDA-core path (lg_simt.cu)
int lg_simt_linear(int epilogue, const float* A0, const float* A1, int K0, const float* W, int T, int N,
                   int K, const int32_t* lens, LgEpi epi, cudaStream_t st);
int lg_simt_attention(const float* Q, const float* K, const float* V, int S, int Lp, const int32_t* lens,
                      int kv_xor, float* ctx, cudaStream_t st);
int lg_simt_assign_lse(const float* md, int S, int Lp, const int32_t* lens, float* lse, cudaStream_t st);
int lg_simt_assign_scores(const float* md, const float* z, const float* lse, int B, int Lp,
                          const int32_t* lens, int R, int C, float* scores, float* sim_out, cudaStream_t st);

// bf16 tcgen05 path (lg_tc_*.cu)
int lg_tc_linear(int epilogue, const __nv_bfloat16* A0, const __nv_bfloat16* A1, int K0,
                 const __nv_bfloat16* W, int T, int N, int K, const int32_t* lens, LgEpi epi,
                 cudaStream_t st);
int lg_tc_linear_v2(int epilogue, const __nv_bfloat16* A0, const __nv_bfloat16* A1, int K0,
                    const __nv_bfloat16* W, int T, int N, int K, const int32_t* lens, LgEpi epi,
                    const void* rot16, const __nv_bfloat16* resid16, cudaStream_t st);
int lg_tc_attention(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                    const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, const int32_t* order, cudaStream_t st);
int lg_tc_attention2(const __nv_bfloat16* Q, const __nv_bfloat16* K, const __nv_bfloat16* V, int S, int Lp,
                     const int32_t* lens, int kv_xor, __nv_bfloat16* ctx, cudaStream_t st);
int lg_tc_assign_lse(const __nv_bfloat16* md, int S, int Lp, const int32_t* lens, float* lse,
                     cudaStream_t st);
int lg_tc_assign_scores(const __nv_bfloat16* md, const float* z, const float* lse, int B, int Lp,
                        const int32_t* lens, int R, int C, float* scores, void* best_ws, cudaStream_t st);
int lg_tc_assign_loss(const __nv_bfloat16* md, const float* z, const float* lse, int B, int Lp, const int32_t* lens,
                      int R, int C, const uint8_t* gt, float* row_pos, float* row_cnt, float* row_exp,
                      int32_t* row_arg, int32_t* col_arg, void* best_ws, cudaStream_t st);
