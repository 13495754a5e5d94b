// Temporary: tcgen05 attention / assignment not yet built.
#include "lg_internal.cuh"
int lg_tc_attention(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, int, int,
                    const int32_t*, int, __nv_bfloat16*, cudaStream_t) { return LGB200_ERR_PRECISION; }
int lg_tc_assign_lse(const __nv_bfloat16*, int, int, const int32_t*, float*, cudaStream_t) { return LGB200_ERR_PRECISION; }
int lg_tc_assign_scores(const __nv_bfloat16*, const float*, const float*, int, int, const int32_t*, int, int,
                        float*, cudaStream_t) { return LGB200_ERR_PRECISION; }
