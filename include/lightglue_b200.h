/*
 * lightglue_b200.h -- C ABI of the B200-native LightGlue matcher hot path.
 *
 * The reference (ipastore/glue-factory-colon) has NO native boundary: every op
 * of gluefactory/models/matchers/lightglue.py is a PyTorch library call.  Each
 * entry point below replaces the library call sites of one kernel family; the
 * reference file:line it replaces is cited per function (paths relative to the
 * reference checkout).  The only caller is the Python host mirror
 * glue_factory_colon_b200/lightglue.py (ctypes).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless stated otherwise.  No entry point
 *     allocates, synchronises or keeps mutable global state; every launch goes
 *     to the `stream` argument (a cudaStream_t passed as void*).
 *   - Return value: 0 on success; LGB200_ERR_* (negative) for bad arguments;
 *     a positive value is a cudaError_t from the launch.
 *   - Token storage ("sequence-major"): S = 2*B sequences, sequence s = 2*b+i
 *     is image i of pair b; every sequence owns Lp rows (Lp % 128 == 0) of
 *     which the first lens[s] are valid.  `lens` is a DEVICE int32[S] so that
 *     point pruning can shrink it without a host round trip; a sequence with
 *     lens[s] == 0 is skipped by every kernel (used for early-exited pairs).
 *   - precision: LGB200_F32 runs hand-written CUDA-core fp32 kernels (the
 *     parity mode, 1e-3 on log_assignment); LGB200_BF16 runs tcgen05/TMEM/TMA
 *     kernels with bf16 operands and fp32 accumulation (the throughput mode);
 *     LGB200_F32X3 is the fp32-accurate mode ON the tensor cores: every MMA
 *     operand is a pair of fp16 planes [2][rows][K] (x = hi + lo, 22 mantissa
 *     bits), every product is three tcgen05 MMAs with fp32 accumulation,
 *     softmax / LayerNorm / GELU(erf) / residuals stay in fp32 (lg_x3*.cu).
 *     In that mode the operand pointers of lgb200_linear (A0, A1, W, out16,
 *     outp0..2) and of lgb200_attention (Q, K, V, ctx) are split planes.
 */
#ifndef LIGHTGLUE_B200_H_
#define LIGHTGLUE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGB200_ABI_VERSION 6

enum { LGB200_F32 = 0, LGB200_BF16 = 1, LGB200_F32X3 = 2 };

enum {
  LGB200_OK = 0,
  LGB200_ERR_SHAPE = -1,     /* unsupported extent / alignment            */
  LGB200_ERR_NULL = -2,      /* required pointer missing                  */
  LGB200_ERR_PRECISION = -3, /* unknown precision / epilogue enum         */
  LGB200_ERR_DRIVER = -4,    /* cuTensorMapEncodeTiled unavailable/failed */
  LGB200_ERR_ARCH = -5       /* device is not sm_100                      */
};

/* epilogues of lgb200_linear */
enum {
  LGB200_EPI_ROWMAJOR = 0, /* out[r,c] = (acc + bias[c]) * scale0 (+ resid[r,c])            */
  LGB200_EPI_HEADS = 1,    /* split into parts of 256 cols, optional rotary, head-major out */
  LGB200_EPI_LN_GELU = 2   /* N == 512: gelu(layernorm(acc + bias) * gamma + beta)          */
};

int lgb200_abi_version(void);
/* 0 if the current device can run the kernels (compute capability 10.x). */
int lgb200_device_ok(void);
const char* lgb200_error_string(int code);
/* Number of kernels this library has launched in the process so far (statistics for bench.py). */
unsigned long long lgb200_launch_count(void);

/* ---- input staging ------------------------------------------------------
 * Replaces descriptors.contiguous() + the implicit [B,N,d] layout,
 * lightglue.py:456-465 (input_proj itself goes through lgb200_linear).
 * src [B, n, dim] fp32 -> rows of sequences (2*b + img): x32 [S,Lp,dim] fp32 and/or
 * x16 (bf16); either may be NULL, not both.  Rows >= n are zero-filled. */
int lgb200_pack_rows(const float* src, int B, int n, int dim, int img, int Lp,
                     float* x32, void* x16, void* stream);

/* ---- positional encoding --------------------------------------------------
 * Replaces normalize_keypoints (lightglue.py:28-40) and
 * LearnableFourierPositionalEncoding.forward (lightglue.py:61-66).
 * kpts [B,n,kdim] fp32 (kdim 2, or 4 with scale/orientation appended),
 * size [B,2] (W,H) or NULL (extent of the valid points), Wr [32,kdim].
 * rot [S,Lp,64] fp32 receives (cos_f, sin_f) pairs, f = 0..31, for sequence
 * 2*b+img; one table serves the 4 heads and all layers (the reference's
 * repeat_interleave(2) is implicit: rotary pair f uses entry f).
 * rot16 [S,Lp,32] (nullable) receives the same pairs as packed fp16 (cos, sin):
 * the 128-byte-per-token table the bf16 QKV epilogue fetches by TMA.
 * At least one of rot / rot16 must be given. */
int lgb200_posenc(const float* kpts, int B, int n, int kdim, const float* size,
                  const float* Wr, const int32_t* lens, int img, int Lp,
                  float* rot, void* rot16, void* stream);

/* ---- linear layers with fused epilogues -------------------------------------
 * Replaces nn.Linear + the elementwise ops around it:
 *   Wqkv + unflatten + rotary   lightglue.py:157-161, 43-50
 *   to_qk / to_v                lightglue.py:196-201, 208
 *   out_proj / to_out           lightglue.py:163, 219
 *   ffn (cat, Linear, LayerNorm, GELU, Linear, residual)  lightglue.py:144-149,164,220-221
 *   final_proj / d**.25         lightglue.py:281-283
 *   input_proj                  lightglue.py:464-465
 * Y[T,N] = A[T,K] . W[N,K]^T, T = S*Lp.  A is A0[T,K0] followed column-wise by
 * A1[T,K-K0] (A1 NULL when K0 == K) -- this is the ffn's torch.cat([x,msg],-1).
 * A and W are fp32 (LGB200_F32) or bf16 (LGB200_BF16); bias/gamma/beta fp32.
 *   ROWMAJOR: out32 (fp32, nullable) / out16 (bf16, nullable), leading dim N;
 *             residual: resid32 [T,N] fp32 or resid16 [T,N] bf16 (at most one).
 *             The bf16 throughput path keeps the residual stream in bf16 only
 *             (resid16 == out16 updates it in place).
 *   HEADS:    N = parts*256, column c = part*256 + head*64 + d (the caller
 *             permutes Wqkv rows from the reference's head*192 + d*3 + part).
 *             Parts < n_rot get the rotary embedding from rot [T,64] fp32 or
 *             rot16 [T,32] packed fp16 (cos,sin) (LGB200_BF16 prefers rot16);
 *             part p is scaled by scale[p] and written to outp[p] laid out
 *             [S,4,Lp,64] in the precision's element type.
 *   LN_GELU:  N = 512; out as ROWMAJOR. */
int lgb200_linear(int precision, int epilogue, const void* A0, const void* A1, int K0,
                  const void* W, const float* bias, int T, int N, int K,
                  const int32_t* lens, int Lp,
                  float scale0, float scale1, float scale2,
                  const float* resid32, const void* resid16, float* out32, void* out16,
                  const float* rot, const void* rot16, int n_rot,
                  void* outp0, void* outp1, void* outp2,
                  const float* gamma, const float* beta, void* stream);

/* ---- attention ---------------------------------------------------------------
 * Replaces Attention.forward / F.scaled_dot_product_attention and the einsum
 * cross attention, lightglue.py:112-129, 203-217.  Q,K,V [S,4,Lp,64]; Q is
 * expected pre-scaled so that softmax uses exp2 (log2(e)/sqrt(64) folded in by
 * the producing lgb200_linear).  Sequence s attends to sequence s ^ kv_xor
 * (0 = self attention, 1 = the other image of the pair = cross attention);
 * keys >= lens[s ^ kv_xor] are masked, a sequence with no keys yields zeros
 * (nan_to_num, lightglue.py:118,217).  ctx [S,Lp,256] token-major, column
 * head*64+d (= transpose(1,2).flatten(-2), lightglue.py:163,218). */
int lgb200_attention(int precision, const void* Q, const void* K, const void* V,
                     int S, int Lp, const int32_t* lens, int kv_xor,
                     void* ctx, void* stream);
/* Same, with a launch order for ragged batches: order[S] (DEVICE int32, a permutation of 0..S-1, or NULL) lists the
 * sequences in the order their query tiles are handed to the SMs.  Sorted by key count, longest first, the kernel's
 * tail consists of the shortest work items (a CTA's duration is proportional to lens[s ^ kv_xor]).  The result does
 * not depend on it; the fp32 kernels ignore it. */
int lgb200_attention_ordered(int precision, const void* Q, const void* K, const void* V,
                             int S, int Lp, const int32_t* lens, const int32_t* order, int kv_xor,
                             void* ctx, void* stream);

/* ---- per-token heads -----------------------------------------------------------
 * Replaces matchability / token-confidence Linear(256,1) (+ sigmoid),
 * lightglue.py:72,75-80,285-286,290-291.  out[r] = dot(x[r,:], w) + b with x
 * [S*Lp,256] fp32 (LGB200_F32) or bf16 (LGB200_BF16), w/b fp32; sigmoid applied
 * when apply_sigmoid != 0; rows >= lens[s] are left untouched. */
int lgb200_rowdot(int precision, const void* x, const float* w, const float* b, int S, int Lp,
                  const int32_t* lens, int apply_sigmoid, float* out, void* stream);

/* ---- log assignment ---------------------------------------------------------------
 * Replaces the similarity einsum + sigmoid_log_double_softmax,
 * lightglue.py:257-269, 284-288.  md [S,Lp,256] = final_proj(desc)/4 (from
 * lgb200_linear), z [S,Lp] = matchability logits.
 *   pass 1 (lgb200_assign_lse): lse[s,l] = logsumexp_j <md[s,l], md[s^1,j]>,
 *          j < lens[s^1]  (row normaliser of image 0, column normaliser of image 1);
 *   pass 2 (lgb200_assign_scores): recomputes the similarity tile and writes
 *          scores [B,R,C] fp32 (R = n0max+1, C = n1max+1) exactly once:
 *          valid block 2*sim - lse0[i] - lse1[j] + logsigmoid(z0[i]) + logsigmoid(z1[j]),
 *          dustbin column C-1 = logsigmoid(-z0), dustbin row R-1 = logsigmoid(-z1),
 *          everything else (padding, corner) 0.
 *          best_ws (nullable, 8*B*(R+C) bytes, LGB200_BF16 only): the kernel also leaves the packed
 *          row / column maxima of the valid block there (computed on exactly the values it writes),
 *          which lgb200_filter_matches accepts with workspace_has_best = 1 and then does not
 *          re-read the score matrix at all. */
int lgb200_assign_lse(int precision, const void* md, int S, int Lp, const int32_t* lens,
                      float* lse, void* stream);
int lgb200_assign_scores(int precision, const void* md, const float* z, const float* lse,
                         int B, int Lp, const int32_t* lens, int R, int C,
                         float* scores, void* best_ws, void* stream);

/* ---- filter_matches -------------------------------------------------------------------
 * Replaces filter_matches, lightglue.py:294-319 (and the scatter back to
 * un-pruned indices, lightglue.py:527-536).  scores [B,R,C] fp32; pair b uses
 * rows < n0 = lens[2b], cols < n1 = lens[2b+1] (lens NULL: n0 = R-1, n1 = C-1).
 * Ties resolve to the lowest index, NaN counts as the maximum (torch.max).
 * ind0/ind1 (int32 [B,ind_ld], nullable) map current rows to original indices;
 * outputs are indexed by ORIGINAL index: m0 [B,N0] int64, m1 [B,N1] int64,
 * ms0 [B,N0], ms1 [B,N1] fp32; untouched entries are -1 / 0.
 * workspace: 8 * B * (R + C) bytes; workspace_has_best != 0 means it already holds the packed maxima
 * written by lgb200_assign_scores (scores may then be NULL). */
int lgb200_filter_matches(const float* scores, int B, int R, int C, const int32_t* lens,
                          float threshold, const int32_t* ind0, const int32_t* ind1, int ind_ld,
                          int N0, int N1, int64_t* m0, int64_t* m1, float* ms0, float* ms1,
                          void* workspace, int workspace_has_best, void* stream);

/* ---- nearest-neighbour matcher ----------------------------------------------------------
 * Replaces NearestNeighborMatcher._forward, gluefactory/models/matchers/nearest_neighbor_matcher.py:63-83
 * (fp32).  d [S,Lp,256] = descriptors packed like the LightGlue activations (rows >= count and columns
 * >= the descriptor dimension zero), lse from lgb200_assign_lse(LGB200_F32, d, ...).
 *   lgb200_nn_scores: similarity [B,R-1,C-1] (nullable) = <d0[i], d1[j]> (:64) and log_assignment [B,R,C] =
 *          log_softmax(sim, rows) + log_softmax(sim, columns) on the valid block, zeros elsewhere incl. the
 *          dustbin row / column (:72-74).
 *   lgb200_nn_match: find_nn (:15-31) for both images -- nearest neighbour by similarity, rejected if
 *          2(1-s1) > ratio^2 * 2(1-s2) (ratio_thresh > 0, more than one candidate) or 2(1-s1) > distance_thresh^2
 *          (distance_thresh > 0); lowest index among equal similarities -- then mutual_check (:34-43) if
 *          `mutual`; scores = (match > -1).  similarity [B,N,M]; workspace: 8 * B * (N + M) bytes. */
int lgb200_nn_scores(const float* d, const float* lse, int B, int Lp, const int32_t* lens, int R, int C,
                     float* similarity, float* log_assignment, void* stream);
int lgb200_nn_match(const float* similarity, int B, int N, int M, const int32_t* lens, float ratio_thresh,
                    float distance_thresh, int mutual, int64_t* workspace, int64_t* m0, int64_t* m1,
                    float* ms0, float* ms1, void* stream);

/* N_pair loss of the nearest-neighbour matcher, forward values.  Replaces
 * NearestNeighborMatcher.loss, nearest_neighbor_matcher.py:85-109:
 * scores = T (2 - sqrt(clamp(2 (1 - sim), 1e-6))), prob0 / prob1 = row / column
 * log-softmax of scores.  similarity [B,N,M] fp32, gt_assignment [B,N,M] bytes
 * (non-zero = match).  Outputs per row: row_sum[b,i] = sum_j a (prob0 + prob1),
 * row_cnt[b,i] = sum_j a; the caller forms nll = -sum_i row_sum / (2 max(sum_i row_cnt, 1)).
 * workspace: B * (N + M) floats. */
int lgb200_npair_loss(const float* similarity, const uint8_t* gt_assignment, int B, int N, int M,
                      float temperature, float* workspace, float* row_sum, float* row_cnt, void* stream);

/* ---- loss-side reductions (forward values only) -------------------
 * Replaces the dense [B,R,C] reductions behind LightGlue.loss, lightglue.py:588-637:
 *   NLLLoss / weight_loss, gluefactory/models/utils/losses.py:6-26 -- row_pos[b,i] = sum_j la[b,i,j] *
 *          gt[b,i,j] and row_cnt[b,i] = sum_j gt[b,i,j] over the inner block (i < R-1, j < C-1);
 *   losses["row_norm"], lightglue.py:606 -- row_exp[b,i] = sum_{j<C} exp(la[b,i,j]);
 *   TokenConfidence.loss, lightglue.py:86-91 -- row_arg[b,i] = argmax_{j<C} la[b,i,j] for i < R-1 and
 *          col_arg[b,j] = argmax_{i<R} la[b,i,j] for j < C-1 (torch.max: lowest index among equal values).
 * gt_assignment is the reference's bool tensor data["gt_assignment"] [B,R-1,C-1] (one byte per entry).
 * Any output pointer may be NULL; per-row results are written without atomics (bit-reproducible). */
int lgb200_loss_reduce(const float* log_assignment, int B, int R, int C, const uint8_t* gt_assignment,
                       float* row_pos, float* row_cnt, float* row_exp, int32_t* row_arg, int32_t* col_arg,
                       void* stream);

/* Fused form for LGB200_BF16 (tcgen05): MatchAssignment pass 2 of lgb200_assign_scores with the reductions above in
 * its epilogue instead of the store -- the [B,R,C] matrix of a layer that only the loss looks at (every layer but
 * the last in training mode, lightglue.py:607-608) is never written.  Same outputs as lgb200_loss_reduce on the
 * matrix lgb200_assign_scores would have produced; md / z / lse as for lgb200_assign_scores; lens entries must be
 * R-1 / C-1 (or lens NULL when R-1 == C-1 == Lp).  workspace: 8 * B * (R + C) bytes.  LGB200_F32 is refused
 * (LGB200_ERR_PRECISION): the fp32 parity mode materialises and calls lgb200_loss_reduce. */
int lgb200_assign_loss(int precision, const void* md, const float* z, const float* lse, int B, int Lp,
                       const int32_t* lens, int R, int C, const uint8_t* gt_assignment, float* row_pos,
                       float* row_cnt, float* row_exp, int32_t* row_arg, int32_t* col_arg, void* workspace,
                       void* stream);

/* ---- adaptive depth (early exit) ---------------------------------------------------------
 * Replaces check_if_stop, lightglue.py:569-580.  conf [S,Lp] = sigmoid token
 * confidences; pair b stops iff 1 - count(conf < thr)/total[b] > depth_conf
 * (fp32 arithmetic as in the reference).  done[b] is set to layer+1 for pairs
 * that stop now (pairs with done[b] != 0 are left alone) and their entries of
 * `lens_active` are zeroed so later kernels skip them. */
int lgb200_exit_check(const float* conf, int B, int Lp, const int32_t* lens,
                      const int32_t* total, float thr, float depth_conf, int layer,
                      int32_t* done, int32_t* lens_active, void* stream);

/* ---- adaptive width (point pruning) --------------------------------------------------------
 * Replaces get_pruning_mask + torch.where + index_select + prune += 1,
 * lightglue.py:506-521, 560-567.  keep[l] = match[s,l] > 1 - width_conf ||
 * (conf != NULL && conf[s,l] <= thr).  Rows are compacted from the *_src
 * buffers into the *_dst buffers (stable order), lens[s] is updated on the
 * device, prune_cnt[s, ind] += 1 for kept rows.  Sequences with
 * lens_active[s] == 0 are copied through unchanged in count (not pruned).
 * Each of the x32 / x16 / rot / rot16 buffer pairs may be NULL (not in use). */
int lgb200_prune_compact(const float* match, const float* conf, float thr, float width_conf,
                         int S, int Lp, int32_t* lens, int32_t* lens_active,
                         const float* x32_src, float* x32_dst,
                         const void* x16_src, void* x16_dst,
                         const float* rot_src, float* rot_dst,
                         const void* rot16_src, void* rot16_dst,
                         const int32_t* ind_src, int32_t* ind_dst,
                         int32_t* prune_cnt, void* stream);

/* ---- fp32-accurate tensor-core mode (LGB200_F32X3): helpers -------------------------------------
 * lgb200_split_rows: x [n] fp32 -> xs [2][n] fp16 planes (hi = fp16(x), lo = fp16(x - hi)); n % 4 == 0.
 * Used for the staged descriptors (lightglue.py:456-465) and after point pruning (:506-521). */
int lgb200_split_rows(const float* x, long long n, void* xs, void* stream);
/* lgb200_split_dynamic: the same for tensors without a fixed range (gradients): xs [2][n] = planes of g x with g the
 * power of two that brings max |x| into [256, 512) (1 for an all-zero tensor); inv_scale[0] = 1 / g on the device,
 * inv_scale[1] is scratch (two floats).  Feeds the three-product tensor-core GEMMs of the backward pass (train.py).
 * colsum_partials (nullable): x is read as [n / cols][cols] and colsum_partials [n_partials][cols] receives per-CTA
 * column sums from the same pass (the caller adds the n_partials rows: the bias gradient dY^T 1); cols % 4 == 0.
 * lgb200_merge_rows: planes [2][n] -> x [n] = (hi + lo) * inv_scale. */
int lgb200_split_dynamic(const float* x, long long n, void* xs, float* inv_scale, int cols, float* colsum_partials,
                         int n_partials, void* stream);
int lgb200_merge_rows(const void* xs, long long n, float inv_scale, float* x, void* stream);
/* Similarity of MatchAssignment (the einsum of lightglue.py:284) for all pairs: mds = split planes of
 * final_proj(desc)/4 [2][S*Lp][256]; sim [B][Lp][Lp] fp32 receives <md[2b,i], md[2b+1,j]> for i < lens[2b],
 * j < lens[2b+1] (tiles past the valid counts are not touched). Lp % 256 == 0. */
int lgb200_x3_similarity(const void* mds, int B, int Lp, const int32_t* lens, float* sim, void* stream);
/* lse [S*Lp]: row normaliser log_softmax(sim, 2) for image 0, column normaliser log_softmax(sim, 1) for image 1
 * (lightglue.py:262-263), natural log; n0 / n1 = counts when lens is NULL. */
int lgb200_x3_assign_lse(const float* sim, int B, int Lp, const int32_t* lens, int n0, int n1, float* lse,
                         void* stream);
/* scores [B][R][C] (lightglue.py:257-269) from sim, z (matchability logits [S*Lp]) and lse. */
int lgb200_x3_assign_scores(const float* sim, const float* z, const float* lse, int B, int Lp,
                            const int32_t* lens, int R, int C, float* scores, void* stream);

/* ---- training path: backward kernels -------------------------------------
 * The reference trains through this path with autograd (train.py -> LightGlue.forward, lightglue.py:484-498, and
 * LightGlue.loss, :588-637).  fp32 (the layouts of LGB200_F32).  Plain GEMMs of the backward pass (dX = dY.W,
 * dW = dY^T.X) are cuBLAS calls on the host side (glue_factory_colon_b200/train.py); the fused forward ops have
 * these fused backward ops:
 *
 * lgb200_attention_bwd: backward of lgb200_attention (flash-style: the score matrix is recomputed tile by tile,
 *   nothing N x M is stored).  Q, K, V [S,4,Lp,64] as given to the forward (Q pre-scaled, log2 domain), ctx / dctx
 *   [S,Lp,256] = forward output and its gradient.  dQ, dK, dV [S,4,Lp,64]: gradients w.r.t. the given (scaled) Q,
 *   K and V; with kv_xor = 1, dK / dV of sequence s collect the queries of sequence s ^ 1.  Rows >= lens are zero.
 *   Runs on tcgen05 in the fp32-accurate split-fp16 mode (csrc/lg_x3_attn_bwd.cu).  workspace:
 *   lgb200_attention_bwd_workspace(S, Lp) floats (row log-sum-exp, <dO, O>, and the fp16 plane pairs of Q, K, V, dO).
 * lgb200_attention_bwd_workspace: *n_floats = the workspace size of lgb200_attention_bwd for (S, Lp).
 * lgb200_heads_bwd: backward of the HEADS epilogue of lgb200_linear (unflatten + rotary + scale, lightglue.py:43-50,
 *   157-161, 196-201).  n_parts = 3: out [T,768] = [scale0 R^T dQ | scale1 R^T dK | scale2 dV] in the packed column
 *   order part*256 + head*64 + d, R^T = inverse rotation by the angles in rot [T,64]; dtheta [T,32] (nullable) +=
 *   the gradient of the rotary angles summed over heads and q / k (accumulates over layers; feeds posenc.Wr).
 *   n_parts = 2 (cross block): out [T,512] = [scale0 (dQ + dK) | scale1 dV]; Q, K, rot, dtheta unused.
 * lgb200_ln_gelu_bwd: backward of GELU(LayerNorm(h) gamma + beta) on rows of 512 (lightglue.py:144-149): h = the
 *   pre-LayerNorm activations, da = gradient of the GELU output; dh = gradient of h, act (nullable) = the GELU
 *   output recomputed (the A operand of d W2); partials [n_partials][1024] receives per-CTA column sums of
 *   (d gamma | d beta) -- the caller adds the n_partials rows.  Rows >= lens get zeros.
 * lgb200_assign_dsim: backward of sigmoid_log_double_softmax (lightglue.py:257-269) w.r.t. the similarity, for a
 *   loss that is linear in log_assignment (NLLLoss, losses.py:6-26): sim [B,m,n] is replaced in place by
 *   2 g_pos[b] gt - r[b,i] exp(sim - lse0[i]) - c[b,j] exp(sim - lse1[j]); lse as written by lgb200_assign_lse
 *   ([S,Lp]: rows of image 0, columns of image 1), r / c = g_pos times the row / column sums of gt. */
int lgb200_attention_bwd(const float* Q, const float* K, const float* V, const float* ctx, const float* dctx,
                         int S, int Lp, const int32_t* lens, int kv_xor, float* dQ, float* dK, float* dV,
                         float* workspace, void* stream);
int lgb200_attention_bwd_workspace(int S, int Lp, long long* n_floats);
int lgb200_heads_bwd(const float* dQ, const float* dK, const float* dV, const float* Q, const float* K,
                     const float* rot, int S, int Lp, const int32_t* lens, int n_parts, float scale0,
                     float scale1, float scale2, float* out, float* dtheta, void* stream);
int lgb200_ln_gelu_bwd(const float* h, const float* gamma, const float* beta, const float* da, int T, int Lp,
                       const int32_t* lens, float* dh, float* act, float* partials, int n_partials,
                       void* stream);
int lgb200_assign_dsim(float* sim, int B, int m, int n, const float* lse, int Lp,
                       const uint8_t* gt_assignment, const float* g_pos, const float* r, const float* c,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTGLUE_B200_H_ */
