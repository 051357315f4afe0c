/* svol_b200 -- C ABI of the B200-native SVOL hot path (libsvol_b200.so, sm_100a).
 *
 * Drop-in boundary for the reference's lib/modeling path (SVANet.forward ->
 * PerFrameMatcher / HungarianMatcher -> SetCriterion).  The reference has no native layer of
 * its own; every op below replaces a PyTorch / scipy call of the reference, cited per entry.
 * Host code (svol_b200/*.py here, or the reference's own train.py / test.py through
 * INTEGRATION.md's stubs) binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named h_*.
 *   - the caller (PyTorch) owns every buffer; nothing is allocated, freed or retained.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous.
 *   - return value: 0 on success; negative SVOL_ERR_* for argument errors; positive values
 *     are cudaError_t codes.  svol_last_error() returns a message for the calling thread.
 *   - bf16 buffers are raw uint16 storage (torch.bfloat16); row-major unless stated.
 *   - there is no CPU fallback: every entry point launches sm_100a kernels.
 */
#ifndef SVOL_B200_H_
#define SVOL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVOL_ABI_VERSION 13

enum {
  SVOL_OK = 0,
  SVOL_ERR_SHAPE = -1,     /* unsupported or inconsistent sizes / alignment */
  SVOL_ERR_NULL = -2,      /* required pointer missing */
  SVOL_ERR_DRIVER = -3,    /* driver entry point (cuTensorMapEncodeTiled) unavailable / failed */
  SVOL_ERR_DEVICE = -4     /* device is not compute capability 10.x */
};

enum { SVOL_ACT_NONE = 0, SVOL_ACT_RELU = 1, SVOL_ACT_GELU = 2 };

typedef uint16_t svol_bf16;

int svol_abi_version(void);
const char* svol_last_error(void);
/* 0 if the current device can run this library (sm_100), SVOL_ERR_DEVICE otherwise. */
int svol_device_check(void);
/* sizeof() of the argument structures as this library was compiled (0 gemm_args, 1 attn_args,
 * 2 match_args, 3 criterion_args, 4 gemm_epilogue, 5 ffn_args, 6 attn_bwd_args) -- lets a foreign-language binding verify its
 * struct layout before the first launch. */
int svol_sizeof_args(int which);

/* ------------------------------------------------------------------------------------------
 * Dense projections.  out = epilogue(A[M,K] x W[N,K]^T), bf16 operands, fp32 accumulation in
 * TMEM (tcgen05.mma).  Replaces nn.Linear (+ F.relu / F.gelu / residual add / nn.LayerNorm):
 *   input projections        lib/modeling/svanet.py:49-60,159-181
 *   attention in/out proj    lib/modeling/cross_modal_transformer.py:88-97,137-141,145-156
 *   FFN fc1/GELU/fc2 + norm  lib/modeling/cross_modal_transformer.py:142-143,157-158,163-179
 *   box-head hidden layers   lib/modeling/svanet.py:144-156
 * Epilogue order: +bias (-> out_pre) -> act -> * f'(dact_src) -> +residual -> LayerNorm(ln_weight, ln_bias, ln_eps) -> stores.
 * Requirements: N % 256 == 0, K % 64 == 0; LayerNorm and the transposed store need N == 256;
 * A, W, out*, residual 16-byte aligned with row pitches (in elements) multiple of 8.
 * ------------------------------------------------------------------------------------------ */
typedef struct svol_gemm_epilogue {
  const float* bias;          /* [N] or NULL */
  int32_t act;                /* SVOL_ACT_* */
  int32_t ld_res;
  const svol_bf16* residual;  /* [M, ld_res] or NULL */
  const float* ln_weight;     /* [N] or NULL (no LayerNorm) */
  const float* ln_bias;       /* [N] */
  float ln_eps;
  int32_t ld_out;
  svol_bf16* out;             /* [M, ld_out] or NULL */
  svol_bf16* out_pos;         /* [M, ld_out] second output = result + pos, or NULL */
  const float* pos;           /* [*, ld_pos] fp32 */
  int32_t ld_pos;
  int32_t pos_row_mod;        /* pos row = row % pos_row_mod (0: pos row = row) */
  svol_bf16* out_vt;          /* per-head transposed output [(M / vt_len) * N, vt_pitch] or NULL */
  int32_t vt_len;             /* tokens per sample (row = b * vt_len + l) */
  int32_t vt_pitch;           /* row pitch of out_vt in elements (multiple of 8, >= vt_len) */
  const float* pos_theta;     /* alternative to pos for out_pos: fp32 [M] angles from svol_posenc_theta; the sine
                                 encoding is evaluated in the epilogue (needs N == 256), no table is read */
  /* training step */
  svol_bf16* out_pre;         /* [M, ld_out] or NULL: the value BEFORE the activation (after bias), stored next to
                                 out = act(...): the FFN's fc1 keeps its pre-activation for the backward.  Excludes out_pos. */
  const svol_bf16* dact_src;  /* [M, ld_dact] or NULL: multiply the accumulator by f'(dact_src) before the residual add --
                                 the activation backward fused into the dgrad GEMM.  dact_mode SVOL_ACT_GELU: dact_src is
                                 the saved pre-activation; SVOL_ACT_RELU: the saved activation output (mask = src > 0). */
  int32_t ld_dact;
  int32_t dact_mode;
} svol_gemm_epilogue;

typedef struct svol_gemm_args {
  const svol_bf16* A;  /* [M, lda] */
  const svol_bf16* W;  /* [N, ldw] */
  int32_t M, N, K, lda, ldw;
  int32_t split_block; /* 0, or (K == 256 only) the first 256-column block that takes its A operand from A2 and writes
                          the per-head transposed output out_vt (columns relative to the split); the blocks before it
                          use A and write out / out_pos [M, split_block * 256].  One launch then computes e.g. the
                          q/k projection of x + pos and the v projection of x (cross_modal_transformer.py:137-139). */
  svol_gemm_epilogue ep;
  const svol_bf16* A2; /* [M, lda2] or NULL */
  int32_t lda2;
  int32_t ld_f32;
  float* out_f32;      /* or NULL.  Weight-gradient mode: out_f32[M, ld_f32] += A x W^T in fp32 (atomic accumulation, the
                          contraction is split over the SMs); excludes every epilogue option.  dW = dY^T X of an
                          nn.Linear with A = dY^T [N_out, rows], W = X^T [K_in, rows] (svol_transpose_bf16). */
  int32_t mn_major;    /* with out_f32 only: the operands are given UNtransposed, A = dY [K, lda] (K rows, M columns) and
                          W = X [K, ldw] (K rows, N columns), the contraction index is the row; tiles are consumed through
                          MN-major shared-memory descriptors, no transposed copies are needed.  M % 8 == 0, any K. */
  int32_t reserved;
} svol_gemm_args;

int svol_gemm_bf16(const svol_gemm_args* args, void* stream);
/* Same contract, plain SIMT kernel (one warp per row).  Test / triangulation aid only. */
int svol_gemm_bf16_plain(const svol_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused FFN block:  out = LayerNorm(x + fc2(GELU_erf(fc1(x))))   (and optionally out_pos = out + pos).
 * Replaces MLP.forward + residual + norm3 / norm6 (lib/modeling/cross_modal_transformer.py:142-143,
 * 157-158,163-179).  The [M, ff] hidden activation stays on the SM (tensor memory -> shared memory);
 * bf16 operands, fp32 accumulation.  Requirements: d == 256, ff % 256 == 0, ff <= 2048, 16-byte aligned
 * buffers with row pitches multiple of 8 elements.
 *   x [M, ldx] bf16; w1 [ff, ldw1] bf16 (fc1.weight); b1 [ff] fp32; w2 [d, ldw2] bf16 (fc2.weight); b2 [d] fp32
 *   ln_weight / ln_bias [d] fp32; out, out_pos [M, ld_out] bf16; pos fp32 rows (row, or row % pos_row_mod)
 *   reserved must be 0.
 * ------------------------------------------------------------------------------------------ */
typedef struct svol_ffn_args {
  const svol_bf16* x;
  const svol_bf16* w1;
  const float* b1;
  const svol_bf16* w2;
  const float* b2;
  const float* ln_weight;
  const float* ln_bias;
  svol_bf16* out;
  svol_bf16* out_pos;       /* or NULL */
  const float* pos;         /* fp32 table for out_pos (rows follow the output rows, or repeat with pos_row_mod) */
  const float* pos_theta;   /* alternative to pos: fp32 [M] angles from svol_posenc_theta; the sine encoding
                               (position_encoding.py:62-71) is then evaluated inside the kernel */
  int32_t M, d, ff, ldx, ldw1, ldw2, ld_out, ld_pos, pos_row_mod;
  float ln_eps;
  int32_t reserved;
} svol_ffn_args;

int svol_ffn_bf16(const svol_ffn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-head attention core, flash style (scores never leave the SM): tcgen05 QK^T and PV
 * with TMEM accumulators, TMA-fed K / V^T ring, online softmax in registers.
 * Replaces the softmax(QK^T/sqrt(dh) + mask) V part of nn.MultiheadAttention at
 * lib/modeling/cross_modal_transformer.py:139 (video self-attention), :147 (query
 * self-attention) and :154 (query -> video cross-attention with key_padding_mask).
 *   q   [B*Lq, ldq]  head h at columns [h*32, h*32+32); ALREADY scaled by log2(e)/sqrt(32)
 *   k   [B*Lk, ldk]  same head layout
 *   vt  [B*H*32, vt_pitch]  V transposed per head (row (b*H+h)*32+d, column = key index);
 *                     columns [Lk, vt_pitch) must be zero
 *   key_mask [B, Lk] float, nonzero = valid key, or NULL (all valid)
 *   out [B*Lq, ldo]  heads concatenated (the input of out_proj)
 * head_dim is fixed at 32 (hidden_dim 256 / 8 heads, lib/configs.py:117-120).
 * ------------------------------------------------------------------------------------------ */
typedef struct svol_attn_args {
  const svol_bf16* q;
  const svol_bf16* k;
  const svol_bf16* vt;
  const float* key_mask;
  svol_bf16* out;
  int32_t B, H, Lq, Lk, ldq, ldk, ldo, vt_pitch;
  float* lse;          /* optional (training forward): base-2 log-sum-exp of the scaled scores, [B, H, lse_pitch] */
  int32_t lse_pitch;   /* >= Lq */
  int32_t reserved;
} svol_attn_args;

int svol_attention_bf16(const svol_attn_args* args, void* stream);
int svol_attention_bf16_plain(const svol_attn_args* args, void* stream);   /* SIMT, tests only */

/* ------------------------------------------------------------------------------------------
 * Row-wise / memory-bound pieces of the head.
 * ------------------------------------------------------------------------------------------ */
/* y = LayerNorm(x) rows of `cols` fp32 -> bf16.  First op of LinearLayer (svanet.py:174-176). */
int svol_layernorm_f32_to_bf16(const float* x, const float* weight, const float* bias, svol_bf16* y,
                               int32_t rows, int32_t cols, float eps, void* stream);

/* The same first LayerNorm for frame features kept in bf16 by the caller (a precomputed feature cache; the reference
 * stores its sketch features precomputed, preprocess/sketch_vit_feature_extractor.py): x [rows, cols] bf16. */
int svol_layernorm_bf16_to_bf16(const svol_bf16* x, const float* weight, const float* bias, svol_bf16* y, int32_t rows,
                                int32_t cols, float eps, void* stream);

/* Backbone hand-off (backbone.py:72-89, model.py:18-22): x [frames, channels, spatial] fp32 is the ResNet
 * trunk's (N*T, C, h, w) feature map as cuDNN leaves it; y [frames*spatial, channels] bf16 = LayerNorm over the channels
 * of token (frame, position), i.e. the first LayerNorm of the head applied to the (N, T*h*w, C) token layout without
 * materialising the reshaped / transposed fp32 copy the reference builds.  channels <= 1024, channels*spatial*4 <= 200 KB. */
int svol_layernorm_nchw_to_bf16(const float* x, const float* weight, const float* bias, svol_bf16* y, int32_t frames,
                                int32_t channels, int32_t spatial, float eps, void* stream);

/* y = [ReLU](Linear(LayerNorm(x))) in fp32 for a handful of rows: the sketch branch of the input
 * projection (svanet.py:56-60,87), one call per LinearLayer.  x [rows,in], w [out,in], y [rows,out]. */
int svol_ln_linear_f32(const float* x, const float* ln_weight, const float* ln_bias, const float* w,
                       const float* b, int32_t relu, float* y, int32_t rows, int32_t in_dim,
                       int32_t out_dim, float eps, void* stream);

/* Sine positional encoding, PositionEmbeddingSine(normalize=True) (position_encoding.py:51-71).
 * mask [B,L] float (nonzero = valid) -> pos [B,L,d] fp32. */
int svol_posenc_sine(const float* mask, float* pos, int32_t B, int32_t L, int32_t d, void* stream);

/* theta[b,l] = cumsum(mask)[b,l] / (sum(mask[b]) + 1e-6) * 2 pi : the per-token angle of the normalised sine
 * encoding (position_encoding.py:55-61); column pairs of the table are (sin, cos)(theta / dim_t).  [B*L] fp32. */
int svol_posenc_theta(const float* mask, float* theta, int32_t B, int32_t L, void* stream);

/* out[r,:] = bf16(x[r % mod,:] (+ pos[r % mod,:])) : broadcast the query embedding over the batch
 * (cross_modal_transformer.py:52-56) and build x + pos operands. */
int svol_add_pos_bf16(const float* x, const float* pos, svol_bf16* out, int32_t rows, int32_t cols,
                      int32_t mod, void* stream);

/* Sketch-conditioned gate (cross_modal_transformer.py:122-127): the sketch->video attention only
 * contributes its head-averaged weights, so K is never materialised:
 *   u[b,h,:]  = (1/sqrt(dh)) * Wk_h^T (Wq_h s_b + bq_h)            svol_gate_vectors
 *   s[b,h,l]  = (x+pos)[b,l,:] . u[b,h,:]                            svol_gate_scores
 *   att[b,l]  = mean_h softmax_l(s[b,h,:])                           (inside svol_gate_apply)
 *   mem       = LayerNorm1(x * (1 + att));  also emits mem + pos     svol_gate_apply
 * in_proj_weight [3d,d], in_proj_bias [3d] are the fp32 parameters of sketch_video_cross_attn. */
int svol_gate_vectors(const float* sketch, const float* in_proj_weight, const float* in_proj_bias,
                      float* u, int32_t B, int32_t d, int32_t H, void* stream);
int svol_gate_scores(const svol_bf16* xpos, const float* u, float* scores, int32_t B, int32_t L,
                     int32_t d, int32_t H, void* stream);
int svol_gate_apply(const svol_bf16* x, const float* scores, const float* ln_weight, const float* ln_bias,
                    const float* pos, svol_bf16* mem, svol_bf16* mem_pos, float* att_out /* [B,L] or NULL */,
                    int32_t B, int32_t L, int32_t d, int32_t H, float eps, void* stream);
/* Same, with the sine positions evaluated in place from theta [B*L] (svol_posenc_theta) instead of a table. */
int svol_gate_apply_theta(const svol_bf16* x, const float* scores, const float* ln_weight, const float* ln_bias,
                          const float* theta, svol_bf16* mem, svol_bf16* mem_pos, float* att_out, int32_t B, int32_t L,
                          int32_t d, int32_t H, float eps, void* stream);

/* svol_gate_scores + the softmax + svol_gate_apply_theta in ONE launch (cross_modal_transformer.py:122-127): a cluster of
 * 8 CTAs per sample keeps the sample's token rows in shared memory (read from HBM once), forms the scores from x + pos
 * in fp32 (positions evaluated from theta) and closes the softmax over the L tokens through distributed shared memory.
 * Available when a sample's rows fit one cluster (svol_gate_fused_supported(L) == 1, L <= 3384); att_out [B,L] and
 * scores_out [B,H,L] are optional (NULL). */
int svol_gate_fused_supported(int32_t L);
int svol_gate_fused(const svol_bf16* x, const float* u, const float* ln_weight, const float* ln_bias, const float* theta,
                    svol_bf16* mem, svol_bf16* mem_pos, float* att_out, float* scores_out, int32_t B, int32_t L,
                    int32_t d, int32_t H, float eps, void* stream);

/* Output heads (svanet.py:125-127): logits = class_embed(hs); boxes = sigmoid(bbox_embed.layers.2(h2))
 * where h2 is the output of the two hidden box-MLP layers (svol_gemm_bf16 with ReLU).
 * hs, h2 [rows, d] bf16; wc [2,d], bc [2], wb [4,d], bb [4] fp32; logits [rows,2], boxes [rows,4] fp32. */
int svol_heads(const svol_bf16* hs, const svol_bf16* h2, const float* wc, const float* bc, const float* wb,
               const float* bb, float* logits, float* boxes, int32_t rows, int32_t d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hungarian matching (lib/modeling/matcher.py) for ALL decoder layers in one launch.
 *
 * A "problem" is one assignment: rows = `rows_per_problem` consecutive queries of one video,
 * columns = the targets [tgt_off[p], tgt_off[p+1]).  PerFrameMatcher (matcher.py:38-119):
 * problems_per_video = T, rows_per_problem = queries per frame.  HungarianMatcher (:131-159):
 * problems_per_video = 1, rows_per_problem = Q.
 *
 * Cost block (only the block-diagonal entries the reference ever reads, matcher.py:92-93) in the
 * reference's fp32 operation order:  C = w_bbox*L1 + w_giou*(-GIoU) + w_class*(-softmax(logits)[0])
 * (matcher.py:59-85, box_utils.py:9-13,24-61), then scipy.optimize.linear_sum_assignment's
 * algorithm (Crouse 2016 shortest augmenting paths on fp64-promoted costs, identical
 * tie-breaking; matcher.py:93,158) run by one warp per problem.
 *
 *   logits [NL,B,Q,2], boxes [NL,B,Q,4] fp32 (cxcywh); tgt_boxes [S,4] fp32 (cxcywh)
 *   tgt_off   [P+1] int32, P = B*problems_per_video : target range of problem p
 *   match_off [P+1] int32 : output range of problem p (min(rows, cols) entries)
 *   cost_ws   fp32 workspace or NULL.  When given, every problem's cost block (rows x cols, row-major) is also
 *             written to it: cost_off[p] (int64, [P+1]) = offset of problem p's block inside one layer's slab of
 *             cost_off[P] floats; total NL*cost_off[P] floats.  Required when rows_per_problem * max_cols * 4 bytes
 *             exceed 96 KB (the blocks are then read back from the workspace instead of shared memory, and are
 *             stored in the solver's working orientation: TRANSPOSED, cols x rows, for problems with rows > cols)
 *             and for mode 1 / 2.
 *   pred_idx, tgt_idx [NL, K] int64, K = match_off[P]: query index inside the video, target index (global, or
 *             local to the video after `localize`)
 *   status    [2] int32, ZERO-INITIALISED ONCE by the caller and then owned by the library: status[0] = result of
 *             the most recent svol_match on this buffer (bit 0: NaN / -inf cost entries -- scipy's "matrix contains
 *             invalid numeric entries"; bit 1: infeasible problem), status[1] = accumulator of the call in flight
 *             (published and cleared by the call's own finalize kernel: no host-side memset per call, the call can
 *             be captured in a CUDA graph).  A problem without a solution still gets valid indices (row r <->
 *             column min(r, cols-1)) so that svol_criterion never dereferences garbage; check status[0].
 *   mode      0: cost blocks + assignment (default); 1: cost blocks only (into cost_ws); 2: assignment from the
 *             blocks already in cost_ws.  1 and 2 exist to time / profile the two halves separately.
 *   solver    0: per-column solver state in registers (problems up to 320 working columns; wider ones fall back);
 *             1: force the shared-memory-state solver (any size; test / triangulation aid)
 *   localize  0: tgt_idx stays global; 1: PerFrameMatcher (matcher.py:114-115): per (layer, video) subtract the
 *             minimum matched global index; 2: HungarianMatcher (matcher.py:158): subtract the video's first target
 *             (video_tgt_off).  video_match_off [B+1] int32 = output range of each video.
 * ------------------------------------------------------------------------------------------ */
typedef struct svol_match_args {
  const float* logits;
  const float* boxes;
  const float* tgt_boxes;
  const int32_t* tgt_off;
  const int32_t* match_off;
  const int64_t* cost_off;
  float* cost_ws;
  int64_t* pred_idx;
  int64_t* tgt_idx;
  int32_t* status;
  int32_t NL, B, Q, problems_per_video, rows_per_problem, max_cols;
  float w_class, w_bbox, w_giou;
  int32_t K;                        /* match_off[P] as known to the host: row pitch of pred_idx / tgt_idx */
  const int32_t* video_match_off;   /* [B+1] or NULL (localize == 0) */
  const int32_t* video_tgt_off;     /* [B+1] or NULL (localize != 2) */
  int32_t mode, solver, localize, reserved;
  const int32_t* order;             /* [P] permutation of the problems or NULL: launch order, e.g. most columns first
                                       (longest-processing-time-first: the launch does not end on its largest problem) */
} svol_match_args;

int svol_match(const svol_match_args* args, void* stream);

/* PerFrameMatcher's localisation quirk (matcher.py:114-115) as a stand-alone call for callers that ran svol_match with
 * localize == 0: per (layer, video) subtract the minimum matched global target index. */
int svol_match_localize(int64_t* tgt_idx, const int32_t* video_match_off, int32_t NL, int32_t B,
                        int32_t K, void* stream);

/* Batched scipy.optimize.linear_sum_assignment (matcher.py:93,158) on caller-supplied fp32 cost matrices, one warp
 * per problem, the solver of svol_match.  Problem p: shape[2p] x shape[2p+1] row-major at cost + cost_off[p];
 * min(rows, cols) assignments written to rows_out / cols_out + out_off[p] (rows ascending, as scipy returns them);
 * status[p] = 0 ok, 1 invalid entries (NaN / -inf), 2 infeasible.  max_small / max_big = upper bounds of min / max(rows,
 * cols) over the batch, max_entries of rows * cols; max_entries * 4 bytes must fit in shared memory (<= ~190 KB).
 * solver as in svol_match_args. */
int svol_lsap_f32(const float* cost, const int64_t* cost_off, const int32_t* shape, int32_t n_problems,
                  int32_t max_small, int32_t max_big, int32_t max_entries, int64_t* rows_out, int64_t* cols_out, const int64_t* out_off,
                  int32_t* status, int32_t solver, void* stream);

/* ------------------------------------------------------------------------------------------
 * SetCriterion losses (lib/modeling/loss.py:39-60,76-103) for ALL decoder layers in one launch.
 *   losses [NL,4] fp32 = (loss_label, class_error, loss_bbox, loss_giou) per layer
 *   pred_idx / tgt_idx [NL,K] int64 as produced above (tgt_idx video-local),
 *   match_video [K] int32 = video of each matched pair (or NULL: derived from video_match_off [B+1] int32, the
 *   output range of each video), video_tgt_off [B+1] int32.
 *   meta: NULL, or a device int32 array whose first element is K -- the launch then does not depend on the batch's
 *   number of matched pairs (one captured CUDA graph serves every batch); idx_pitch = fixed row pitch (>= K) of
 *   pred_idx / tgt_idx in that case (0: pitch = K).  Out-of-range indices are clamped, never dereferenced.
 * svol_criterion_backward writes d(sum_i w_i * loss_i)/d(logits, boxes) with per-layer weights
 * grad_w [NL,3] = (w_label, w_bbox, w_giou) (train.py:227-228).
 * ------------------------------------------------------------------------------------------ */
typedef struct svol_criterion_args {
  const float* logits;
  const float* boxes;
  const float* tgt_boxes;
  const int64_t* pred_idx;
  const int64_t* tgt_idx;
  const int32_t* match_video;
  const int32_t* video_tgt_off;
  float* losses;
  int32_t NL, B, Q, K;
  float eos_coef;
  int32_t idx_pitch;
  const int32_t* video_match_off;
  const int32_t* meta;
  void* scratch;     /* NULL: one CTA per layer.  Else >= svol_criterion_scratch_bytes(NL, B) bytes, ZERO-INITIALISED ONCE by
                        the caller (needs video_match_off): one CTA per (video, layer), fp64 partial sums combined in video
                        order by the last CTA to arrive, which also resets the arrival counters */
} svol_criterion_args;
/* NL * B * 4 doubles + NL int32 counters, rounded up to 16 bytes */
int64_t svol_criterion_scratch_bytes(int32_t NL, int32_t B);

int svol_criterion(const svol_criterion_args* args, void* stream);
int svol_criterion_backward(const svol_criterion_args* args, const float* grad_w, float* grad_logits,
                            float* grad_boxes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Inference post-processing (test.py:133-158): foreground score = softmax(logits)[0],
 * clamp(cxcywh->xyxy, 0, 1), per-frame groups of q_per_frame queries sorted by score
 * (descending, stable).  out [B, Q, 5] fp32 = (x0,y0,x1,y1,score) in sorted order,
 * order [B, Q] int32 = source query (within its frame) of each output row.
 * ------------------------------------------------------------------------------------------ */
int svol_postprocess(const float* logits, const float* boxes, float* out, int32_t* order, int32_t B,
                     int32_t Q, int32_t q_per_frame, void* stream);


/* ------------------------------------------------------------------------------------------
 * Evaluation metrics (lib/evaluate/eval.py:20-117, lib/evaluate/utils.py:36-202) on the device arrays the
 * forward produced.  pred = svol_postprocess output viewed as [B*T, q_per_frame, 5] (per-frame score-sorted x0,y0,x1,y1,
 * score); evaluated frame f reads pred[frame_index[f]]; gt [n_gt, 4] fp32 xyxy, frame f owns gt[gt_off[f] : gt_off[f+1]];
 * evaluation unit (video + sketch) u owns frames [frame_off[u], frame_off[u+1]).  Predictions are rounded to 4 decimals
 * (test.py:161) and all comparisons are float64 in the reference's operation order.
 *   svol_eval_max_iou            max1 / max5 [n_gt] fp64: column maxima of the reference's (k, n_f) IoU array for k = 1, 5
 *                                (eval.py:75-90, including compute_iou_batch_cross's tile / repeat / reshape pairing)
 *   svol_eval_average_precision  ap [units, 10] fp64: interpolated AP at IoU 0.50:0.05:0.95 after the greedy, score-ordered
 *                                matching of utils.py:118-202
 * ------------------------------------------------------------------------------------------ */
int svol_eval_max_iou(const float* pred, const int32_t* frame_index, const float* gt, const int32_t* gt_off, int32_t frames,
                      int32_t n_gt, int32_t q_per_frame, double* max1, double* max5, void* stream);
int svol_eval_average_precision(const float* pred, const int32_t* frame_index, const float* gt, const int32_t* gt_off,
                                const int32_t* frame_off, int32_t units, int32_t q_per_frame, int32_t max_frames, int32_t max_gt,
                                double* ap, void* stream);

/* ==========================================================================================
 * TRAINING STEP (train.py:216-232: forward in train mode, criterion, loss.backward(), AdamW).
 * The reference gets its backward from torch.autograd over lib/modeling/svanet.py and
 * lib/modeling/cross_modal_transformer.py; here every backward op is an explicit entry point.
 * Dense gradients reuse svol_gemm_bf16:  dX = dY W  (W operand = the transposed weight) and
 * dW = dY^T X  (both operands transposed by svol_transpose_bf16, contraction over the token rows).
 * Activation gradients are bf16, parameter gradients fp32 (accumulated: zero them first).
 * ========================================================================================== */

/* y = LayerNorm(z) (and y_pos = y + pos), bf16 [rows, 256]; pos = fp32 table (row, or row % pos_mod) or theta
 * (fp32 [rows] angles of svol_posenc_theta).  Training forward of norm1..norm6 / the input projection's second
 * LayerNorm (cross_modal_transformer.py:127,141,143,149,156,158; svanet.py:174-176) when z must be kept. */
int svol_layernorm_bf16(const svol_bf16* z, const float* weight, const float* bias, svol_bf16* y, svol_bf16* y_pos,
                        const float* pos, int32_t pos_mod, const float* theta, int32_t rows, int32_t cols, float eps,
                        float drop_p, const int64_t* seed, int32_t site, void* stream);

/* Train-mode Dropout of the input projections (svanet.py:168-170: LayerNorm -> Dropout -> Linear).  The mask is
 * counter based and stateless: element idx (row * cols + col) of dropout site `site` is kept iff the top 32 bits of
 * splitmix64(idx + (8 * *seed + site) * 0x9E3779B97F4A7C15) are >= drop_p * 2^32; kept values are scaled by
 * 1 / (1 - drop_p).  `seed` is a DEVICE scalar (the host bumps it every step; the plans replay as CUDA graphs).  The
 * backward entry points recompute the mask from the same (seed, site) instead of storing it.  drop_p == 0 disables it.
 * Dropout variants of the two inference-path LayerNorm entry points: */
int svol_layernorm_f32_to_bf16_dropout(const float* x, const float* weight, const float* bias, svol_bf16* y, int32_t rows,
                                       int32_t cols, float eps, float drop_p, const int64_t* seed, int32_t site, void* stream);
int svol_ln_linear_f32_dropout(const float* x, const float* ln_weight, const float* ln_bias, const float* w, const float* b,
                               int32_t relu, float* y, int32_t rows, int32_t in_dim, int32_t out_dim, float eps, float drop_p,
                               const int64_t* seed, int32_t site, void* stream);

/* LayerNorm backward, cols = 256, 512, 768 or 1024.  z = forward input (bf16, or fp32 when z_is_f32), optionally times
 * (1 + att[row]) (the sketch gate, cross_modal_transformer.py:124-126: then dx = dz * (1 + att) and
 * datt[row] = sum_c dz * z_in).  dy = dy1 + dy2 + dy3 (dy2, dy3 optional).  dgamma / dbeta [cols] are accumulated. */
int svol_layernorm_backward(const void* z, int32_t z_is_f32, const float* att, const svol_bf16* dy1,
                            const svol_bf16* dy2, const svol_bf16* dy3, const float* gamma, svol_bf16* dx, float* datt,
                            float* dgamma, float* dbeta, int32_t rows, int32_t cols, float eps, float drop_p, const int64_t* seed,
                            int32_t site, void* stream);

/* y = GELU_erf(x) elementwise (F.gelu, cross_modal_transformer.py:163-179), n % 8 == 0. */
int svol_gelu_bf16(const svol_bf16* x, svol_bf16* y, int64_t n, void* stream);
/* out = dy * f'(saved).  mode SVOL_ACT_RELU: saved = activation output; SVOL_ACT_GELU: saved = pre-activation. */
int svol_act_backward(const svol_bf16* dy, const svol_bf16* saved, svol_bf16* out, int64_t n, int32_t mode, void* stream);

/* out[c, r] = in[r, c] (bf16; out pitch ld_out >= rows, columns >= rows untouched) and, if colsum != NULL,
 * colsum[c] += sum_r in[r, c] (the bias gradient of the Linear whose output gradient `in` is). */
int svol_transpose_bf16(const svol_bf16* in, int32_t ld_in, int32_t rows, int32_t cols, svol_bf16* out, int32_t ld_out,
                        float* colsum, void* stream);
int svol_colsum_bf16(const svol_bf16* in, int32_t ld_in, int32_t rows, int32_t cols, float* colsum, void* stream);

/* Attention backward (autograd of nn.MultiheadAttention's core at cross_modal_transformer.py:139,147,154).
 *   q [B*Lq, ldq] (pre-scaled as in svol_attention_bf16), k [B*Lk, ldk], v [B*Lk, ldv]: head h at columns [32h, 32h+32)
 *   kt [B*8*32, kt_pitch], qt, d_ot [B*8*32, qt_pitch]: per-head transposed K, Q and dO (svol_gemm_bf16 out_vt);
 *        columns beyond the sequence must be zero
 *   o, d_o [B*Lq, 256]: forward output and its gradient;  lse [B, 8, stat_pitch]: from svol_attention_bf16, entries
 *        [Lq, stat_pitch) must be +inf;  delta [B, 8, stat_pitch]: workspace, entries [Lq, stat_pitch) must be 0;
 *        stat_pitch % 64 == 0
 *   dq = gradient w.r.t. the UNSCALED query projection (x Wq^T + bq), dk, dv: [B*L, ld_*] bf16. */
typedef struct svol_attn_bwd_args {
  const svol_bf16* q;
  const svol_bf16* k;
  const svol_bf16* v;
  const svol_bf16* kt;
  const svol_bf16* qt;
  const svol_bf16* o;
  const svol_bf16* d_o;
  const svol_bf16* d_ot;
  const float* lse;
  float* delta;
  const float* key_mask;
  svol_bf16* dq;
  svol_bf16* dk;
  svol_bf16* dv;
  int32_t B, H, Lq, Lk, ldq, ldk, ldv, ld_o, ld_do, ld_dq, ld_dk, ld_dv, kt_pitch, qt_pitch, stat_pitch, reserved;
} svol_attn_bwd_args;
int svol_attention_backward_bf16(const svol_attn_bwd_args* args, void* stream);

/* Backward of svol_heads: dhs_cls = dlogits Wc; dh2 = relu'(h2) * ((dboxes * boxes * (1 - boxes)) Wb);
 * dwc [2,d], dbc [2], dwb [4,d], dbb [4] accumulated. */
int svol_heads_backward(const svol_bf16* hs, const svol_bf16* h2, const float* wc, const float* wb, const float* boxes,
                        const float* dlogits, const float* dboxes, svol_bf16* dhs_cls, svol_bf16* dh2, float* dwc,
                        float* dbc, float* dwb, float* dbb, int32_t rows, int32_t d, void* stream);

/* Backward of the sketch gate from datt [B*L] (produced by svol_layernorm_backward with att):
 *   dscores [B,H,L] workspace;  dx_out = dx_in + sum_h dscores u  (gradient w.r.t. the layer input x, through
 *   both x * (1 + att) and the scores of x + pos);  du [B,H,d] is overwritten. */
int svol_gate_backward(const svol_bf16* xpos, const float* u, const float* scores, const float* datt,
                       const svol_bf16* dx_in, svol_bf16* dx_out, float* dscores, float* du, int32_t B, int32_t L,
                       int32_t d, int32_t H, void* stream);
/* Backward of svol_gate_vectors: d_in_proj_weight [3d,d], d_in_proj_bias [3d], dsketch [B,d] accumulated. */
int svol_gate_vectors_backward(const float* sketch, const float* in_proj_weight, const float* in_proj_bias,
                               const float* du, float* d_in_proj_weight, float* d_in_proj_bias, float* dsketch,
                               int32_t B, int32_t d, int32_t H, void* stream);
/* Backward of svol_ln_linear_f32 (y = forward output).  dx [rows,in] is written (required: it doubles as the
 * kernel pair's workspace); parameter gradients are accumulated. */
int svol_ln_linear_f32_backward(const float* x, const float* ln_weight, const float* ln_bias, const float* w,
                                const float* y, const float* dy, int32_t relu, float* dx, float* d_ln_weight,
                                float* d_ln_bias, float* dw, float* db, int32_t rows, int32_t in_dim, int32_t out_dim,
                                float eps, float drop_p, const int64_t* seed, int32_t site, void* stream);
/* acc[r % mod, :] += g[r, :] summed over rows (bf16 -> fp32, 256 columns): query-embedding gradient. */
int svol_batch_sum(const svol_bf16* g, float* acc, int32_t rows, int32_t cols, int32_t mod, void* stream);
/* dst[i] (+)= scale * src[i]: bf16 weight-gradient GEMM output -> fp32 parameter gradient. */
int svol_accum_bf16(const svol_bf16* src, float* dst, int64_t n, float scale, int32_t accumulate, void* stream);
/* Refreshes the packed operand copies of the launch plans from the fp32 parameters in one launch.  Job k copies
 * src [rows, cols] fp32 (contiguous) to dst, multiplying rows [0, scaled_rows) by scale (the attention query rows carry
 * log2(e)/sqrt(dh)), as bf16 (SVOL_PACK_BF16) or fp32, optionally transposed (dst [cols, rows]).  jobs is a DEVICE array. */
enum { SVOL_PACK_BF16 = 1, SVOL_PACK_TRANSPOSE = 2 };
typedef struct svol_pack_job {
  const float* src;
  void* dst;
  int32_t rows, cols, scaled_rows, flags;
  float scale;
  int32_t reserved;
} svol_pack_job;
int svol_pack_weights(const svol_pack_job* jobs, int32_t n_jobs, void* stream);

/* Fused AdamW (torch.optim.AdamW, train.py:71-78) over one flat fp32 buffer; g is multiplied by grad_scale first. */
int svol_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int32_t step, float grad_scale, void* stream);

/* The same update over a flat buffer made of parameter SEGMENTS with per-group hyper-parameters (torch.optim
 * param_groups: train.py:72-96 builds them, the lr schedulers of train.py:129-137 rewrite group['lr']).
 *   seg_end   [n_seg] int64 (device): exclusive end offset of every segment, ascending, multiples of 4; the last = n
 *   seg_group [n_seg] int32 (device): group index of the segment, or -1 = skip it entirely (torch.optim.AdamW does
 *             not touch a parameter whose .grad is None: no weight decay, no moment update)
 *   groups    HOST array of n_groups <= SVOL_ADAMW_MAX_GROUPS entries, read during the call */
#define SVOL_ADAMW_MAX_GROUPS 8
typedef struct svol_adamw_group {
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;        /* 1-based step count of the group's active parameters (bias correction) */
} svol_adamw_group;
int svol_adamw_segments(float* p, const float* g, float* m, float* v, int64_t n, const int64_t* seg_end,
                        const int32_t* seg_group, int32_t n_seg, const svol_adamw_group* groups, int32_t n_groups,
                        float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVOL_B200_H_ */
