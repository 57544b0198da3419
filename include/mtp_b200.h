/* mtp_b200 — C ABI of the B200-native ViT/RVSA backbone kernels (libmtp_b200.so, sm_100a).
 *
 * The reference (ViTAE-Transformer/MTP) has no FFI of its own: the boundary it exposes for this path is the Python
 * class ViT_Win_RVSA_V3_WSZ7 (Multi-Task_Pretrain/backbone/vit_win_rvsa_v3_wsz7.py:587-817, "[V]" below).  Each entry
 * point here replaces the ATen/cuBLAS/cuDNN call sequence behind one piece of that class and cites it.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory unless the name ends in _host
 *  - the caller owns all memory (tensors and workspaces); the library never allocates or frees device memory and never
 *    synchronises; all work is enqueued on the cudaStream_t passed last (CUDA-graph capturable)
 *  - return 0 on success, a negative MTP_ERR_* code otherwise; mtp_last_error() gives the message (thread-local)
 *  - "tok" tensors are token-major row-major matrices [T, C] with T = B*Hp*Wp (image-major, then row, then column)
 *  - bf16 = __nv_bfloat16 bits (uint16), f32 = float
 */
#ifndef MTP_B200_H_
#define MTP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mtp_stream_t; /* == cudaStream_t */

enum {
  MTP_OK = 0,
  MTP_ERR_INVALID = -1,  /* bad argument / unsupported shape */
  MTP_ERR_CUDA = -2,     /* CUDA runtime / driver error at launch */
  MTP_ERR_NO_DEVICE = -3 /* no sm_100 device */
};

const char* mtp_last_error(void);
int mtp_version(void);
/* SM count of the current device (cached). */
int mtp_num_sms(void);
/* Programmatic dependent launch for every kernel of the library (default on): each kernel may be scheduled while its stream
 * predecessor drains and holds at griddepcontrol.wait until the predecessor's results are visible.  Off = plain stream order. */
int mtp_set_pdl(int enabled);
/* SM budget of the persistent / wave-sized kernels (GEMM grids and static tile schedules, LayerNorm-backward waves): launches
 * made while a limit n > 0 is set use at most n SMs.  The data-parallel trainer lowers it during the backward so that the NCCL
 * all-reduce running beside it (NCCL_MAX_CTAS channels) does not push a 148-CTA persistent GEMM into a second wave.  0 = all. */
int mtp_set_sm_limit(int n);

/* ---------------------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores:  acc[m,n] = sum_k A[m,k] * B[n,k]   (bf16 in, fp32 accumulate in TMEM)
 * replaces nn.Linear / Conv2d(k16,s16) / ConvTranspose2d(k2,s2) and their autograd dgrad/wgrad:
 *   [V]:78,87,96,109 (Attention qkv/proj)  [V]:256,262,390,430 (RVSA qkv/proj)  [V]:50-60 (Mlp fc1/fc2)
 *   [V]:529,536 (PatchEmbed.proj)          [V]:642,645,649 (fpn ConvTranspose2d)
 * Operand storage:
 *   a_mn_major = 0: A stored [M, lda] (k contiguous)     a_mn_major = 1: A stored [K, lda] (m contiguous)
 *   b_mn_major = 0: B stored [N, ldb] (k contiguous)     b_mn_major = 1: B stored [K, ldb] (n contiguous)
 * so forward is (0,0), dgrad dX = dY * W is (0,1), wgrad dW = dY^T * X is (1,1).
 * lda/ldb in elements, multiples of 8; base pointers 16-byte aligned; N % 8 == 0.
 * ------------------------------------------------------------------------------------------------------------- */
enum mtp_epilogue_mode {
  MTP_EPI_BF16 = 0,         /* out_bf16 = acc + bias                                                            */
  MTP_EPI_BF16_GELU = 1,    /* out2_bf16 = acc + bias (if out2); out_bf16 = gelu_erf(acc + bias)     [V]:56-57  */
  MTP_EPI_F32_RESID = 2,    /* out_f32 = aux_f32[m,n] + row_scale[m / rows_per_group] * (acc + bias) [V]:508-509 */
  MTP_EPI_F32_POS = 3,      /* out_f32 = acc + bias + aux_f32[m % pos_rows, n]                       [V]:793-794 */
  MTP_EPI_F32 = 4,          /* out_f32 (+)= acc + bias            (wgrad; accumulate flag)                       */
  MTP_EPI_BF16_DGELU = 5,   /* out_bf16 = acc * gelu'(aux_bf16[m,n])  (fc2 dgrad fused with GELU backward)       */
  MTP_EPI_BF16_PIXSHUF = 6  /* ConvTranspose2d(k2,s2) scatter: n = (dy*2+dx)*ps_cout + co, m = (b,y,x) ->
                               out_bf16[((b*2*ps_h + 2y+dy)*2*ps_w + 2x+dx) * ldo + co] = acc + bias[co]  [V]:642 */
};

typedef struct mtp_epilogue {
  int mode;
  int ldo;                /* leading dimension (elements) of out / out2 / aux (when aux is [M, ldo]-shaped) */
  const float* bias;      /* [N], or [ps_cout] applied with period ps_cout when ps_cout > 0 (any mode), or NULL */
  void* out;
  void* out2;             /* GELU mode: optional pre-activation copy */
  const void* aux;
  const float* row_scale; /* RESID: per-group multiplier (DropPath keep/keep_prob), NULL = 1 */
  int rows_per_group;     /* RESID: tokens per image */
  int pos_rows;           /* POS: rows of the positional table */
  int accumulate;         /* F32: add into out */
  int ps_h, ps_w, ps_cout; /* PIXSHUF geometry; ps_cout > 0 alone (other modes): bias[n % ps_cout] -- the ConvTranspose2d GEMM whose
                             output row holds the 4 sub-pixels side by side shares one [Cout] bias across them */
  float* colsum;          /* BF16 / BF16_DGELU: optional [N] fp32, += column sums of the stored values (the bias gradient of the
                             Linear whose cotangent this GEMM produces); 16-byte aligned */
  float* sumsq;           /* F32: optional scalar, += sum of squares of the stored outputs (gradient-norm clipping without a
                             separate pass over the weight gradients) */
  int hilo;               /* fp32-class mode ("fp32x3", forward only): A [M, 2K] and B [N, 2K] hold every value as TWO bf16 words, hi = bf16(v)
                             in columns [0, K) and lo = bf16(v - hi) in [K, 2K) (K-major, lda/ldb >= 2K); the kernel accumulates
                             A_hi B_hi + A_hi B_lo + A_lo B_hi in fp32 (three passes over K through the same tcgen05 pipeline: relative error
                             ~2^-16 instead of 2^-9).  bf16 outputs are written as hi | lo pairs too, see out_lo_offset */
  int out_lo_offset;      /* > 0: BF16 / BF16_GELU epilogues also store lo = bf16(v - bf16(v)) at out[m * ldo + out_lo_offset + n] */
  int b_static;           /* 1: operand B is NOT written by the kernels just before this one in the stream (weights, activations
                             saved earlier): its first tiles may be fetched before the programmatic-dependent-launch wait */
} mtp_epilogue;

int mtp_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N, int K,
                  const mtp_epilogue* ep, int force_bn /* 0 = heuristic */, mtp_stream_t stream);

/* Two INDEPENDENT GEMMs in one persistent launch (e.g. the dgrad dX = dY W and the wgrad dW = dY^T X of one nn.Linear, which
 * only share an input): the tiles of both are spread over the SMs by one longest-processing-time schedule, so SMs left idle
 * by one problem's tile count work on the other, and a single launch / prologue / epilogue tail is paid. */
typedef struct mtp_gemm_desc {
  const void* A; int lda; int a_mn_major;
  const void* B; int ldb; int b_mn_major;
  int M, N, K;
  const mtp_epilogue* ep;
} mtp_gemm_desc;
int mtp_gemm_bf16_dual(const mtp_gemm_desc* g0, const mtp_gemm_desc* g1, int force_bn, mtp_stream_t stream);
/* Host-only planning (no CUDA call, works without a GPU): the tile configuration (width + 1000 for pairs), CTA count and modelled
 * makespan in SM cycles the scheduler picks for one problem (M1 = 0) or a grouped pair, after checking that its work schedule
 * assigns every tile exactly once. */
int mtp_gemm_plan(int M0, int N0, int K0, int b0_mn_major, int M1, int N1, int K1, int b1_mn_major, int force_bn, int* out_config,
                  int* out_ctas, double* out_cycles);
/* Tile configuration the heuristic (or force_bn) selected for the most recent GEMM launch: width + 1000 for cta_group::2 pairs. */
int mtp_gemm_last_config(void);
/* Tuning aid: device buffer [grid][8] of int64 that receives per-CTA globaltimer stamps of the pipeline phases (NULL = off). */
int mtp_gemm_set_debug(void* device_buffer);
/* Tuning aid: 0 normal; 1 = skip the TMA loads (isolates the MMA pipeline; results are garbage); 2 = skip the MMAs;
 * 3 = skip the epilogue stores; 4 = empty launches (bench.py measures the GEMM share of a graph-replayed step with it). */
int mtp_gemm_set_debug_mode(int mode);

/* ---------------------------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound; one warp per token row; fp32 statistics).
 * mtp_layernorm_fwd replaces nn.LayerNorm(eps=1e-6) norm1/norm2 ([V]:484,496,596) and, with x_is_bf16 + fuse_gelu,
 * Norm2d -> nn.GELU of fpn1 ([V]:576-584,643-644).  y is bf16 (the next GEMM's A operand); mean/rstd may be NULL.
 * mtp_layernorm_bwd: dx = dres_f32 (optional residual-path gradient, [V]:508-509) + LN backward of dy; dgamma/dbeta
 * are ACCUMULATED (atomicAdd) so the caller zeroes them once per step.  Supported variants: (f32 x, f32 dx, no gelu)
 * and (bf16 x, bf16 dx, gelu).  C % 128 == 0, C <= 1024.
 * ------------------------------------------------------------------------------------------------------------- */
int mtp_layernorm_fwd(const void* x, int x_is_bf16, const float* gamma, const float* beta, void* y_bf16, float* mean,
                      float* rstd, int rows, int C, float eps, int fuse_gelu, mtp_stream_t stream);
int mtp_layernorm_bwd(const void* dy_bf16, const void* x, int x_is_bf16, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, const float* dres_f32, void* dx, int dx_is_bf16,
                      float* dgamma, float* dbeta,
                      /* optional fused mtp_scale_cast_bf16 of dx (f32 variant only; all NULL/0 = off): cast_out_bf16[r,c] =
                       * bf16(dx[r,c] * cast_row_scale[r / cast_rows_per_group]), cast_colsum += its column sums */
                      const float* cast_row_scale, int cast_rows_per_group, void* cast_out_bf16, float* cast_colsum,
                      /* optional (f32 variant): AvgPool backward of the RVSA sampling heads folded in -- dy[t, :] is used as
                       * dy[t, :] + pool_add[window(t), :] / 49 for the [pool_h, pool_w] token grid (rows = B * pool_h * pool_w);
                       * pool_add = the dpooled block mtp_rvsa_sampling_bwd leaves in its workspace */
                      const float* pool_add, int pool_h, int pool_w,
                      int rows, int C, int fused_gelu, mtp_stream_t stream);
/* out_bf16[r,c] = in[r,c] * row_scale[r / rows_per_group] (DropPath backward, [V]:31-39); colsum (optional) += column sums
 * of the scaled values (bias gradient of the Linear that produced the branch). */
int mtp_scale_cast_bf16(const float* in, const float* row_scale, int rows_per_group, void* out_bf16, float* colsum,
                        int rows, int C, mtp_stream_t stream);
/* colsum[c] += sum_r in_bf16[r*ld + c]   (bias gradients) */
int mtp_colsum_bf16(const void* in_bf16, int ld, float* colsum, int rows, int C, mtp_stream_t stream);
int mtp_cast_f32_bf16(const float* in, void* out_bf16, size_t n, mtp_stream_t stream);
int mtp_add_bf16_into_f32(const void* in_bf16, float* out, size_t n, mtp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Rotated varied-size window attention (RVSA), forward.  RotatedVariedSizeWindowAttention.forward, [V]:287-433.
 * mtp_rvsa_sampling_fwd: AvgPool2d(7) over the zero-padded LN'd tokens -> LeakyReLU -> the three 1x1 convs
 *   ([V]:228-243,347,354-368).  yn_bf16 [T, C]; w_off/w_scale [2nH, C], w_angle [nH, C] (Conv2d weights flattened);
 *   pooled (out) [B*nWin, C] pre-activation means (intermediate of the two kernels; saved for backward);
 *   params (out) [B*nWin, nH, 8] = (ox, oy, sx, sy, theta, -, -, -) with the [V]:359-360 divisions applied.
 * mtp_rvsa_attn_fwd: coords ([V]:372-385) -> bilinear K/V gather ([V]:397-404) -> scores + decomposed rel-pos with the
 *   UNscaled q ([V]:410-412) + bias table ([V]:414-418) -> softmax -> PV -> un-window + crop ([V]:420-428).
 *   qkv_bf16 [T, 3C] (q | k | v, head-major, hd = 64); rel_pos_h/w [13, 64]; bias_table [169, nH];
 *   out_bf16 [T, C]; lse (optional out) [B*nWin*nH, 49] log-sum-exp per query row (saved for backward).
 * ------------------------------------------------------------------------------------------------------------- */
int mtp_rvsa_sampling_fwd(const void* yn_bf16, const float* w_off, const float* b_off, const float* w_scale,
                          const float* b_scale, const float* w_angle, const float* b_angle, float* pooled, float* params,
                          int B, int h, int w, int C, int nH, mtp_stream_t stream);
int mtp_rvsa_attn_fwd(const void* qkv_bf16, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                      const float* bias_table, void* out_bf16, float* lse, int B, int h, int w, int C, int nH,
                      mtp_stream_t stream);

/* Dense attention with decomposed rel-pos bias (q scaled first), flash-style.  Attention.forward [V]:90-111 +
 * calc_rel_pos_spatial [V]:142-193.  rel_pos_h [2gh-1, 64], rel_pos_w [2gw-1, 64], both NULL = no bias (the mmdet /
 * mmrotate finetune twins).  lse (optional out) [B, nH, N]. */
int mtp_full_attn_fwd(const void* qkv_bf16, const float* rel_pos_h, const float* rel_pos_w, void* out_bf16, float* lse, int B,
                      int gh, int gw, int C, int nH, mtp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Layout kernels.  mtp_patchify: image (B, cin, H, W) f32|bf16 -> patch rows [T, cin*256] bf16 (A operand of the
 * patch-embed GEMM, [V]:529,536-539).  mtp_tok_to_nchw / mtp_nchw_to_tok: token-major matrix <-> NCHW map
 * ([V]:807), `level` = number of nested k2/s2 transposed convolutions whose 4 sub-pixels sit side by side in the row
 * (0: ld >= C; 1, 2: ld >= 4C).  mtp_maxpool2_tok_*: fpn4 = MaxPool2d(2,2) on the token grid ([V]:654).
 * ------------------------------------------------------------------------------------------------------------- */
int mtp_patchify(const void* img, int img_is_bf16, void* out_bf16, int B, int cin, int H, int W, mtp_stream_t stream);

/* f2: MTP_DataPreprocessor folded into the patch gather -- replaces Multi-Task_Pretrain/preprocessing.py:145-187
 * (mmengine ImgDataPreprocessor.forward: `_batch_input[[2,1,0]]` when bgr_to_rgb, `.float()`, `(x - mean) / std`; padding is the identity
 * at H,W == img_size) + PatchEmbed's im2col ([V]:536-539).  img_u8: uint8 [B, cin, H, W] (hwc = 0) or [B, H, W, cin] (hwc = 1);
 * mean / stdv: HOST arrays of cin floats in OUTPUT channel order (models.py:38-39); flip_channels = bgr_to_rgb | rgb_to_bgr. */
int mtp_patchify_u8(const void* img_u8, int hwc, int flip_channels, const float* mean, const float* stdv, void* out_bf16, int B, int cin,
                    int H, int W, mtp_stream_t stream);
int mtp_tok_to_nchw(const void* tok, int tok_is_bf16, int ld, void* out, int out_is_bf16, int B, int h, int w, int C, int level,
                    mtp_stream_t stream);
int mtp_nchw_to_tok(const void* in, int in_is_bf16, void* tok, int tok_is_bf16, int ld, int accumulate, int B, int h, int w, int C,
                    int level, mtp_stream_t stream);
int mtp_maxpool2_tok_fwd(const float* x, float* y, int B, int h, int w, int C, mtp_stream_t stream);
int mtp_maxpool2_tok_bwd(const float* x, const float* dy, float* dx, int B, int h, int w, int C, mtp_stream_t stream);
/* Stand-in objective when no decoder heads are attached (the reference's heads are third-party code): for one bf16
 * feature map of n elements (n % 8 == 0)  *loss += 0.5 * mean(f^2)  and  grad = f / n. */
int mtp_sqloss_fwd_bwd(const void* feat_bf16, void* grad_bf16, float* loss, size_t n, mtp_stream_t stream);
/* weighted: loss += weight * 0.5 * mean(f^2), grad = weight * f / n  (one stand-in head of the three-task step, models.py:327-335) */
int mtp_sqloss_fwd_bwd_w(const void* feat_bf16, void* grad_bf16, float* loss, size_t n, float weight, mtp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Attention backward (autograd of the forward entry points above; everything is recomputed from qkv + lse).
 * All parameter-gradient outputs (d_rel_pos_*, d_bias_table, dw_*, db_*) are ACCUMULATED into (+=).
 * mtp_rvsa_attn_bwd: dqkv_bf16 [T, 3C] is fully written (q slot directly; k|v slots from an fp32 scatter scratch);
 *   dparams [B*nWin, nH, 8] receives d(ox, oy, sx, sy, theta) (slots 5..7 zero); d_qkv_bias (optional, [3C]) += column sums
 *   of dqkv (the qkv bias gradient, [V]:390); workspace >= mtp_rvsa_bwd_workspace_bytes().  scratch_zeroed = 1: the caller keeps
 *   the workspace between calls, it is all-zero on entry (zero it once) and is handed back all-zero (no memset per call).
 * mtp_rvsa_sampling_bwd: backward of the pooled 1x1-conv heads ([V]:228-243): accumulates the six conv gradients and
 *   adds the AvgPool-path gradient into dyn_bf16 [T, C] (the cotangent of the LN1 output) -- or, with dyn_bf16 = NULL, leaves
 *   dpooled [B*nWin, C] fp32 at workspace + B*nWin*5*nH floats for mtp_layernorm_bwd(pool_add = ...) to fold in;
 *   workspace >= mtp_rvsa_sampling_bwd_workspace_bytes().
 * mtp_full_attn_bwd: dqkv_bf16 fully written; workspace >= mtp_full_attn_bwd_workspace_bytes().
 * ------------------------------------------------------------------------------------------------------------- */
size_t mtp_rvsa_bwd_workspace_bytes(int B, int h, int w, int C, int nH);
int mtp_rvsa_attn_bwd(const void* qkv_bf16, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                      const float* bias_table, const float* lse, const void* dout_bf16, void* dqkv_bf16, float* dparams,
                      float* d_rel_pos_h, float* d_rel_pos_w, float* d_bias_table, float* d_qkv_bias, void* workspace,
                      int scratch_zeroed, int B, int h, int w, int C, int nH, mtp_stream_t stream);
size_t mtp_rvsa_sampling_bwd_workspace_bytes(int B, int h, int w, int C, int nH);
/* mtp_rvsa_attn_bwd + mtp_rvsa_sampling_bwd (dyn = NULL form) of one window block with their four follow-up kernels fused into one launch
   (the reference: autograd of `[V]:372-428` and of the three sampling heads `[V]:228-243`).  dpooled [n_bw][C] fp32 is left in
   sampling_workspace at offset n_bw * 5 * nH floats (n_bw = B * windows per image) for mtp_layernorm_bwd's pool_add. nH even. */
int mtp_rvsa_attn_bwd_fused(const void* qkv_bf16, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                            const float* bias_table, const float* lse, const void* dout_bf16, void* dqkv_bf16, float* dparams,
                            float* d_rel_pos_h, float* d_rel_pos_w, float* d_bias_table, float* d_qkv_bias, void* workspace,
                            const float* pooled, const float* w_off, const float* w_scale, const float* w_angle, float* dw_off,
                            float* db_off, float* dw_scale, float* db_scale, float* dw_angle, float* db_angle,
                            void* sampling_workspace, int B, int h, int w, int C, int nH, mtp_stream_t stream);
int mtp_rvsa_sampling_bwd(const float* dparams, const float* pooled, const float* w_off, const float* w_scale,
                          const float* w_angle, float* dw_off, float* db_off, float* dw_scale, float* db_scale, float* dw_angle,
                          float* db_angle, void* dyn_bf16, void* workspace, int B, int h, int w, int C, int nH,
                          mtp_stream_t stream);
size_t mtp_full_attn_bwd_workspace_bytes(int B, int gh, int gw, int nH);
int mtp_full_attn_bwd(const void* qkv_bf16, const float* rel_pos_h, const float* rel_pos_w, const float* lse,
                      const void* out_bf16, const void* dout_bf16, void* dqkv_bf16, float* d_rel_pos_h, float* d_rel_pos_w,
                      void* workspace, int B, int gh, int gw, int C, int nH, mtp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Optimizer tail over FLAT fp32 buffers (parameters, gradients, Adam moments share one layout; every parameter starts
 * at a multiple of 64 elements).  Replaces clip_grad_norm_(5) + AdamW.step() + CosineAnnealingLR.step()
 * (Multi-Task_Pretrain/main_pretrain.py:786-787,832) and the layer-decay parameter groups
 * (mmcv_custom/layer_decay_optimizer_constructor_vit.py:18-78).
 * state (device, 2 floats): [0] step counter, [1] sum of squares of the gradients.  Per step:
 *   mtp_optim_step_begin(state)            ++step, sumsq = 0
 *   mtp_sumsq_f32(grads, n, state + 1)     (after the gradient all-reduce)
 *   mtp_adamw_step(...)                    grad = g * grad_scale * min(1, max_norm / (sqrt(sumsq) * grad_scale + 1e-6));
 *                                          lr = cosine(step; lr0, eta_min, t_max) * group_lr_scale[group];
 *                                          decoupled weight decay group_weight_decay[group]; p_bf16 (optional) mirrors p.
 * chunk_group[i] = parameter group of elements [64 i, 64 i + 64).
 * ------------------------------------------------------------------------------------------------------------- */
int mtp_optim_step_begin(float* state, mtp_stream_t stream);
int mtp_sumsq_f32(const float* x, size_t n, float* out, mtp_stream_t stream);
int mtp_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, const uint8_t* chunk_group,
                   const float* group_lr_scale, const float* group_weight_decay, const float* state, size_t n, float lr0,
                   float eta_min, int t_max, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                   mtp_stream_t stream);
/* data-parallel variant (bf16 gradient buckets): elements [bf16_from, n) take their (rank-summed) gradient from the bf16
 * buffer g_bf16 (same indexing as g), the rest from the fp32 buffer g.  g_bf16 = NULL: identical to mtp_adamw_step. */
int mtp_adamw_step_mixed(float* p, const float* g, const void* g_bf16, size_t bf16_from, float* m, float* v, void* p_bf16,
                         const uint8_t* chunk_group, const float* group_lr_scale, const float* group_weight_decay, const float* state,
                         size_t n, float lr0, float eta_min, int t_max, float beta1, float beta2, float eps, float max_norm,
                         float grad_scale, mtp_stream_t stream);
int mtp_sumsq_bf16(const void* x_bf16, size_t n, float* out, mtp_stream_t stream);
int mtp_add_f32(const float* in, float* out, size_t n, mtp_stream_t stream);      /* out += in */

/* ---------------------------------------------------------------------------------------------------------------
 * fp32-class forward mode ("fp32x3", precision="fp32x3" on the module; forward only).  Every bf16 tensor of the fast path is stored as
 * hi | lo bf16 word pairs ([rows, 2K]: hi = bf16(v) in columns [0, K), lo = bf16(v - hi) in [K, 2K)); GEMMs run with mtp_epilogue.hilo
 * (three tcgen05 passes, fp32 accumulate), attention in fp32.  Used to meet BASELINE north_star's "forward within 1e-3 rel of the
 * reference" at ViT-L depth -- the reference itself is fp32 (main_pretrain.py never enters autocast) -- and as a full-depth logic check.
 * Same operators as above: [V]:536-539 (patchify), :596 (LayerNorm), :287-433 (RVSA), :90-111 (dense attention), :807-811 (maps). */
int mtp_split_hilo(const float* in, int ld_in, void* out_hilo, size_t rows, int K, mtp_stream_t stream);
int mtp_patchify_hilo(const float* img, void* out_hilo, int B, int cin, int H, int W, mtp_stream_t stream);
/* x: fp32 [rows, C], or hi|lo words where logical row r sits in physical row r / x_sub (pitch x_ld) at column (r % x_sub) * C, lo words
 * x_lo_offset further on; y_hilo: [rows, C | C] */
int mtp_layernorm_fwd_hilo(const void* x, int x_is_hilo, int x_ld, int x_lo_offset, int x_sub, const float* gamma, const float* beta,
                           void* y_hilo, int rows, int C, float eps, int fuse_gelu, mtp_stream_t stream);
int mtp_rvsa_sampling_fwd_hilo(const void* yn_hilo, const float* w_off, const float* b_off, const float* w_scale, const float* b_scale,
                               const float* w_angle, const float* b_angle, float* pooled, float* params, int B, int h, int w, int C, int nH,
                               mtp_stream_t stream);
int mtp_rvsa_attn_fwd_hilo(const void* qkv_hilo, const float* params, const float* rel_pos_h, const float* rel_pos_w,
                           const float* bias_table, void* out_hilo, int B, int h, int w, int C, int nH, mtp_stream_t stream);
int mtp_full_attn_fwd_hilo(const void* qkv_hilo, const float* rel_pos_h, const float* rel_pos_w, void* out_hilo, int B, int gh, int gw,
                           int C, int nH, mtp_stream_t stream);
int mtp_tok_to_nchw_hilo(const void* tok_hilo, int ld, int lo_offset, float* out, int B, int h, int w, int C, int level, mtp_stream_t stream);

/* measurement aid: an empty kernel launch on `stream` (keeps a skipped kernel's place in a captured step; tools/step_breakdown.py) */
int mtp_empty_launch(mtp_stream_t stream);
/* GEMM kernel variant: 1 = persistent warp-specialised kernel, 2 = one tile per CTA with two CTAs per SM (gemm.cu, "Variant 2") */
int mtp_gemm_set_variant(int v);
/* measurement aid (tools/turnaround_probe.py): do-nothing kernel with a configurable footprint; stamps[grid][4] = entry, ready, done (globaltimer ns) */
int mtp_probe_launch(long long* stamps, int grid, int threads, int smem_bytes, int tmem_cols, int spin_ns, int pdl_early, mtp_stream_t stream);
/* same, plus end-of-kernel work: n_loads x 512 B global reads and n_tmem_ld accumulator reads per warp, then store_bytes per CTA written to buf
   as a 128-row tile of a matrix with row pitch ldo bytes (pattern 0: 512 contiguous bytes per warp instruction, 1: 8 rows x 64 B, 2: 32 rows x 16 B,
   3: 2 rows x 256 B) */
int mtp_probe_launch2(long long* stamps, int grid, int threads, int smem_bytes, int tmem_cols, int spin_ns, int pdl_early, void* buf,
                      int store_bytes, int store_pattern, int ldo, int n_tmem_ld, int n_loads, mtp_stream_t stream);
/* tuning aid: cap the depth of the GEMM operand ring (0 = as deep as the shared-memory budget allows) */
int mtp_gemm_set_max_stages(int n);

#ifdef __cplusplus
}
#endif
#endif /* MTP_B200_H_ */
