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
  const float* bias;      /* [N] (PIXSHUF: [ps_cout]) or NULL */
  void* out;
  void* out2;             /* GELU mode: optional pre-activation copy */
  const void* aux;
  const float* row_scale; /* RESID: per-group multiplier (DropPath keep/keep_prob), NULL = 1 */
  int rows_per_group;     /* RESID: tokens per image */
  int pos_rows;           /* POS: rows of the positional table */
  int accumulate;         /* F32: add into out */
  int ps_h, ps_w, ps_cout;
} mtp_epilogue;

int mtp_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N, int K,
                  const mtp_epilogue* ep, int force_bn /* 0 = heuristic */, mtp_stream_t stream);

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
                      float* dgamma, float* dbeta, int rows, int C, int fused_gelu, mtp_stream_t stream);
/* out_bf16[r,c] = in[r,c] * row_scale[r / rows_per_group] (DropPath backward, [V]:31-39); colsum (optional) += column sums
 * of the scaled values (bias gradient of the Linear that produced the branch). */
int mtp_scale_cast_bf16(const float* in, const float* row_scale, int rows_per_group, void* out_bf16, float* colsum,
                        int rows, int C, mtp_stream_t stream);
/* colsum[c] += sum_r in_bf16[r*ld + c]   (bias gradients) */
int mtp_colsum_bf16(const void* in_bf16, int ld, float* colsum, int rows, int C, mtp_stream_t stream);
int mtp_cast_f32_bf16(const float* in, void* out_bf16, size_t n, mtp_stream_t stream);
int mtp_add_bf16_into_f32(const void* in_bf16, float* out, size_t n, mtp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MTP_B200_H_ */
