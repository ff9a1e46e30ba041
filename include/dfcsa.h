/* dfcsa.h — flat C ABI of libdfcsa.so: the B200 (sm_100a) kernels behind the DFC-SA-Res-Block hot path.
 *
 * The reference (YukiHataRin/DFC-SA-UNet) is pure PyTorch and has no FFI of its own; every entry point below
 * replaces a group of ATen call sites on the reference's hot path and cites them (file:line, relative to the
 * reference root).  A maintainer binds these with ctypes; the Python package
 * dfc-sa-unet_b200/dfcsa does exactly that.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer owned by the caller (PyTorch's allocator); the library never
 *     allocates or frees device memory and keeps no pointer past the call;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*) and never synchronises, so
 *     calls are legal under CUDA-graph capture;
 *   - return value 0 = success, otherwise a DFCSA_ERR_* code; dfcsa_last_error() gives the message for the
 *     calling thread.  No exception crosses the boundary and there is no CPU fallback;
 *   - activations are NHWC "pixel-major" matrices: element (pixel m, channel c) lives at base[m*ld + c] where
 *     m = (b*H + h)*W + w and ld >= C is the pixel pitch in elements (a channel slice of a wider concat buffer
 *     is just base+offset with the wide pitch: torch.cat on the reference path becomes zero-copy);
 *   - 16-bit tensors: DFCSA_F16 for forward activations / forward weights, DFCSA_BF16 for gradients.
 */
#ifndef DFCSA_H_
#define DFCSA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFCSA_VERSION 100

enum { DFCSA_OK = 0, DFCSA_ERR_BAD_ARG = 1, DFCSA_ERR_CUDA = 2, DFCSA_ERR_UNSUPPORTED = 3 };
enum { DFCSA_F32 = 0, DFCSA_F16 = 1, DFCSA_BF16 = 2 };
/* how the K axis of one input segment walks the source tensor */
enum {
  DFCSA_TAP_1x1 = 0,   /* one tap, same pixel */
  DFCSA_TAP_3x3 = 1,   /* 9 taps, tap t reads pixel (h + t/3 - 1, w + t%3 - 1), zero outside the image */
  DFCSA_TAP_2x2S2 = 2  /* 4 taps, tap t reads pixel (2h + t/2, 2w + t%2) of a (2H x 2W) source (ConvT dgrad) */
};
enum { DFCSA_OUT_DIRECT = 0, DFCSA_OUT_CONVT2x2 = 1 };
enum { DFCSA_BACKEND_TC = 0, DFCSA_BACKEND_SIMT = 1 };

int         dfcsa_version(void);
const char* dfcsa_last_error(void);
/* 1 if the current device is sm_100 (tcgen05 kernels usable) */
int         dfcsa_device_ok(void);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution:  out[m, n] (+)= sum_seg sum_tap sum_c  src_seg[pix(m, tap), c] * w[n, koff(seg,tap) + c]
 * Replaces F.conv2d 3x3 / 1x1 forward and their input-gradient (dgrad) on
 *   models/unet_dfc_sa_res.py:58 (conv_branch), :66 (attn_branch 1x1), :74 (gate), :81 (fusion_conv),
 *   :88 (residual_conv) and nn.ConvTranspose2d forward/dgrad at :147,150,153,156.
 * w is "packed": [N, Ktot] with K contiguous, K ordered (segment, tap, channel)  (see dfcsa_pack_weights).
 * Backend TC: tcgen05.mma (kind::f16, fp32 accumulate in TMEM), operands staged by TMA, persistent over tiles.
 *             Requires 16-bit src/w, every segment's channel count % 64 == 0, ld % 8 == 0, N % 8 == 0.
 * Backend SIMT: fp32 FMA tiles; any channel counts / dtypes (the Ci=3 first layer, tiny test networks).
 * Optional epilogue: + bias[n]; per-channel double-precision sum / sum-of-squares of the fp32 result
 * (BatchNorm batch statistics, reference :60,67,75,82) accumulated with atomics into stats[0:N], stats[N:2N].
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;     /* source activation / gradient */
  int64_t     ld;      /* pixel pitch in elements */
  int32_t     channels;/* K channels contributed per tap */
  int32_t     tap_mode;/* DFCSA_TAP_* */
} dfcsa_seg_t;

/* BatchNorm finalize folded into the GEMM that produces the statistics (tcgen05 backend only): the LAST CTA to flush its
 * partial sums (a ticket counter, no waiting) turns the complete sums into scale / shift / mean / invstd and updates the
 * running statistics - exactly what dfcsa_bn_finalize does as a launch of its own, minus the launch and the dependency
 * bubble behind every convolution (36 per training step).  Channels = the first `channels` output columns. */
typedef struct {
  const float* gamma; const float* beta; const float* conv_bias;   /* conv_bias optional */
  float* running_mean; float* running_var;                         /* optional */
  float momentum, eps;
  int64_t count;                                                   /* samples per channel (= M) */
  float* scale; float* shift; float* mean; float* invstd;          /* outputs, [channels] each */
  uint32_t* ticket;                                                /* device counter, zero before the launch */
  int32_t channels; int32_t pad_;
} dfcsa_bn_fold_t;

/* Optional fused inference epilogue of dfcsa_conv_gemm (DFCSA_BACKEND_TC, fp16 DIRECT output, no statistics, no accumulate),
 * applied to the result after the bias and the activation:
 *   DFCSA_EPI_GATE_MIX : out = s * p + (1 - s) * q,  s = sigmoid(result).  With p = local_feat, q = attn_feat the gate conv
 *                        writes `fused` itself (reference models/unet_dfc_sa_res.py:104-106); the gate logits never reach HBM
 *   DFCSA_EPI_RESIDUAL : out = result + scale[0] * p.  With p = residual_conv(x) the fusion conv writes the block output
 *                        (reference :110-114); the fusion conv's own output never reaches HBM
 * p, q: fp16 [M, N] views with row pitch ld (elements), 16-byte aligned rows; scale: device pointer to one float. */
enum { DFCSA_EPI_NONE = 0, DFCSA_EPI_GATE_MIX = 1, DFCSA_EPI_RESIDUAL = 2 };
typedef struct {
  int32_t mode; int32_t pad_;
  const void* p; const void* q;
  int64_t ld;
  const float* scale;
} dfcsa_conv_epi_t;

typedef struct {
  int32_t B, H, W;          /* output pixel grid; M = B*H*W */
  int32_t n_seg;
  dfcsa_seg_t seg[3];
  int32_t src_dtype;        /* dtype of all segments */
  int32_t w_dtype;          /* dtype of w (16-bit for TC; F32 for SIMT) */
  const void* w;            /* packed weights [N, Ktot] */
  int32_t N;
  int32_t out_dtype;
  void*   out;
  int64_t ld_out;
  int32_t out_mode;         /* DFCSA_OUT_DIRECT, or DFCSA_OUT_CONVT2x2: n = q*Co + co goes to pixel (2h+q/2, 2w+q%2), channel co */
  int32_t accumulate;       /* out += result (read-modify-write) */
  const float* bias;        /* optional: [N] (DIRECT) or [Co] (CONVT2x2) */
  double* stats;            /* optional: [2*N] */
  void*   shadow;           /* optional bf16 copy of the result at the same coordinates (pitch ld_shadow): the operand
                               the weight-gradient GEMM reads later (kind::f16 cannot mix fp16 and bf16) */
  int64_t ld_shadow;
  /* Epilogue activation, applied after the bias and before the store (inference path: eval-mode BatchNorm is folded into
   * the packed weights and the bias, reference inference.py:100, so conv + BN + ReLU is ONE launch):
   * act = 0 none, 1 ReLU on the output columns n < act_cols (act_cols = 0: all N; the attention-branch and residual
   * 1x1 convs share one GEMM over [W2 ; W5] and only the first half has a ReLU).  DIRECT output mode only. */
  int32_t act;
  int32_t act_cols;
  /* statistics only for the output columns n < stats_cols (0 = all N): the GEMM over [W2 ; W5] feeds a BatchNorm with
   * its first half only, the residual half needs no sums (stats keeps its [2*N] layout) */
  int32_t stats_cols;
  int32_t pad_;
  const dfcsa_bn_fold_t* bn;  /* optional (host pointer, read during the call): fold the BatchNorm finalize into this launch;
                                 needs stats and DFCSA_BACKEND_TC */
  const dfcsa_conv_epi_t* epi; /* optional (host pointer, read during the call): fused inference epilogue, see above */
} dfcsa_conv_params_t;

int dfcsa_conv_gemm(const dfcsa_conv_params_t* p, int backend, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weight gradient:  dw[n, tap*C + c] += alpha * sum_m  dy[pix_dy(m, tap'), n] * x[pix_x(m, tap), c]
 * Replaces the weight-gradient half of conv2d / conv_transpose2d backward for the call sites above.
 *   x_tap_mode  = DFCSA_TAP_3x3 : 3x3 conv weight grad (x shifted by the tap), dy at the output pixel
 *   x_tap_mode  = DFCSA_TAP_1x1 and dy_tap_mode = DFCSA_TAP_1x1 : 1x1 conv
 *   dy_tap_mode = DFCSA_TAP_2x2S2 : ConvTranspose2d 2x2/s2 (dy on the (2H x 2W) grid, x on the (H x W) grid)
 * dw is fp32 [N, taps*C] (row pitch ld_dw) and is accumulated with atomics: zero it first.
 * Backend TC: both operands are MN-major TMA tiles (pixels on the UMMA K axis), split over pixel ranges.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, H, W;          /* pixel grid of x */
  const void* x;  int64_t ld_x;  int32_t C;  int32_t x_dtype;  int32_t x_tap_mode;
  const void* dy; int64_t ld_dy; int32_t N;  int32_t dy_dtype; int32_t dy_tap_mode;
  float*  dw; int64_t ld_dw;
  const float* alpha;       /* optional device scalar multiplier */
  /* optional second gradient that reads the SAME x in the same launch (1x1 taps only), e.g. [W4 ; W3] over z or
   * [W5 ; W2] over the block input:   dw2[n, c - c_begin2] += alpha2 * sum_m dy2[m, n] * x[m, c],  c_begin2 <= c < C */
  const void* dy2; int64_t ld_dy2; int32_t N2; int32_t c_begin2;
  float*  dw2; int64_t ld_dw2;
  const float* alpha2;
} dfcsa_wgrad_params_t;

int dfcsa_conv_wgrad(const dfcsa_wgrad_params_t* p, int backend, void* stream);

/* Host-only (no CUDA call): the pixel-split choice of the tcgen05 weight-gradient launch for `items` independent
 * (tap row, n tile, c tile) work items over `pix_blocks` 64-pixel blocks on a device with `sms` SMs.  The kernel runs
 * one CTA per SM, so the grid items * splits is chosen to fill 1..4 WHOLE waves (a grid of 2*sms + 1 CTAs would cost
 * three).  Exposed so that the choice can be regression-tested without a GPU. */
int dfcsa_wgrad_plan(int64_t items, int64_t pix_blocks, int32_t sms, int32_t* splits, int64_t* blocks_per_split);

/* dst[i0*ld_dst + i1*D2 + i2] = scale * src[i0*s0 + i1'*s1 + i2*s2], i1' = flip ? D1-1-i1 : i1
 * (ld_dst = 0 means a dense destination, ld_dst = D1*D2).
 * Re-lays fp32 master weights ([Co,Ci,kh,kw], ConvT [Ci,Co,2,2]) into the packed K-major GEMM operands
 * (forward: [Co,(tap,ci)]; dgrad: [Ci,(flipped tap,co)]) and packed weight gradients back. */
int dfcsa_permute3(const void* src, int src_dtype, void* dst, int dst_dtype,
                   int64_t D0, int64_t D1, int64_t D2, int64_t s0, int64_t s1, int64_t s2,
                   int flip1, const float* scale, int64_t ld_dst, void* stream);

/* The same re-layout for a whole network in ONE launch: a device table of jobs (all weight tensors of the model,
 * built once by the host) and the prefix sum of their sizes in 1024-element chunks (n_jobs + 1 entries). */
typedef struct {
  const void* src; void* dst; const float* scale;
  int64_t D0, D1, D2, s0, s1, s2, ld_dst;
  int32_t src_dtype, dst_dtype, flip1, pad_;
  const float* row_scale;   /* optional [D0]: dst row i0 is additionally multiplied by row_scale[i0] (BatchNorm scale folded
                               into the output channels of a forward weight matrix for the inference path) */
} dfcsa_pack_job_t;
int dfcsa_pack_jobs(const dfcsa_pack_job_t* jobs_dev, int32_t n_jobs, const int64_t* chunk_prefix_dev,
                    int64_t total_chunks, void* stream);

/* Strided batched fp32 GEMM  C[b] = alpha * A[b] * B[b] + beta * C[b]  with arbitrary element strides
 * (transposes are strides).  Used for the pooled attention products, reference
 * models/unet_dfc_sa_res.py:28-33 (q/k/v 1x1 convs on the pooled map, bmm(Q,K), bmm(V,A^T)) and their backward. */
typedef struct {
  int32_t batch, M, N, K;
  const float* A; int64_t a_b, a_m, a_k;
  const float* B; int64_t b_b, b_k, b_n;
  float*       C; int64_t c_b, c_m, c_n;
  const float* bias_n;      /* optional bias along N */
  const float* bias_m;      /* optional bias along M */
  float alpha, beta;
} dfcsa_sgemm_params_t;
int dfcsa_sgemm(const dfcsa_sgemm_params_t* p, void* stream);

/* row softmax over the last dim (reference :31) and its backward  dS = A * (dA - rowsum(dA*A)).
 * x / dy are fp32; y (and dx) may be fp32, fp16 or bf16 (the 16-bit forms feed the tensor-core attention GEMMs). */
int dfcsa_softmax_rows(const float* x, void* y, int y_dtype, int64_t rows, int32_t cols, void* stream);
int dfcsa_softmax_rows_bwd(const void* y, int y_dtype, const float* dy, void* dx, int dx_dtype, int64_t rows, int32_t cols,
                           void* stream);

/* One-pass softmax backward for the large attention maps: dx = y * (dy - D[row]) with D[row] = sum_j dy*y supplied as
 * sum_c dO[row, c] * O[row, c] (dfcsa_rowdot; identical because O = P V).  y / dy / dx: fp32 or 16-bit. */
int dfcsa_softmax_rows_bwd_d(const void* y, int y_dtype, const void* dy, int dy_dtype, const float* D, void* dx, int dx_dtype,
                             int64_t rows, int32_t cols, void* stream);
/* out[r] = sum_c a[r, c] * b[r, c]  (fp32, dense rows) */
int dfcsa_rowdot(const float* a, const float* b, int64_t rows, int32_t cols, float* out, void* stream);

/* The whole attention core for small pooled maps (N = P*P <= 32; P = 4 in the DFC-SA-Res-Block configs) in one kernel
 * per direction, one CTA per image, fp32: attn = softmax(q k^T), o = attn v (reference models/unet_dfc_sa_res.py:28-34)
 * and dq, dk, dv from d_o.  qkv / dqkv: [B*N, ld] rows (q[0:Cq] | k[Cq:2Cq] | v[2Cq:2Cq+C]); attn [B,N,N]; o, d_o [B*N, C]. */
int dfcsa_attn_small_fwd(const float* qkv, int64_t ld, int32_t B, int32_t N, int32_t Cq, int32_t C,
                         float* attn, float* o, void* stream);
/* dbq / dbk / dbv (optional, fp32 [Cq] / [Cq] / [C], ACCUMULATED with atomics): the bias gradients of the q / k / v
 * convolutions = column sums of dq / dk / dv over tokens and images, so no separate reduction pass is needed. */
int dfcsa_attn_small_bwd(const float* qkv, int64_t ld, const float* attn, const float* d_o, int32_t B, int32_t N,
                         int32_t Cq, int32_t C, float* dqkv, float* dbq, float* dbk, float* dbv, void* stream);

/* Batched GEMM on tcgen05:  C[b] = A[b] * B[b],  b < batch,  A: M x K, B: K x N (as a matrix product), 16-bit operands of
 * one dtype, fp32 accumulation.  Storage of an operand is either K-major (element (m,k) at base + b*a_b + m*ld_a + k)
 * or MN-major (element (m,k) at base + b*a_b + k*ld_a + m): transposes are free.  C is row-major [M, N] with pitch ld_c.
 * Used for the attention products softmax(Q K^T) V and their backward (reference models/unet_dfc_sa_res.py:30-33,
 * models/unet_dfc_sa_ablation_attention.py:20-24) whenever N = P*P >= 64.  Pitches / batch strides in elements,
 * multiples of 8; N % 8 == 0. */
/* Optional epilogues that keep the [N, N] intermediates of attention out of HBM:
 *   ROWSTATS:    nothing is stored in C; every tile writes (max, sum exp(x - max)) of its part of each row to
 *                rowstat[b, part, m, 2] (dfcsa_bgemm_rowstat_parts(N) parts per row); dfcsa_lse_combine -> lse[b, m];
 *   EXP:         C = exp(A B - rowvec[b, m])       rowvec = lse: the normalised probabilities, usually 16-bit;
 *   SOFTMAX_BWD: C = aux * (A B - rowvec[b, m])    aux = probabilities (16-bit, laid out like C; may alias C when the
 *                element sizes match), rowvec = D = rowdot(dO, O): with A B = dO V^T this is dS, the softmax backward. */
enum { DFCSA_BGEMM_EPI_NONE = 0, DFCSA_BGEMM_EPI_ROWSTATS = 1, DFCSA_BGEMM_EPI_EXP = 2, DFCSA_BGEMM_EPI_SOFTMAX_BWD = 3 };
typedef struct {
  int32_t batch, M, N, K;
  const void* A; int64_t a_b, ld_a; int32_t a_mn_major;
  const void* B; int64_t b_b, ld_b; int32_t b_mn_major;
  int32_t ab_dtype;
  void* C; int64_t c_b, ld_c; int32_t c_dtype;
  int32_t epi_mode;
  float* rowstat;
  const float* rowvec;
  const void* aux; int32_t aux_dtype;
} dfcsa_bgemm_params_t;
int dfcsa_bgemm(const dfcsa_bgemm_params_t* p, void* stream);
int dfcsa_bgemm_rowstat_parts(int32_t N);
int dfcsa_lse_combine(const float* rowstat, int32_t parts, int32_t batch, int32_t M, float* lse, void* stream);

/* Fused attention forward for large N:  o[b] = exp(q k^T - lse) v  with the [N, N] probabilities kept on chip (TMEM ->
 * registers -> shared memory as the A operand of the second tcgen05 product) - reference models/unet_dfc_sa_res.py:30-33,
 * models/unet_dfc_sa_ablation_attention.py:20-24.  qkv: fp16 rows (q[Cq] | k[Cq] | v[C]) with pitch ld, [batch*N] rows;
 * lse: [batch*N] row log-sum-exp of q k^T (DFCSA_BGEMM_EPI_ROWSTATS pass + dfcsa_lse_combine); o: fp32 [batch, N, C].
 * 8 <= Cq <= 64, C in {64, 128}. */
int dfcsa_attn_pv_fused(const void* qkv, int64_t ld, int32_t batch, int32_t N, int32_t Cq, int32_t C, const float* lse,
                        float* o, void* stream);

/* Fused attention backward for large N (autograd of the same reference lines): with P = exp(q k^T - lse) and
 * D = rowdot(dO, O):  dV = P^T dO,  dS = P o (dO V^T - D),  dQ = dS K,  dK = dS^T Q - two tcgen05 kernels (one per
 * accumulation direction) that rebuild P and dS tile by tile on chip; no [N, N] tensor is read or written.
 * qkv16: the forward's fp16 rows (q | k | v), pitch ld16; qkvb: the same rows in bf16, pitch ldb; dO: bf16 [batch*N, C],
 * pitch ld_do; dqkv: fp32 rows (dq[Cq] | dk[Cq] | dv[C]) with pitch ld_out.  Cq in {8,...,32}, C in {64, 128}, N % 8 == 0. */
int dfcsa_attn_bwd_fused(const void* qkv16, int64_t ld16, const void* qkvb, int64_t ldb, const void* dO, int64_t ld_do,
                         int32_t batch, int32_t N, int32_t Cq, int32_t C, const float* lse, const float* D,
                         float* dqkv, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BatchNorm2d (reference :60,67,75,82; ATen batch_norm semantics: biased variance for normalisation, unbiased
 * for running_var, momentum 0.1, eps 1e-5).
 * finalize: from the conv epilogue's double sums -> per-channel scale/shift (y = x*scale + shift), saved
 *   mean / invstd for backward, running-stat update.  conv_bias (which the conv kernels do not add in front of
 *   a train-mode BN because it cancels) is folded into the running mean so eval mode stays exact.
 * eval:     scale/shift from running statistics (+ conv bias).
 * ---------------------------------------------------------------------------------------------------------- */
int dfcsa_bn_finalize(const double* sum, const double* sumsq, int64_t count, int32_t C,
                      const float* gamma, const float* beta, const float* conv_bias,
                      float* running_mean, float* running_var, float momentum, float eps,
                      float* scale, float* shift, float* mean, float* invstd, void* stream);
int dfcsa_bn_eval_affine(int32_t C, const float* gamma, const float* beta, const float* conv_bias,
                         const float* running_mean, const float* running_var, float eps,
                         float* scale, float* shift, void* stream);

/* y = a [+ b] + res_scale * r with the optional 2x2 max pool: the output stage of the ablation blocks that end in a plain sum
 * (AdditionFusionBlock, reference models/unet_dfc_sa_ablation_fusion.py:40-55; AttentionOnlyBlock,
 * models/unet_dfc_sa_ablation_branches.py:60-69).  b may be NULL. */
int dfcsa_sum_out_fwd(const void* a, int64_t ld_a, const void* b, int64_t ld_b, const void* r, int64_t ld_r, int32_t B,
                      int32_t H, int32_t W, int32_t C, const float* res_scale, void* y, int64_t ld_y, void* yp,
                      int64_t ld_yp, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused bandwidth kernels of the DFC-SA block forward (reference models/unet_dfc_sa_res.py:95-116, :20-39).
 * All activations fp16 NHWC, per-channel BN affine (scale, shift) fp32.
 * ---------------------------------------------------------------------------------------------------------- */
/* a = relu(bn2(A0)); pooled = adaptive_avg_pool2d(a, P) (reference :24).  Separable: rows then columns.
 * tmp is [B, H, P, C] fp32 scratch; pooled is [B, P, P, C] fp32.
 * with_masks != 0 (training): tmp and pooled hold THREE planes, [3][B, H, P, C] and [3][B, P, P, C]: plane 0 as above,
 * plane 1 = the window means of m = [bn2(A0) > 0], plane 2 = the window means of m * A0 - what dfcsa_pool_window_terms
 * needs in the backward pass. */
int dfcsa_bnrelu_pool_fwd(const void* a0, int64_t ld, int32_t B, int32_t H, int32_t W, int32_t C,
                          const float* scale, const float* shift, int32_t P,
                          float* tmp, float* pooled, int32_t with_masks, void* stream);
/* L = relu(bn1(L0)) -> z[:, C:2C];  A = gamma*bilinear_up(o) + relu(bn2(A0)) -> z[:, 2C:3C]
 * (reference :97,:99,:36,:38).  o is [B, P, P, C] fp32.  zb: optional bf16 shadow of z (same channel offsets).
 * l0 may be NULL: only the A half is computed (inference path: L comes straight out of the folded conv + ReLU). */
int dfcsa_branch_act_fwd(const void* l0, int64_t ld_l0, const void* a0, int64_t ld_a0,
                         int32_t B, int32_t H, int32_t W, int32_t C,
                         const float* scale1, const float* shift1, const float* scale2, const float* shift2,
                         const float* o, int32_t P, const float* gamma,
                         void* z, int64_t ld_z, void* zb, int64_t ld_zb, void* stream);
/* g = sigmoid(bn3(G0)); z[:, 0:C] = g*L + (1-g)*A  (reference :104,:106); zb: optional bf16 shadow of z */
int dfcsa_gate_mix_fwd(const void* g0, int64_t ld_g0, int64_t M, int32_t C,
                       const float* scale3, const float* shift3, void* z, int64_t ld_z, void* zb, int64_t ld_zb,
                       void* stream);
/* y = relu(bn4(F0)) + res_scale*R (reference :110,:114) and optionally yp = maxpool2x2(y) (reference :164);
 * yb / ypb: optional bf16 shadows */
int dfcsa_block_out_fwd(const void* f0, int64_t ld_f0, const void* r, int64_t ld_r,
                        int32_t B, int32_t H, int32_t W, int32_t C,
                        const float* scale4, const float* shift4, const float* res_scale,
                        void* y, int64_t ld_y, void* yp, int64_t ld_yp,
                        void* yb, int64_t ld_yb, void* ypb, int64_t ld_ypb, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Block backward (autograd of the same lines).  Gradients bf16 NHWC; per-channel reductions in double.
 * ---------------------------------------------------------------------------------------------------------- */
/* dy = dskip (optional) + maxpool-backward(dyp) (optional; argmax recomputed from y);
 * d4 = dy * [bn4(F0) > 0];  red4[0:C] += sum d4, red4[C:2C] += sum d4*xhat4;  *drs += sum dy*R.
 * dy_out (bf16) receives the combined dy when it has more than one source (may alias dskip). */
int dfcsa_block_out_bwd_reduce(const void* dskip, int64_t ld_dskip, const void* dyp, int64_t ld_dyp,
                               const void* y, int64_t ld_y,
                               const void* f0, int64_t ld_f0, const void* r, int64_t ld_r,
                               int32_t B, int32_t H, int32_t W, int32_t C,
                               const float* scale4, const float* shift4, const float* mean4, const float* invstd4,
                               void* dy_out, int64_t ld_dy, double* red4, double* drs, void* stream);
/* generic BN(+activation) backward apply:
 *   d = dy * act'(.)   with act = relu (mode 0: mask from x*scale+shift > 0) or none (mode 2)
 *   dx = gamma*invstd * (d - red[0:C]/count - xhat * red[C:2C]/count)   (bf16) */
int dfcsa_bn_bwd_apply(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t M, int32_t C,
                       const float* scale, const float* shift, const float* mean, const float* invstd,
                       const float* gamma, const double* red, int32_t act_mode,
                       void* dx, int64_t ld_dx, void* stream);
/* gate/mix backward, pass 1: with dz = [df | dL' | dA'] (gradient of the fusion conv input), z = [f | L | A]:
 *   g = sigmoid(bn3(G0)); dS = df*(L-A)*g*(1-g); red3 += (sum dS, sum dS*xhat3) */
int dfcsa_gate_mix_bwd_reduce(const void* dz, int64_t ld_dz, const void* z, int64_t ld_z,
                              const void* g0, int64_t ld_g0, int64_t M, int32_t C,
                              const float* scale3, const float* shift3, const float* mean3, const float* invstd3,
                              double* red3, void* stream);
/* pass 2: dG0 = gamma3*invstd3*(dS - mean(dS) - xhat3*mean(dS*xhat3)) -> dg0 (bf16).  Only df = dz[:, 0:C] is read: the
 * dL / dA columns of dz are produced afterwards by one two-segment dgrad GEMM over [dF0 | dG0] (no read-modify-write). */
int dfcsa_gate_mix_bwd_apply(const void* dz, int64_t ld_dz, const void* z, int64_t ld_z,
                             const void* g0, int64_t ld_g0, int64_t M, int32_t C,
                             const float* scale3, const float* shift3, const float* mean3, const float* invstd3,
                             const float* gamma3, const double* red3,
                             void* dg0, int64_t ld_dg0, void* stream);
/* branch backward, pass 1.  First completes the gradients of the two branches in place with the gate-mix terms:
 *   dL = dz[:,C:2C] += df*g,  dA = dz[:,2C:3C] += df*(1-g),  g = sigmoid(bn3(G0)), df = dz[:,0:C];  then
 *   red1 += (sum d1, sum d1*xhat1) with d1 = dL*[L>0];
 *   dgamma += sum dA*U (U = bilinear_up(o));  do = gamma * bilinear_up^T(dA)  ([B,P,P,C] fp32, separable, tmp [B,H,P,C]) */
int dfcsa_branch_bwd_reduce1(void* dz, int64_t ld_dz, const void* l0, int64_t ld_l0, const void* g0, int64_t ld_g0,
                             int32_t B, int32_t H, int32_t W, int32_t C,
                             const float* scale1, const float* shift1, const float* mean1, const float* invstd1,
                             const float* scale3, const float* shift3,
                             const float* o, int32_t P, const float* gamma,
                             double* red1, double* dgamma, float* tmp, float* d_o, void* stream);
/* pass 2 (after the pooled-attention backward produced dpooled [B,P,P,C]):
 *   da = dA + adaptive_avg_pool^T(dpooled);  d2 = da*[a>0];  red2 += (sum d2, sum d2*xhat2) */
int dfcsa_branch_bwd_reduce2(const void* dz, int64_t ld_dz, const void* a0, int64_t ld_a0,
                             int32_t B, int32_t H, int32_t W, int32_t C,
                             const float* scale2, const float* shift2, const float* mean2, const float* invstd2,
                             const float* dpooled, int32_t P, double* red2, void* stream);
/* The same reduction without a gather pass.  adaptive_avg_pool^T is linear, so its part of the two sums only needs the
 * per-window means the forward pooling pass can emit (dfcsa_bnrelu_pool_fwd with_masks):
 *   sum_pix poolT(dp)[pix] m[pix] = sum_w dp[w] mean_w(m),   sum_pix poolT(dp)[pix] m[pix] A0[pix] = sum_w dp[w] mean_w(m A0)
 * bn_bwd_reduce:     red += (sum d m, sum d m xhat), m = [x*scale+shift > 0], over a plain (dy, x) pair (here: dA, A0);
 * pool_window_terms: red += the window part; means = planes 1 and 2 of the forward's pooled buffer ([2][B, P, P, C]). */
int dfcsa_bn_bwd_reduce(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, int64_t M, int32_t C,
                        const float* scale, const float* shift, const float* mean, const float* invstd, double* red,
                        void* stream);
int dfcsa_pool_window_terms(const float* dpooled, const float* means, int32_t B, int32_t P, int32_t C, const float* mean,
                            const float* invstd, double* red, void* stream);
/* pass 3: dL0, dA0 (bf16) from the two reductions */
int dfcsa_branch_bwd_apply(const void* dz, int64_t ld_dz, const void* l0, int64_t ld_l0,
                           const void* a0, int64_t ld_a0, int32_t B, int32_t H, int32_t W, int32_t C,
                           const float* scale1, const float* shift1, const float* mean1, const float* invstd1,
                           const float* gamma1, const double* red1,
                           const float* scale2, const float* shift2, const float* mean2, const float* invstd2,
                           const float* gamma2, const double* red2,
                           const float* dpooled, int32_t P,
                           void* dl0, int64_t ld_dl0, void* da0, int64_t ld_da0, void* stream);
/* BN affine gradients from a reduction: dgamma = red[C:2C], dbeta = red[0:C] (fp32 out) */
int dfcsa_bn_param_grads(const double* red, int32_t C, float* dgamma, float* dbeta, void* stream);
/* The same for ALL small parameter gradients of one block in one launch.  red is the block's backward reduction buffer
 * [red1 (2C) | red2 (2C) | red3 (2C) | red4 (2C) | d res_scale | d gamma] (double); output k of BatchNorm i is written only
 * when its pointer is non-NULL: dbeta_i = red_i[0:C], dgamma_i = red_i[C:2C]; *drs = red[8C], *dgam = red[8C+1]. */
int dfcsa_block_param_grads(const double* red, int32_t C, float* dg1, float* db1, float* dg2, float* db2, float* dg3,
                            float* db3, float* dg4, float* db4, float* drs, float* dgam, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Layout / misc
 * ---------------------------------------------------------------------------------------------------------- */
/* NCHW fp32 [B,C,H,W] -> NHWC (dst dtype, pitch ld) and back (reference tensors at the module boundary) */
int dfcsa_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int64_t ld,
                       int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
int dfcsa_nhwc_to_nchw(const void* src, int src_dtype, int64_t ld, float* dst,
                       int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
/* per-channel column sums of a [M, C] matrix (bias gradients): out[c] += sum_m x[m,c] */
int dfcsa_colsum(const void* x, int x_dtype, int64_t ld, int64_t M, int32_t C, float* out, void* stream);
/* y = cast(x) elementwise on [M, C] matrices with pitches */
int dfcsa_cast2d(const void* x, int x_dtype, int64_t ld_x, void* y, int y_dtype, int64_t ld_y,
                 int64_t M, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * General bilinear re-size and adaptive average pool on NHWC tensors of any storage dtype (one thread per output, no
 * tuning: both users are OFF the 224 / 512 / 1024 hot path).
 *   - the reference's re-size after a ConvTranspose2d whose output does not match the skip tensor, i.e. inputs whose side
 *     is not a multiple of 16 (models/unet_dfc_sa_res.py:180-181, F.interpolate(..., mode="bilinear", align_corners=False));
 *   - LightSelfAttention / FullResolutionAttention called as modules of their own (models/unet_dfc_sa_res.py:20-39):
 *     adaptive_avg_pool2d (:24) without the BatchNorm + ReLU the block kernels fuse in front of it, and
 *     gamma * interpolate(out) + x (:36-38).
 * resize:      dst[b,y,x,c] = alpha * bilinear(src: Hi x Wi -> Ho x Wo)[b,y,x,c] + add[b,y,x,c]   (alpha: optional device
 *              scalar, add: optional tensor laid out like dst)
 * resize_bwd:  dsrc = alpha * bilinear^T(ddst)          (every element of dsrc is written)
 * pool:        pooled[b,i,j,c] (fp32, dense [B,P,P,C]) = adaptive_avg_pool2d(src, P)
 * pool_bwd:    dst = add + adaptive_avg_pool2d^T(dpooled)   (add optional)
 * ---------------------------------------------------------------------------------------------------------- */
int dfcsa_resize_bilinear(const void* src, int src_dtype, int64_t ld_src, int32_t B, int32_t Hi, int32_t Wi, int32_t C,
                          void* dst, int dst_dtype, int64_t ld_dst, int32_t Ho, int32_t Wo, const float* alpha,
                          const void* add, int add_dtype, int64_t ld_add, void* stream);
int dfcsa_resize_bilinear_bwd(const void* ddst, int ddst_dtype, int64_t ld_ddst, int32_t B, int32_t Hi, int32_t Wi, int32_t C,
                              void* dsrc, int dsrc_dtype, int64_t ld_dsrc, int32_t Ho, int32_t Wo, const float* alpha,
                              void* stream);
int dfcsa_adaptive_pool(const void* src, int src_dtype, int64_t ld_src, int32_t B, int32_t H, int32_t W, int32_t C, int32_t P,
                        float* pooled, void* stream);
int dfcsa_adaptive_pool_bwd(const float* dpooled, int32_t B, int32_t H, int32_t W, int32_t C, int32_t P, const void* add,
                            int add_dtype, int64_t ld_add, void* dst, int dst_dtype, int64_t ld_dst, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * bce_dice loss (reference utils/trainer.py:124 sigmoid; utils/metrics.py:74-78 BCELoss + dice_loss :19-24;
 * hard metrics :228-236).  sums (double[8]): 0 bce_sum, 1 sum p*t, 2 sum p, 3 sum t, 4 sum [p>.5]*t, 5 sum [p>.5].
 * from_logits=1: x are logits (p = sigmoid(x)); 0: x are probabilities (the calculate_metrics boundary).
 * ---------------------------------------------------------------------------------------------------------- */
int dfcsa_bce_dice_sums(const float* x, const float* t, int64_t n, int from_logits, double* sums, void* stream);
/* out[0]=loss, out[1]=bce, out[2]=dice_loss, out[3]=hard iou, out[4]=hard dice  (fp32, device) */
int dfcsa_bce_dice_finalize(const double* sums, int64_t n, float w_bce, float w_dice, float smooth,
                            float* out, void* stream);
/* The same per SAMPLE (validation: the reference re-runs calculate_metrics on every sample of every batch and copies each
 * to the host, utils/trainer.py:229-245): x, t are [n_samples, n_per_sample]; sums double[n_samples][8] (zeroed by the
 * caller); out fp32 [n_samples][5] as above. */
int dfcsa_bce_dice_sums_batched(const float* x, const float* t, int64_t n_per_sample, int32_t n_samples, int from_logits,
                                double* sums, void* stream);
int dfcsa_bce_dice_finalize_batched(const double* sums, int64_t n_per_sample, int32_t n_samples, float w_bce, float w_dice,
                                    float smooth, float* out, void* stream);
/* dx = gout * dloss/dx  (wrt logits if from_logits else wrt probabilities); gout optional device scalar */
int dfcsa_bce_dice_bwd(const float* x, const float* t, int64_t n, int from_logits, const double* sums,
                       float w_bce, float w_dice, float smooth, const float* gout,
                       void* dx, int dx_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer (reference utils/trainer.py:149 clip_grad_norm_(1.0); train.py:73-78 SGD momentum/weight decay).
 * Works on a device table of tensor descriptors so one launch covers all 235 parameter tensors.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { float* w; float* g; float* m; int64_t n; } dfcsa_param_t;
/* sumsq[0] += sum over all tensors of g^2 (double) */
int dfcsa_grad_sumsq(const dfcsa_param_t* table_dev, int32_t n_tensors, int64_t max_n, double* sumsq, void* stream);
/* c = min(1, max_norm/(sqrt(sumsq*gscale^2)+1e-6)); g = c*gscale*g + wd*w; m = first ? g : mom*m + g; w -= lr*m.
 * If sumsq is NaN / inf the step is skipped entirely (the reference skips a batch whose loss is NaN,
 * utils/trainer.py:134-139; here the decision is taken on the device from the reduced gradient norm). */
int dfcsa_sgd_step(const dfcsa_param_t* table_dev, int32_t n_tensors, int64_t max_n, const double* sumsq,
                   float gscale, float max_norm, float lr, float momentum, float weight_decay, int first_step,
                   void* stream);

/* dst[i] += src[i] over flat fp32 buffers: gradient accumulation across the micro-batches of one optimizer step (the C4
 * configuration's 128 / 256 images per GPU at 512^2 exceed one forward's activation memory; not in the reference) */
int dfcsa_accumulate(float* dst, const float* src, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * GPU-side data path: the reference's per-sample transform chain, utils/data_loader.py:25-74
 * (ExtResize -> ExtRandomRotation -> ExtRandomHorizontalFlip -> ExtToTensor -> ExtNormalize, executed there by Pillow and
 * torchvision in DataLoader workers), for a ragged batch of decoded uint8 images.  Bit-exact with Pillow 12.2: 22-bit
 * fixed-point separable BILINEAR resize with byte intermediates, NEAREST resize by repeated-addition coordinates,
 * double-precision BILINEAR rotation truncated to bytes, 16.16 fixed-point NEAREST rotation, then
 * image = (byte / 255 - mean) / std (fp32, NCHW) and mask = (byte / 255 > 0.5).
 * The random decisions stay on the host (the caller draws them in the reference's order and, like Image.rotate, turns an
 * angle into rot_mode + the inverse affine matrix; dfcsa/data_loader.py: rotate_plan). */
enum { DFCSA_ROT_NONE = 0, DFCSA_ROT_AFFINE = 1, DFCSA_ROT_90 = 2, DFCSA_ROT_180 = 3, DFCSA_ROT_270 = 4 };
typedef struct {
  const uint8_t* img;     /* device, [h, w, 3] RGB bytes, tightly packed */
  const uint8_t* mask;    /* device, [h, w] bytes (mode "L"), or NULL */
  int32_t h, w;           /* source size */
  int32_t rot_mode;       /* DFCSA_ROT_*; 90 / 270 only for square outputs (as in Image.rotate) */
  int32_t flip;           /* horizontal flip, applied after the rotation */
  double a[6];            /* DFCSA_ROT_AFFINE: output -> input matrix in the resized image's pixel coordinates */
} dfcsa_sample_t;
int64_t dfcsa_preprocess_workspace_bytes(int32_t n, int32_t max_src_h, int32_t max_src_w, int32_t out_h, int32_t out_w);
/* samples_dev: n records in device memory; mean3 / std3: host arrays; img_out [n, 3, out_h, out_w] fp32,
 * mask_out [n, 1, out_h, out_w] fp32 or NULL; workspace: 256-byte aligned device scratch of at least
 * dfcsa_preprocess_workspace_bytes(...) bytes.  Four launches, no synchronisation. */
int dfcsa_preprocess(const dfcsa_sample_t* samples_dev, int32_t n, int32_t max_src_h, int32_t max_src_w, int32_t out_h,
                     int32_t out_w, const float* mean3, const float* std3, float* img_out, float* mask_out,
                     void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFCSA_H_ */
