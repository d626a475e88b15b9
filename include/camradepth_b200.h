/*
 * camradepth_b200 -- C ABI of the B200 (sm_100a) kernels behind the CamRaDepth hot path.
 *
 * The reference (TUMFTM/CamRaDepth) has NO native/FFI layer: every op on the path is an eager
 * ATen call made from Python.  The boundary a maintainer binds is therefore this
 * library, loaded with ctypes from the drop-in `CamRaDepth` nn.Module.  Each
 * entry point names the reference code it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless stated; no allocation
 *     inside the library (workspaces are passed in); all launches go to `stream`.
 *   - return 0 on success, a cudaError_t (>0) on a launch error, < 0 on an argument error.
 *   - activations are NHWC ("tokens" (B,N,C) are NHWC with H*W = N); `ld*` = elements between
 *     consecutive pixels, so a tensor may be a channel slice of a wider concat buffer.  Channel
 *     counts handed to kernels are multiples of 8 (callers pad with zero channels).
 *   - dtype codes: CRD_F32 = 0, CRD_BF16 = 1.  Statistics, residual streams, losses, optimizer
 *     state and parameter gradients are always fp32.
 */
#ifndef CAMRADEPTH_B200_H
#define CAMRADEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* crd_stream_t; /* cudaStream_t */

#define CRD_API __attribute__((visibility("default")))

#define CRD_F32 0
#define CRD_BF16 1
#define CRD_ACT_NONE 0
#define CRD_ACT_GELU 1
#define CRD_ACT_SIGMOID 2

CRD_API int crd_version(void);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
CRD_API unsigned long long crd_launch_count(void);
/* 1 if the tcgen05/TMA kernels can run on the current device (sm_100) */
CRD_API int crd_has_tcgen05(void);

/* ---------------------------------------------------------------- layout (boundary NCHW <-> NHWC)
 * Replaces nothing in the reference (which is NCHW throughout); this is the price of the drop-in
 * NCHW fp32 nn.Module surface (CamRaDepth.py:173-176). */
CRD_API int crd_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int B, int C, int H, int W, int ld_dst,
                     crd_stream_t stream);
CRD_API int crd_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int B, int C, int H, int W, int ld_src,
                     crd_stream_t stream);

/* ---------------------------------------------------------------- convolution as implicit GEMM
 * y[b,oh,ow,n] = sum_{kh,kw,c} x[b, oh*stride-pad+kh, ow*stride-pad+kw, c] * w[n][(kh*KW+kw)*Cin + c]
 * (transposed=1: x is read at ((oh+pad-kh)/stride, (ow+pad-kw)/stride) when divisible -- the data
 * gradient of a strided conv).  Replaces F.conv2d / F.conv1d at simplified_attention.py:35,41,92,
 * 98,101,108,184,321; utils.py:225,286,288; CamRaDepth.py:129,133,155,159 and their autograd
 * backward (aten convolution_backward). */
typedef struct {
  int B, H, W;          /* tensor being read */
  int Cin, ldx;         /* channels read (multiple of 8), pixel stride */
  int Ho, Wo, Cout;     /* tensor being written */
  int ldy;
  int KH, KW, stride, pad;
  int transposed;
  int in_dtype;         /* dtype of x and w */
  int out_dtype;        /* dtype of y (fwd) / dy (wgrad) */
  int act;              /* CRD_ACT_NONE or CRD_ACT_SIGMOID applied after bias */
  int accumulate;       /* y += result */
  int out_nchw;         /* write y as NCHW (B,Cout,Ho,Wo); ldy ignored */
  int w_tap_stride;     /* tensor-core path only: elements between taps in a weight row (0 = Cin) ...       */
  int w_koff;           /* ... and first column used inside each tap: w[n][tap*w_tap_stride + w_koff + c]   */
} crd_conv_desc;

/* generic CUDA-core path (any shape; the only path in fp32-exact mode) */
CRD_API int crd_conv_fwd(const crd_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                 crd_stream_t stream);
/* dw[n][(kh*KW+kw)*Cin + c] += sum_pixels dy[.,n] * x[.,c]  (fp32, atomically accumulated; caller zeroes) */
CRD_API int crd_conv_wgrad(const crd_conv_desc* d, const void* x, const void* dy, float* dw, crd_stream_t stream);

/* tcgen05 / TMEM / TMA path (bf16 in, fp32 accumulate): stride-1 KHxKW "same" convs and 1x1 GEMMs.
 * Same contract as crd_conv_fwd (transposed=1 gives the stride-1 dgrad).  gn_sums (optional, fp32
 * [B][Cout][2]) receives per-(sample, channel) sum / sum-of-squares of the fp32 accumulators. */
CRD_API int crd_conv_fwd_tc(const crd_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                    float* gn_sums, crd_stream_t stream);
/* Seg_Block heads whose logits are only consumed by argmax (seg_conv_stage_4 / unsup_stage_4 / unsup_final,
 * CamRaDepth.py:128-135,155-162 with utils.py:95-100): 3x3 conv (Cout <= 128) whose accumulator read-out takes the
 * per-pixel argmax over the first ncls channels (first maximum wins) and writes argmax / ncls to up to two bf16
 * NHWC channels (map0 / map1, pixel strides ld0 / ld1 in elements) and / or an fp32 (B,1,H,W) map; the logits
 * are never materialised. */
CRD_API int crd_conv_argmax_tc(const crd_conv_desc* d, const void* x, const void* w, const float* bias, int ncls,
                       void* map0, int ld0, void* map1, int ld1, float* map_f32, crd_stream_t stream);
CRD_API int crd_conv_wgrad_tc(const crd_conv_desc* d, const void* x, const void* dy, float* dw, crd_stream_t stream);
/* 1x1 contractions with a bias (encoder q / k / fc1 / fc2 Conv1d, simplified_attention.py:16-21,66-67): the weight
 * gradient and the bias gradient db[co] += sum_pixels dY from one pass over dY (an extra MMA against ones). */
CRD_API int crd_conv_wgrad_bias_tc(const crd_conv_desc* d, const void* x, const void* dy, float* dw, float* db,
                           crd_stream_t stream);

/* weight repacking between the reference's parameter layout and the kernels' K-major layout.
 * mode 0 (fwd):   dst[co][tap][map[ci]] = w[co][ci][tap]            dst is [Cout][KH*KW][Cin_p]
 * mode 1 (dgrad): dst[map[ci]][tap][co] = w[co][ci][tap]            dst is [Cin_p][KH*KW][Cout_p]
 * mode 2 (dgrad of an im2col GEMM): dst[tap][map[ci]][co] = w[co][ci][tap]   dst is [KH*KW*Cin_p][Cout_p]
 * map == NULL is the identity; dst must be zero-filled by the caller where unmapped. */
CRD_API int crd_weight_pack(const float* w, void* dst, int dst_dtype, const int* map, int Cout, int Cin, int taps,
                    int Cin_p, int Cout_p, int mode, crd_stream_t stream);
/* All packed copies in one launch (refresh after an optimizer step, runner.py:232).  `table` is a DEVICE array of
 * n_items + 1 rows of 12 int64: {w, dst, map, first block, Cout, Cin, taps, Cin_p, Cout_p, mode, dst dtype,
 * Cout*Cin*taps}; an item owns crd_weight_pack_blocks(Cout, Cin, taps) consecutive blocks (one per brick of output
 * channels x input channels x all taps; taps <= 128) and row n_items holds n_blocks; the rows are followed by n_blocks
 * int64 values, the item index of every block. */
CRD_API int crd_weight_pack_blocks(int Cout, int Cin, int taps);
/* dst[pix][0..C) = 0 for npix pixels of an NHWC buffer with pixel stride ld (dst points at the first channel to clear):
 * the zero padding / not-yet-written tail channels of the feature buffers ([features | depth | seg maps | pad],
 * CamRaDepth.py:137-162), without filling the whole buffer */
CRD_API int crd_zero_channels(void* dst, int ld, int dtype, int C, long long npix, crd_stream_t stream);
CRD_API int crd_weight_pack_batch(const long long* table, int n_items, int n_blocks, crd_stream_t stream);
/* Strided convolutions (patch embeddings k7s4/k3s2, spatial-reduction convs k=s; simplified_attention.py:68,
 * 158-160) run as GEMMs on the tensor-core path: gather the patches once, multiply, scatter the data gradient.
 * col is [B*Ho*Wo][KH*KW*Cin]; col2im is the gather-form adjoint. */
CRD_API int crd_im2col(const void* x, void* col, int dtype, int B, int H, int W, int Cin, int ldx, int Ho, int Wo,
               int KH, int KW, int stride, int pad, crd_stream_t stream);
CRD_API int crd_col2im(const void* dcol, void* dx, int dtype, int accumulate, int B, int H, int W, int Cin, int lddx,
               int Ho, int Wo, int KH, int KW, int stride, int pad, crd_stream_t stream);
/* grad[co][ci][tap] (+)= dwp[co][tap][map[ci]] */
CRD_API int crd_weight_unpack_grad(const float* dwp, float* grad, const int* map, int Cout, int Cin, int taps,
                           int Cin_p, int accumulate, crd_stream_t stream);
/* db[n] += sum_m dy[m][n] */
CRD_API int crd_col_sum(const void* dy, int dtype, float* db, long long M, int N, int ld, crd_stream_t stream);

/* ---------------------------------------------------------------- GroupNorm (+GELU, +Dropout2d)
 * nn.GroupNorm(C/16, C) everywhere: simplified_attention.py:23-24,70,117-118,162; utils.py:208-215.
 * Protocol: per-(b,c) sums -> finalize to a per-(b,c) affine (a,b) -> apply in the consumer. */
CRD_API int crd_chan_stats(const void* x, int dtype, float* sums /*[B][C][2], accumulated*/, int B, long long N,
                   int C, int ld, crd_stream_t stream);
CRD_API int crd_gn_finalize(const float* sums, const float* gamma, const float* beta, float* ab /*[B][C][2]*/,
                    float* mean_rstd /*[B][G][2]*/, float* xbar /*[B][C] or NULL: token mean of the output*/,
                    int B, int C, int G, long long N, float eps, crd_stream_t stream);
/* y = act(a*x+b) * post[b][c] */
CRD_API int crd_affine_act(const void* x, int in_dtype, void* y, int out_dtype, const float* ab, const float* post,
                   int act, int B, long long N, int C, int ldx, int ldy, crd_stream_t stream);
/* backward: dz = (dy + addbc[b][c]) * post * act'(a*x+b);  pq[b][c] += (sum dz, sum dz*x).
 * dz_out (optional, layout/dtype of dy, may alias dy) receives dz so that crd_gnact_bwd_apply can run with
 * act = NONE, post = addbc = NULL on it (the activation derivative is evaluated once, not twice). */
CRD_API int crd_gnact_bwd_reduce(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                         const float* post, const float* addbc, int act, float* pq, void* dz_out, int B,
                         long long N, int C, int lddy, int ldx, crd_stream_t stream);
/* coef[b][c][3] = (A, Bq, Cq) with dx = A*dz + Bq*x + Cq ; dgamma/dbeta[c] += ... */
CRD_API int crd_gn_bwd_finalize(const float* pq, const float* mean_rstd, const float* gamma, float* coef,
                        float* dgamma, float* dbeta, int B, int C, int G, long long N, crd_stream_t stream);
CRD_API int crd_gnact_bwd_apply(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab,
                        const float* post, const float* addbc, int act, const float* coef, void* dx,
                        int dx_dtype, int accumulate, int B, long long N, int C, int lddy, int ldx, int lddx,
                        crd_stream_t stream);

/* GroupNorm forward / backward in ONE launch each for tensors that fit the L2 (every GroupNorm of the encoder):
 * one CTA per (sample, channel tile of whole groups) sweeps its [N][tile] slab twice -- statistics, in-CTA finalize,
 * apply -- instead of the three launches of the protocol above.  Same arithmetic.
 *   fwd: y = act(a*x + b) * post.  sums_in (optional): per-(b,c) sums already produced by the conv read-out
 *        (crd_conv_fwd_tc gn_sums), then the statistics sweep is skipped.  y == NULL: finalize only.
 *        ab_out [B][C][2], mean_rstd_out [B][G][2], xbar_out [B][C] (each optional) are written for later use.
 *   bwd: as crd_gnact_bwd_reduce + crd_gn_bwd_finalize + crd_gnact_bwd_apply; with an activation dy is overwritten
 *        by dz (it is consumed here).  dx may not alias dy. */
CRD_API int crd_gn_fused_supported(int B, long long N, int C, int G);
CRD_API int crd_gn_fused_fwd(const void* x, int x_dtype, void* y, int y_dtype, const float* gamma, const float* beta,
                     const float* sums_in, const float* post, int act, float* ab_out, float* mean_rstd_out,
                     float* xbar_out, int B, long long N, int C, int G, int ldx, int ldy, float eps,
                     crd_stream_t stream);
CRD_API int crd_gn_fused_bwd(void* dy, int dy_dtype, const void* x, int x_dtype, const float* ab, const float* mean_rstd,
                     const float* gamma, const float* post, const float* addbc, int act, void* dx, int dx_dtype,
                     int accumulate, float* dgamma, float* dbeta, int B, long long N, int C, int G, int lddy, int ldx,
                     int lddx, crd_stream_t stream);

/* ---------------------------------------------------------------- encoder block pieces
 * DWConv (simplified_attention.py:313-323) fused with the preceding GroupNorm apply (Mlp.norm1, :36). */
CRD_API int crd_dwconv3x3_fwd(const void* x, int dtype, const float* ab, const float* w /*[C][9]*/, const float* bias,
                      void* y, int B, int H, int W, int C, crd_stream_t stream);
CRD_API int crd_dwconv3x3_bwd_input(const void* dy, int dtype, const float* w, void* dxn, int B, int H, int W, int C,
                            crd_stream_t stream);
/* fused backward: dxn (input gradient) and dw/db (accumulated) from ONE pass over dy */
CRD_API int crd_dwconv3x3_bwd(const void* dy, int dtype, const void* x, const float* ab, const float* w, void* dxn,
                      float* dw, float* db, int B, int H, int W, int C, crd_stream_t stream);
CRD_API int crd_dwconv3x3_bwd_weight(const void* dy, int dtype, const void* x, const float* ab, float* dw, float* db,
                             int B, int H, int W, int C, crd_stream_t stream);
/* Attention_MaxPool (simplified_attention.py:90-109), algebraically reduced:
 * s[b,n] = scale * sum_h max_m q_h[b,n,:].k_h[b,m,:] ; idx = argmax key per head */
CRD_API int crd_attn_qkmax_fwd(const void* q, const void* k, int dtype, float* s, unsigned short* idx, int B, int N,
                       int M, int C, int heads, float scale, crd_stream_t stream);
/* tcgen05 variant of the score (bf16 q/k, head_dim <= 64 and a multiple of 8; more than 256 keys run as chunks
 * of 256 with a running max): per head one
 * Q_h K_h^T GEMM into tensor memory, max/argmax taken in the accumulator read-out (simplified_attention.py:96-105).
 * Returns 1 (nothing launched) when the shape is not covered; crd_attn_qkmax_fwd dispatches to it. */
CRD_API int crd_attn_qkmax_fwd_tc(const void* q, const void* k, float* s, unsigned short* idx, int B, int N, int M,
                          int C, int heads, float scale, crd_stream_t stream);
CRD_API int crd_attn_qkmax_bwd(const float* ds, const void* q, const void* k, int dtype, const unsigned short* idx,
                       void* dq, float* dk /*[B][M][C] accumulated*/, int B, int N, int M, int C, int heads,
                       float scale, crd_stream_t stream);
/* pv[b][o] = sum_c Wp[o][c] * xbar[b][c]   (proj applied to the token-mean "v", :103,108) */
CRD_API int crd_attn_pv_fwd(const float* xbar, const float* Wp, float* pv, int B, int C, crd_stream_t stream);
/* dWp[o][c] += sum_b dpv[b][o]*xbar[b][c] ; dxbar[b][c] = dxbar_scale * sum_o Wp[o][c]*dpv[b][o] */
CRD_API int crd_attn_pv_bwd(const float* dpv, const float* xbar, const float* Wp, float* dWp, float* dxbar,
                    float dxbar_scale, int B, int C, crd_stream_t stream);
/* xout = x + dp[b] * (pv[b][c]*s[b][n] + bp[c])   (Block.forward :143) */
CRD_API int crd_attn_out_residual(const float* x, const float* pv, const float* s, const float* bp, const float* dp,
                          float* xout, int B, int N, int C, crd_stream_t stream);
/* given dx (grad of xout): ds[b][n] = dp*sum_c dx*pv ; dpv[b][c] += dp*sum_n dx*s ; dbp[c] += dp*sum dx */
CRD_API int crd_attn_out_bwd(const float* dx, const float* pv, const float* s, const float* dp, float* ds, float* dpv,
                     float* dbp, float* tmp /*[B][C][2] scratch*/, int B, int N, int C, crd_stream_t stream);
/* xout = x + dp[b]*y   (Block.forward :144) ; and its backward dy = dp[b]*dx */
CRD_API int crd_residual_add(const float* x, const void* y, int y_dtype, const float* dp, float* xout, int B,
                     long long N, int C, crd_stream_t stream);
CRD_API int crd_scale_cast(const float* dx, const float* dp, void* dy, int dy_dtype, int B, long long N, int C,
                   crd_stream_t stream);
/* generic elementwise accumulate: dst(f32) += src */
CRD_API int crd_add_f32(float* dst, const void* src, int src_dtype, long long n, crd_stream_t stream);

/* ---------------------------------------------------------------- decoder pieces
 * nn.Upsample(scale_factor=2, mode='bicubic') (utils.py:241,251), align_corners=False, A=-0.75. */
CRD_API int crd_bicubic2x_fwd(const void* x, void* y, int dtype, int B, int H, int W, int C, int ldx, int ldy,
                      crd_stream_t stream);
CRD_API int crd_bicubic2x_bwd(const void* dy, void* dx, int dtype, int accumulate, int B, int H, int W, int C, int lddy,
                      int lddx, crd_stream_t stream);
/* Depth_Activation.conv_2 (utils.py:283,288): 3x3, Cin -> 1, bias; y fp32 (B,1,H,W) */
CRD_API int crd_conv3x3_c1_fwd(const void* x, int dtype, const float* w /*[9][Cin]*/, const float* bias, float* y,
                       int B, int H, int W, int Cin, int ldx, crd_stream_t stream);
CRD_API int crd_conv3x3_c1_bwd(const float* dy, const void* x, int dtype, const float* w, void* dx, float* dw,
                       float* db, int B, int H, int W, int Cin, int ldx, int lddx, crd_stream_t stream);
/* same, with x = the OUTPUT of the sigmoid that feeds this conv (Depth_Activation, utils.py:285-289): dx receives the
 * gradient with respect to the sigmoid's INPUT, dx = dgrad * x * (1 - x) (one launch for the 32-channel bf16 case) */
CRD_API int crd_conv3x3_c1_bwd_sigmoid(const float* dy, const void* x, int dtype, const float* w, void* dx, float* dw,
                               float* db, int B, int H, int W, int Cin, int ldx, int lddx, crd_stream_t stream);
/* dx = dy * y * (1-y) (y = sigmoid output) */
CRD_API int crd_sigmoid_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, crd_stream_t stream);
/* Seg_Block (utils.py:95-100): map = argmax_c(logits)/ncls ; written to a channel of an NHWC buffer and/or
 * an fp32 (B,1,H,W) tensor */
CRD_API int crd_argmax_map(const void* logits, int dtype, int ld, int ncls, void* dst, int dst_dtype, int ld_dst,
                   float* dst_f32, long long npix, crd_stream_t stream);

/* ---------------------------------------------------------------- losses
 * MaskedSmoothL1Loss (loss_funcs.py:77-91), MaskedMSELoss (:36-46), MaskedFocalLoss (:14-34). */
CRD_API int crd_masked_l1_fwd(const float* pred, const float* target, float* acc /*[3]: sum smoothl1, count, sum sq*/,
                      long long n, crd_stream_t stream);
CRD_API int crd_masked_l1_bwd(const float* pred, const float* target, const float* acc, const float* gout, float* dpred,
                      long long n, crd_stream_t stream);
CRD_API int crd_ce_fwd(const float* logits /*NCHW*/, const long long* target, float* acc /*[2]: sum nll, count*/,
               int B, int C, long long HW, int ignore_index, crd_stream_t stream);
CRD_API int crd_ce_bwd(const float* logits, const long long* target, const float* acc, const float* gout, float gamma,
               float* dlogits, int B, int C, long long HW, int ignore_index, crd_stream_t stream);
/* out[0] = acc[0]/acc[1] ; (focal) out[0] = (1-exp(-ce))^gamma * ce ; (rmse) out[1] = sqrt(acc[2]/acc[1]) */
CRD_API int crd_loss_finalize(const float* acc, float* out, int kind, float gamma, crd_stream_t stream);

/* ---------------------------------------------------------------- test-mode metrics (runner.py:442-492)
 * pred clipped to [0,1], both scaled by max_depth, valid = 0 < gt <= thr1 (max_distances[0], runner.py:455-457);
 * second set additionally gt >= thr2 (the "<= 50 m" subset in inverse-depth space, :473-475).
 * acc: 8 floats scratch, out[6] = RMSE, MAE, REL x 2. */
CRD_API int crd_depth_metrics(const float* pred, const float* gt, float* acc, float* out, long long n,
                      float max_depth, float thr1, float thr2, crd_stream_t stream);
/* conf[t][p] += #pixels with label t (!= ignore_index) predicted as p = argmax_c logits (NCHW) -- IoU input */
CRD_API int crd_confusion(const float* logits, const long long* target, float* conf, int B, int C, long long HW,
                  int ignore_index, crd_stream_t stream);

/* ---------------------------------------------------------------- input pipeline (dataloader.py:202-257)
 * inverse-normalised lidar GT, its zero-ignoring 3x3/s2 min-pool pyramid, ImageNet normalisation of the
 * uint8 HWC camera image into channels [0,3) of the (B,Ctot,H,W) fp32 network input.
 * mean3 / std3 are HOST pointers to 3 floats. */
CRD_API int crd_gt_normalize(const float* d, float* g, long long n, float max_depth, crd_stream_t stream);
CRD_API int crd_minpool3x3s2(const float* x, float* y, int B, int H, int W, crd_stream_t stream);
CRD_API int crd_image_normalize(const unsigned char* img, float* out, int B, int H, int W, int Ctot,
                        const float* mean3, const float* std3, crd_stream_t stream);

/* segmentation ground truth resized with nearest-neighbour sampling (dataloader.py:262-268: skimage resize, order 0,
 * no anti-aliasing): dst[b][oh][ow] = src[b][floor((oh + .5) * Hi / Ho)][floor((ow + .5) * Wi / Wo)], int64 out;
 * src is uint8 (src_is_u8 = 1) or int64. */
CRD_API int crd_seg_resize_nearest(const void* src, int src_is_u8, long long* dst, int B, int Hi, int Wi, int Ho, int Wo,
                           crd_stream_t stream);
/* network input written directly in the engine's layout (NHWC bf16, ld channels per pixel, multiple of 8):
 * [ImageNet-normalised RGB from the uint8 HWC image | Ce <= 5 fp32 planes of extra (B,Ce,H,W) | zeros].
 * Removes the NCHW fp32 -> NHWC bf16 pack of the nn.Module boundary.  mean3 / std3 are HOST pointers. */
CRD_API int crd_pack_input_nhwc(const unsigned char* img, const float* extra, void* dst, int B, int H, int W, int Ce,
                        int ld, const float* mean3, const float* std3, crd_stream_t stream);
/* dst[pix][0..C) = src[pix][0..C) for every pixel: a (small, arbitrarily aligned) channel slice of one NHWC buffer
 * into a channel slice of another (the network input into the full-resolution concat buffers, CamRaDepth.py:146,163) */
CRD_API int crd_copy_channels(const void* src, int ld_src, void* dst, int ld_dst, int dtype, int C, long long npix,
                      crd_stream_t stream);
/* Stochastic masks of one training forward in ONE launch: timm DropPath scales (one per drop_path CALL: n_dp rows of
 * B values, keep probability keep_dp[row]; simplified_attention.py:143-144) followed by n_d2 Dropout2d scale planes
 * of B*C2 values (keep probability keep_d2; CamRaDepth.py:96).  Value = 1/keep with probability keep, else 0.
 * Philox4x32-10 keyed by state[0] (seed) and state[1] (step counter, advanced by the kernel: CUDA-graph replays
 * draw fresh masks).  state is DEVICE memory. */
CRD_API int crd_make_masks(float* out, const float* keep_dp, int n_dp, int B, int n_d2, int C2, float keep_d2,
                   unsigned long long* state, crd_stream_t stream);

/* ---------------------------------------------------------------- diffGradNorm (diffGradNorm.py:41-113)
 * multi-tensor: table rows describe (param, grad, exp_avg, exp_avg_sq, previous_grad, numel). */
typedef struct {
  float* p; const float* g; float* m; float* v; float* prev;
  long long numel;
} crd_opt_tensor;
typedef struct { int tensor; int pad; long long start; } crd_opt_chunk;   /* chunk of CRD_OPT_CHUNK elements */
#define CRD_OPT_CHUNK 16384
/* sumsq: nchunks floats, one partial sum of squares per chunk (no atomics; crd_diffgradnorm_update adds the partials
 * of a tensor in chunk order, so the update is deterministic and data-parallel replicas stay bit-identical) */
CRD_API int crd_mt_sumsq(const crd_opt_tensor* table, const crd_opt_chunk* chunks, int nchunks, float* sumsq,
                 crd_stream_t stream);
/* step_size[0] (DEVICE scalar, so a CUDA-graph replay sees fresh values) = lr*sqrt(1-beta2^t)/((1-beta1^t)+1e-8)
 * (diffGradNorm.py:108); egn_in/egn_out: per-tensor exp_grad_norm before/after (ping-pong so every chunk of a
 * tensor sees the same input) */
CRD_API int crd_diffgradnorm_update(const crd_opt_tensor* table, const crd_opt_chunk* chunks, int nchunks,
                            const float* sumsq, const float* egn_in, float* egn_out, const float* step_size,
                            float beta1, float beta2, float eps, crd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
