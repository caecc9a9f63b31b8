/*
 * glowk.h -- C ABI of the B200 (sm_100a) Glow flow kernels.
 *
 * Drop-in boundary for the hot path of corenel/pytorch-glow.
 * The reference has no FFI of its own: its boundary is the nn.Module surface of
 * network/module.py + network/model.py, all of which bottoms out in ATen calls.
 * Each entry point below replaces the ATen call sequence of the cited reference
 * lines; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - flow state is contiguous NCHW fp32 exactly as in the reference;
 *   - "rows" matrices are [P][ld] with P = N*H*W pixels (pixel-major, channel
 *     contiguous) and are the operands of the coupling-network GEMMs; their
 *     element type is selected by `act_dtype` (GLOWK_F32 | GLOWK_BF16);
 *   - `stream` is a cudaStream_t passed as void* (torch's current stream);
 *   - return value: 0 = ok, otherwise a GLOWK_E* code; glowk_last_error() gives
 *     the message for the calling thread.  No entry point keeps global mutable
 *     state, synchronises the device, or changes the current device.
 */
#ifndef GLOWK_H_
#define GLOWK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLOWK_OK 0
#define GLOWK_EINVAL 1   /* bad shape / argument (reference: Python assert) */
#define GLOWK_ECUDA 2    /* CUDA launch / runtime error */
#define GLOWK_EUNSUP 3   /* valid request this build cannot serve */

#define GLOWK_F32 0
#define GLOWK_BF16 1

/* GEMM epilogues (glowk_gemm) */
#define GLOWK_EPI_STORE 0        /* out = acc                               */
#define GLOWK_EPI_ACTNORM_RELU 1 /* out = relu((acc + bias[n]) * exp(f*logs[n])): Conv2d+ActNorm+ReLU, module.py:252-260,315 */
#define GLOWK_EPI_ACTNORM 2      /* same without the ReLU (stand-alone Conv2d module) */
#define GLOWK_EPI_ZEROS 3        /* out = (acc + bias[n]) * exp(f*logs[n]): Conv2dZeros, module.py:295-296 */
#define GLOWK_EPI_RELU_BWD 4     /* g = acc*[y>0]; out = g*exp(f*logs[n]); dlogs[n] += f*sum g*y; dbias[n] += exp(f*logs[n])*sum g */

const char* glowk_last_error(void);
int glowk_version(void);
/* 1 if the tcgen05/TMA (bf16) GEMM path is usable on the current device. */
int glowk_has_tcgen05(void);
/* Profiling aid: with GLOWK_GEMM_DEBUG=64 in the environment CTA 0 of the tcgen05 GEMM records how many cycles its
 * TMA / MMA / epilogue roles spent waiting on each other; this copies the 16 counters to a HOST array (it
 * synchronises the device; see gemm_sm100.cu for the meaning of each slot). */
int glowk_debug_gemm_trace(unsigned long long* out16_host);

/* ---- ActNorm: network/module.py:34-84,122-149 ------------------------------------------
 * fwd: y = (x + bias[c]) * exp(f*logs[c]);  rev: y = x * exp(-f*logs[c]) - bias[c].
 * x,y: [N,C,HW] (y may alias x).  The logdet term HW*sum(f*logs) is glowk_logdet_finish's. */
int glowk_actnorm(const float* x, float* y, const float* bias, const float* logs,
                  float logscale_factor, int64_t N, int64_t C, int64_t HW, int reverse, void* stream);

/* Data-dependent init, module.py:86-120: bias = -mean_{n,p} x;
 * logs = log(scale / (sqrt(mean (x+bias)^2) + 1e-6)) / f.  Strided so that both NCHW
 * (sN=C*HW, sC=HW, sP=1) and pixel-major rows (sN=HW*ld, sC=1, sP=ld) inputs work.
 * If `relu_after_prev` ... (not used).  Deterministic (one CTA per channel). */
int glowk_actnorm_init(const void* x, int act_dtype, int64_t N, int64_t C, int64_t HW,
                       int64_t sN, int64_t sC, int64_t sP, float scale, float logscale_factor,
                       float* bias_out, float* logs_out, void* stream);

/* The init variants the call above does not cover: batch_variance (module.py:112-113: ONE variance, the mean of
 * (x+bias)^2 over the whole tensor, for all channels) and the first call arriving in the reverse direction
 * (module.py:143-146 with 44-45, 62-63: logs from the raw second moment mean x^2, then bias = -mean(x * exp(-f*logs))). */
int glowk_actnorm_init_ex(const float* x, int64_t N, int64_t C, int64_t HW, int64_t sN, int64_t sC, int64_t sP,
                          float scale, float logscale_factor, int batch_variance, int reverse, float* bias_out,
                          float* logs_out, void* stream);

/* ---- Invertible 1x1 conv weight prep: module.py:356-357,365 (torch.det / .inverse) -------
 * LU with partial pivoting of the CxC matrix W (one CTA, fp64 internally):
 * logabsdet_out[0] = log|det W|; winv_out (nullable) = W^-1. */
int glowk_invconv_prepare(const float* w, int64_t C, float* logabsdet_out, float* winv_out, void* stream);
/* Same for `batch` matrices stored back to back (w: [batch][C][C]; logabsdet_out: [batch];
 * winv_out: [batch][C][C] or null): one launch, one CTA per matrix.  A FlowModel factorises all the
 * invconv weights of one channel count at once (FlowModel.encode's loop calls module.py:357 K*L times). */
int glowk_invconv_prepare_batched(const float* w, int64_t batch, int64_t C, float* logabsdet_out,
                                  float* winv_out, void* stream);

/* LU parameterisation (north_star; the reference raises NotImplementedError, module.py:336-337):
 * W = P . L . (U + diag(sign_s*exp(log_s))), L unit-lower (strict lower part of l), U strictly
 * upper part of u.  w_out = W, winv_out (nullable) = W^-1 by two triangular solves,
 * logabsdet_out[0] = sum(log_s). */
int glowk_invconv_lu_assemble(const float* p, const float* l, const float* u, const float* sign_s,
                              const float* log_s, int64_t C, float* w_out, float* winv_out,
                              float* logabsdet_out, void* stream);
/* Gradient chain of that parameterisation (autograd of W = P Lf Uf): from dW ([C][C], the gradient wrt the assembled
 * weight, logdet term included) accumulate dl += tril(P^T dW Uf^T, -1), du += triu(Lf^T P^T dW, 1),
 * dlog_s += diag(Lf^T P^T dW) * sign_s * exp(log_s).  One small kernel per FlowStep. */
int glowk_invconv_lu_grads(const float* dw, const float* p, const float* l, const float* u, const float* sign_s,
                           const float* log_s, int64_t C, float* dl, float* du, float* dlog_s, void* stream);

/* ---- Fused ActNorm + channel mix: model.py:94-103 (fwd) / 142-152 (rev) ------------------
 * fwd: z[n,o,p] = sum_i W[o,i] * ((x[n,i,p] + bias[i]) * exp(f*logs[i]))          (mix)
 *      z[n,o,p] =               (x[n,idx[o],p] + bias[idx[o]]) * exp(f*logs[idx[o]]) (perm, bit-exact gather)
 * rev: x[n,o,p] = (sum_i Winv[o,i] z[n,i,p]) * exp(-f*logs[o]) - bias[o]           (w = W^-1 / idx = inverse idx)
 * Exactly one of w / idx is non-null.  bias/logs null => plain Invertible1x1Conv / Permutation2d. */
int glowk_actnorm_mix(const float* x, float* z, const float* w, const int64_t* idx,
                      const float* bias, const float* logs, float logscale_factor,
                      int64_t N, int64_t C, int64_t HW, int reverse, void* stream);

/* ---- Squeeze2d: module.py:551-591 (bit-exact index map) ---------------------------------
 * fwd: y[n, c*f*f + fh*f + fw, i, j] = x[n, c, i*f+fh, j*f+fw];  x: [N,C,H,W]. reverse = unsqueeze
 * with x: [N,C,H,W] -> y: [N, C/f^2, H*f, W*f].  sN = batch stride of x in elements (C*H*W when
 * contiguous; larger for the channel-sliced view Split2d returns, module.py:111). */
int glowk_squeeze2d(const float* x, float* y, int64_t N, int64_t C, int64_t H, int64_t W,
                    int64_t sN, int factor, int reverse, void* stream);

/* ---- rows (pixel-major) <-> NCHW ---------------------------------------------------------
 * im2col for a SAME-padded kxk conv (k = 1 or 3), module.py:209-212,252:
 * dst[p][tap*Cin + ci] = src[n, c0+ci, y+ky-pad, x+kx-pad] (0 outside), columns >= k*k*Cin zeroed
 * up to ld.  sN = batch stride of src in elements (>= (c0+Cin)*H*W).  flip=1 mirrors the taps (transposed conv, used by the backward pass). */
int glowk_im2col(const float* src, int64_t N, int64_t sN, int64_t c0, int64_t Cin, int64_t H, int64_t W,
                 int ksize, int flip, void* dst, int act_dtype, int64_t ld, void* stream);
/* Same gather from a pixel-major fp32 source [P][ld_src] (channels c0..c0+Cin). */
int glowk_im2col_rows(const float* src, int64_t ld_src, int64_t N, int64_t c0, int64_t Cin, int64_t H,
                      int64_t W, int ksize, int flip, void* dst, int act_dtype, int64_t ld, void* stream);
/* glowk_im2col_rows (3x3, bf16 output, even Cin / c0 / ld_src) that additionally writes 1.0 into the zero-padding
 * column ones_col (9*Cin <= ones_col < ld; -1 = none).  The conv (zero weight in that column) and its dgrad are
 * unaffected; the weight-gradient GEMM dW = d^T a1 then carries sum_p d[p][n] -- the bias gradient of the ActNorm
 * behind the conv (module.py:34-50) -- in column ones_col, so the dgrad epilogue needs no column sum
 * (see glowk_conv_actnorm_finish_batched). */
int glowk_im2col_rows_ones(const float* src, int64_t ld_src, int64_t N, int64_t c0, int64_t Cin, int64_t H,
                           int64_t W, int ksize, int flip, void* dst, int act_dtype, int64_t ld, int64_t ones_col,
                           void* stream);
/* dst[n,c,p] = rows[(n*HW+p)*ld + c] for c < C (rows of type act_dtype). */
int glowk_rows_to_nchw(const void* rows, int act_dtype, int64_t ld, float* dst, int64_t N, int64_t C,
                       int64_t HW, void* stream);
/* 3x3 tap gather-sum: dst[n,c0+c,y,x] (+)= sum_tap P[pix(n,y+ky-1,x+kx-1)][tap*C + c], the
 * second half of a conv computed as nine pointwise GEMMs.  flip=1 mirrors taps. */
int glowk_tapsum_to_nchw(const float* P, int64_t ldp, float* dst, int64_t N, int64_t Ctot, int64_t c0,
                         int64_t C, int64_t H, int64_t W, int flip, int accumulate, void* stream);

/* Conv weight packing, fp32 [O][I][k][k] (module.py nn.Conv2d layout) -> GEMM B operands:
 *   layout 0: dst[o][tap*I + i]      ([O][ld], ld >= k*k*I)   forward, taps folded into K
 *   layout 1: dst[tap*O + o][i]      ([rows][ld], ld >= I)    forward, taps folded into N
 *   layout 2: dst[tap*I + i][o]      transpose of layout 0    (dgrad)
 *   layout 3: dst[i][tap*O + o]      transpose of layout 1    (dgrad)
 * Padding rows/cols are zeroed: dst has `rows` rows of `ld` elements. */
int glowk_pack_conv_weight(const float* w, int64_t O, int64_t I, int ksize, int layout,
                           void* dst, int act_dtype, int64_t rows, int64_t ld, void* stream);

/* ---- GEMM: out[M][N] = epilogue(A[M][K] . B[N][K]^T) --------------------------------------
 * The coupling network's convs (module.py:300-319) as pixel-major GEMMs.  act_dtype GLOWK_BF16
 * runs the tcgen05/TMEM/TMA kernel (bf16 operands, fp32 accumulate); GLOWK_F32 runs the fp32
 * CUDA-core kernel used for strict-parity runs.  A: [M][lda], B: [N][ldb] both K-contiguous.
 * out_dtype selects the element type of out ([M][ldo]).  bias/logs: fp32 [N] epilogue vectors.
 * EPI_RELU_BWD additionally reads y ([M][ldy], act_dtype) and accumulates dlogs/dbias (fp32 [N]). */
int glowk_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype,
               int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const float* logs,
               float logscale_factor, const void* y, int64_t ldy, float* dlogs, float* dbias,
               void* out, int out_dtype, int64_t ldo, void* stream);
/* Same with an explicit thread-block-cluster shape for the tcgen05 path: cluster_m m-tiles share (TMA-multicast)
 * the B tile and cluster_n n-tiles share the A tile; powers of two, cluster_m*cluster_n <= 8; 0 = library default. */
int glowk_gemm_ex(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype,
                  int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const float* logs,
                  float logscale_factor, const void* y, int64_t ldy, float* dlogs, float* dbias,
                  void* out, int out_dtype, int64_t ldo, int cluster_m, int cluster_n, void* stream);

/* Weight-gradient GEMM: dW[Mo][No] (+)= A[P][Mo]^T . B[P][No]  (reduction over pixels, fp32 out,
 * split over CTAs with atomic accumulation).  A: [P][lda], B: [P][ldb] of act_dtype. */
int glowk_gemm_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, int act_dtype,
                     int64_t P, int64_t Mo, int64_t No, float* dW, int64_t lddw, void* stream);

/* ---- Fused coupling network f() (module.py:300-319): conv1 3x3 -> ActNorm -> ReLU -> conv2 1x1 -> ActNorm ->
 * ReLU -> conv3 3x3 as ONE tcgen05 kernel per 128-pixel tile (hidden = 512, bf16 operands, fp32 accumulate).
 * The 512-wide hidden activations never leave the SM: h1 is written back to tensor memory (tcgen05.st) and is
 * the TMEM A operand of conv2, h2 goes through a 16 KB shared-memory chunk straight into conv3's MMAs.
 *   a1 : [M][lda]   bf16  im2col rows of z1 (glowk_im2col_rows), K1 = padded 9*Cin (multiple of 64, <= 256)
 *   w1 : [512][ldw1], w2: [512][ldw2], w3: [N3][ldw3]   bf16 GEMM-layout weights (glowk_pack_conv_weight
 *        layouts 0, 0, 1); N3 = padded 9*Cout (multiple of 16, <= 128; or a multiple of 32 <= 256)
 *   bias*, logs*, f* : the two hidden ActNorms (module.py:238-239, 34-84)
 *   p3 : [M][ldp3]  fp32  conv3 in tap form (input of glowk_rows_coupling)
 *   h1_save / h2_save : NULL when sampling; [M][ldh] bf16 to keep the activations for glowk_cnet_backward
 * Results are bit-identical to three glowk_gemm calls (EPI_ACTNORM_RELU, EPI_ACTNORM_RELU, EPI_STORE).
 * glowk_cnet_fused_supported(backward, K1, hidden, N3) tells whether a shape is served (else: glowk_gemm). */
/* Profiling aid: with GLOWK_CNET_DEBUG=1 CTA 0 of the fused kernels records the cycles its roles spent waiting on
 * each other (layout: csrc/cnet_fused_sm100.cu); this copies the 16 counters to a HOST array (synchronises). */
int glowk_debug_cnet_trace(unsigned long long* out16_host);
/* Same for the event timeline of one tile (GLOWK_CNET_DEBUG bit 16; 64 SM-clock stamps). */
int glowk_debug_cnet_timeline(unsigned long long* out64_host);
int glowk_cnet_fused_supported(int backward, int64_t K1, int64_t hidden, int64_t N3);
int glowk_cnet_forward(const void* a1, int64_t lda, const void* w1, int64_t ldw1, const void* w2, int64_t ldw2,
                       const void* w3, int64_t ldw3, int64_t M, int64_t K1, int64_t hidden, int64_t N3,
                       const float* bias1, const float* logs1, float f1, const float* bias2, const float* logs2,
                       float f2, float* p3, int64_t ldp3, void* h1_save, void* h2_save, int64_t ldh, void* stream);
/* Same with conv1 as an IMPLICIT GEMM (module.py:252 F.conv2d 3x3, SAME zero padding): the kernel gathers the
 * im2col tile of channels c0 .. c0+Cin-1 of the pixel-major flow state z ([N*H*W][ld_z] fp32) itself -- exactly what
 * glowk_im2col_rows would have written (k = tap*Cin + ci, bf16 rounding, zero padding columns, ones_col >= 9*Cin set
 * to 1.0 or -1) -- so neither that kernel nor its output exist when sampling.  a1_save (nullable, [M][lda] bf16)
 * receives the tile for the weight gradient of conv1 when training. */
int glowk_cnet_forward_implicit(const float* z, int64_t ld_z, int64_t c0, int64_t Cin, int64_t N, int64_t H, int64_t W,
                                int64_t ones_col, void* a1_save, int64_t lda, const void* w1, int64_t ldw1,
                                const void* w2, int64_t ldw2, const void* w3, int64_t ldw3, int64_t K1, int64_t hidden,
                                int64_t N3, const float* bias1, const float* logs1, float f1, const float* bias2,
                                const float* logs2, float f2, float* p3, int64_t ldp3, void* h1_save, void* h2_save,
                                int64_t ldh, void* stream);
/* Adjoint chain of the same network (autograd of module.py:300-319) in one kernel:
 *   d2 = [h2 > 0] * (d3col . w3t^T) * exp(f2*logs2)   -> TMEM A operand + stored [M][ldh] bf16 (wgrad operand)
 *   d1 = [h1 > 0] * (d2 . w2t^T) * exp(f1*logs1)      -> shared-memory chunk + stored [M][ldh] bf16
 *   da1 = d1 . w1t^T                                  -> [M][ldda1] bf16 (conv1 dgrad in im2col form)
 * d3col: [M][ldd3] bf16 flipped im2col of du, K3 = padded 9*Cout (multiple of 64, <= 256); w3t: [512][ldw3t],
 * w2t: [512][ldw2t], w1t: [K1p][ldw1t] (glowk_pack_conv_weight layouts 3, 2, 2); K1p multiple of 16, <= 128.
 * dbias2 / dbias1 (nullable, fp32 [512]): += column sums of the stored d2 / d1 (ActNorm bias gradients). */
int glowk_cnet_backward(const void* d3col, int64_t ldd3, const void* w3t, int64_t ldw3t, const void* w2t,
                        int64_t ldw2t, const void* w1t, int64_t ldw1t, int64_t M, int64_t K3, int64_t hidden,
                        int64_t K1p, const float* logs2, float f2, const float* logs1, float f1, const void* h2,
                        const void* h1, void* d2, void* d1, int64_t ldh, void* da1, int64_t ldda1, float* dbias2,
                        float* dbias1, void* stream);

/* glowk_cnet_backward with dgrad3's operand gathered in-kernel: d3col = flipped 3x3 im2col of du ([N*H*W][ldu] fp32,
 * the gradient of Conv2dZeros' output from glowk_rows_coupling_bwd) -- exactly glowk_im2col_rows(flip = 1).  The tile is
 * also stored to d3col_save ([M][ldd3] bf16): the conv3 weight-gradient GEMM reads it. */
int glowk_cnet_backward_implicit(const float* du, int64_t ldu, int64_t Cout, int64_t N, int64_t H, int64_t W,
                                 void* d3col_save, int64_t ldd3, const void* w3t, int64_t ldw3t, const void* w2t,
                                 int64_t ldw2t, const void* w1t, int64_t ldw1t, int64_t K3, int64_t hidden, int64_t K1p,
                                 const float* logs2, float f2, const float* logs1, float f1, const void* h2,
                                 const void* h1, void* d2, void* d1, int64_t ldh, void* da1, int64_t ldda1,
                                 float* dbias2, float* dbias1, void* stream);

/* The same two calls with the ReLU masks of the hidden activations as BITS.  The training forward writes, next to
 * h1 / h2 (still needed by the weight-gradient GEMMs), one bit per element; the backward chain then reads 1/16 of the
 * bytes for its two ReLU' masks (the kernel is bound by its HBM traffic, of which the bf16 masks were 45 %).
 * mask1 / mask2: glowk_cnet_relu_mask_bytes(M) bytes each, 8-byte aligned, opaque layout
 * ([tile of 128 rows][8 column groups of 64][128 rows] x 64 bits).  Results are bit-identical to the unmasked calls. */
int64_t glowk_cnet_relu_mask_bytes(int64_t M);
int glowk_cnet_forward_masked(const void* a1, int64_t lda, const void* w1, int64_t ldw1, const void* w2, int64_t ldw2,
                              const void* w3, int64_t ldw3, int64_t M, int64_t K1, int64_t hidden, int64_t N3,
                              const float* bias1, const float* logs1, float f1, const float* bias2, const float* logs2,
                              float f2, float* p3, int64_t ldp3, void* h1_save, void* h2_save, int64_t ldh, void* mask1,
                              void* mask2, void* stream);
int glowk_cnet_forward_implicit_masked(const float* z, int64_t ld_z, int64_t c0, int64_t Cin, int64_t N, int64_t H,
                                       int64_t W, int64_t ones_col, void* a1_save, int64_t lda, const void* w1,
                                       int64_t ldw1, const void* w2, int64_t ldw2, const void* w3, int64_t ldw3,
                                       int64_t K1, int64_t hidden, int64_t N3, const float* bias1, const float* logs1,
                                       float f1, const float* bias2, const float* logs2, float f2, float* p3,
                                       int64_t ldp3, void* h1_save, void* h2_save, int64_t ldh, void* mask1, void* mask2,
                                       void* stream);
int glowk_cnet_backward_implicit_masked(const float* du, int64_t ldu, int64_t Cout, int64_t N, int64_t H, int64_t W,
                                        void* d3col_save, int64_t ldd3, const void* w3t, int64_t ldw3t, const void* w2t,
                                        int64_t ldw2t, const void* w1t, int64_t ldw1t, int64_t K3, int64_t hidden,
                                        int64_t K1p, const float* logs2, float f2, const float* logs1, float f1,
                                        const void* h2, const void* h1, void* d2, void* d1, int64_t ldh, void* da1,
                                        int64_t ldda1, float* dbias2, float* dbias1, const void* mask2,
                                        const void* mask1, void* stream);

/* ---- Coupling: model.py:105-115 (fwd) / 131-140 (rev) --------------------------------------
 * h[n,co,y,x] = (u + bias3[co]) * exp(f*logs3[co]),  u = 3x3 tap gather-sum of P (ldp floats/row,
 * column tap*Cout+co), i.e. Conv2dZeros (module.py:295-296).
 * affine (Cout = C):  shift = h[2j], scale = sigmoid(h[2j+1] + 2);
 *    fwd z2 = (z2 + shift)*scale, partial += sum log scale ; rev z2 = z2/scale - shift, partial -= ...
 * additive (Cout = C/2): fwd z2 += h[j]; rev z2 -= h[j].
 * z: [N,C,H,W], channels C/2.. are updated IN PLACE.  partials: [N][nblk] per-CTA sums of
 * log(scale) for sample n (nblk = glowk_coupling_nblk(H*W)); summed by glowk_logdet_finish.
 * h_save (nullable): [P][Cout] fp32 copy of h for the backward pass. */
int64_t glowk_coupling_nblk(int64_t HW);
int glowk_coupling(const float* P, int64_t ldp, const float* bias3, const float* logs3,
                   float logscale_factor, float* z, float* partials, float* h_save,
                   int64_t N, int64_t C, int64_t H, int64_t W, int affine, int reverse, void* stream);

/* logdet_out[n] = logdet_in[n] + sign * ( HW * (sum_c f*logs[c] + logabsdet[0]) ) + sum_b partials[n][b]
 * (module.py:77-82, 357-367; model.py:114,140).  logs/logabsdet/partials may each be null.
 * `sign` = +1 forward, -1 reverse (partials already carry their own sign). */
int glowk_logdet_finish(const float* logdet_in, float* logdet_out, const float* logs, int64_t C,
                        float logscale_factor, const float* logabsdet, const float* partials,
                        int64_t nblk, int64_t HW, float sign, int64_t N, void* stream);

/* ---- Split2d / GaussianDiag: module.py:437-483, 511-536 -----------------------------------
 * h: [P][ldh] fp32 rows holding Conv2dZeros(z1) (mean = h[2j], logs = h[2j+1], 'cross' split);
 * x: [N,C,H,W]; z2 = x[:, C/2:].  out[n] = logdet_in[n] + sum_{j,p} -0.5(log2pi + 2 logs + (z2-mean)^2/exp(2 logs)).
 * h == null => mean = logs = 0 over all C channels of x (the Glow top prior, model.py:435-438). */
int glowk_gaussian_logp(const float* h, int64_t ldh, const float* x, int64_t N, int64_t C, int64_t HW,
                        int64_t c0, int64_t Cz, const float* logdet_in, float* logdet_out, void* stream);
/* Loss head (network/model.py:425-427, 435-450, 496-498; plain N(0, I) top prior): z [N][D] is the top latent, ld [N]
 * the flow's accumulated logdet started from ZERO (null = 0); per sample
 *   objective = ld[n] + c0 + sum log N(z[n]; 0, I)     (c0 = -log(n_bins) * D_x: the objective's initial value)
 *   nll[n]    = -objective / denom                     (denom = ln 2 * D_x: bits per dimension)
 * and, if loss != null, loss[0] = mean_n nll[n] (Glow.generative_loss), summed in index order by the last CTA;
 * ticket: one zero-initialised uint32 that the kernel leaves at zero. */
int glowk_nll_head(const float* z, const float* ld, float c0, float denom, int64_t N, int64_t D, float* nll,
                   float* loss, void* ticket, void* stream);
/* Adjoint of glowk_nll_head.  g_loss = dL/dloss (device scalar), g_nll = dL/dnll [N], dz_in = dL/dz [N][D]; each may be
 * null (both g's null: the loss gradient is 1).  coef_n = (g_loss / N + g_nll[n]) / denom;
 * dz = dz_in + z * coef_n, dld[n] = -coef_n (the gradient of the flow's logdet output). */
int glowk_nll_head_bwd(const float* z, const float* g_loss, const float* g_nll, const float* dz_in, float denom,
                       int64_t N, int64_t D, float* dz, float* dld, void* stream);
/* Reverse: out[:, :C/2] = z1, out[:, C/2:] = mean + exp(logs) * eps   (eps already scaled by eps_std). */
int glowk_split2d_sample(const float* h, int64_t ldh, const float* z1, const float* eps, float* out,
                         int64_t N, int64_t Chalf, int64_t HW, void* stream);

/* ================================ backward (training) =====================================
 * The reference trains through torch autograd over the ATen ops of the lines cited above
 * (network/trainer.py:138-140 loss.backward()).  These are the analytic adjoints of the forward
 * kernels; parameter gradients are ACCUMULATED (atomically) into the given fp32 buffers. */

/* Adjoint of glowk_coupling (+ the Conv2dZeros scale): y = step output, hrows = h saved by the forward
 * ([P][Cout]), dy = grad wrt y, dld = grad wrt this step's per-sample logdet ([N], nullable).
 * Writes dz [N,C,H,W] (dz1 = dy1; the conv dgrad is added by glowk_tapsum_to_nchw) and du rows
 * [P][Cout] = grad wrt the tap-summed conv output; accumulates dlogs3/dbias3 [Cout]. */
int glowk_coupling_bwd(const float* y, const float* hrows, const float* dy, const float* dld,
                       const float* logs3, float logscale_factor, float* dz, float* du, float* dlogs3,
                       float* dbias3, int64_t N, int64_t C, int64_t H, int64_t W, int affine, void* stream);

/* Adjoint of Split2d forward (module.py:526-530): dz1 = grad wrt the returned z1 ([N,C/2,H,W], batch
 * stride dz1_sN; nullable), dld = grad wrt logdet [N].  Writes dx [N,C,H,W] and du rows [P][ldu]. */
int glowk_split2d_bwd(const float* x, const float* hrows, int64_t ldh, const float* dz1, int64_t dz1_sN,
                      const float* dld, const float* logs_p, float logscale_factor, float* dx, float* du,
                      int64_t ldu, float* dlogs_p, float* dbias_p, int64_t N, int64_t C, int64_t HW, void* stream);

/* Adjoint of glowk_actnorm_mix (forward direction): dx = s*(W^T dz); dw += sum_p dz a^T;
 * dlogs += f*sum da*a; dbias += sum da*s, with a = (x+bias)*s recomputed from the step input x. */
int glowk_actnorm_mix_bwd(const float* x, const float* dz, const float* w, const int64_t* idx,
                          const float* bias, const float* logs, float logscale_factor, float* dx, float* dw,
                          float* dlogs, float* dbias, int64_t N, int64_t C, int64_t HW, void* stream);

/* Gradient of the sample-independent logdet terms: dlogs[c] += f*HW*G, dw += HW*G*W^-T, G = sum_n dld[n]. */
int glowk_logdet_param_grad(const float* dld, int64_t N, int64_t HW, float logscale_factor, float* dlogs,
                            int64_t C, const float* winv, float* dw, void* stream);

/* Inverse of glowk_pack_conv_weight for gradients: grad[O][I][k][k] (+)= src (packed layout, fp32). */
int glowk_unpack_weight_grad(const float* src, int64_t ld, int64_t O, int64_t I, int ksize, int layout,
                             float* grad, int accumulate, void* stream);

/* ---- Fused optimizer step over a flat fp32 arena: trainer.py:142-150 + torch.optim.Adam (builder.py:10-13)
 * clip_norm: grads = clamp(grads, +-clip_value) in place (clip_grad_value_), then workspace[0] = global
 * L2 norm, workspace[1] = min(1, max_norm/(norm+1e-6)) (clip_grad_norm_).  workspace: device floats,
 * glowk_optim_workspace_floats() long.
 * adam: grads *= norm_coef[1]; m,v,params updated (no weight decay, no amsgrad).  sched_dev (nullable):
 * device [lr, 1-beta1^step, sqrt(1-beta2^step)] overriding lr/step so a CUDA graph can be replayed. */
int64_t glowk_optim_workspace_floats(void);
int glowk_optim_clip_norm(float* grads, int64_t n, float clip_value, float max_norm, float* workspace, void* stream);
int glowk_optim_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     const float* norm_coef, const float* sched_dev, float lr, float beta1, float beta2,
                     float eps, int64_t step, void* stream);
/* torch.optim.Adamax (network/builder.py:10-13, the reference's `optimizer: "adamax"` choice) over the same arenas:
 * grads *= norm_coef[1]; m = beta1*m + (1-beta1)*g; u = max(beta2*u, |g| + eps); p -= lr/(1-beta1^step) * m/u. */
int glowk_optim_adamax(float* params, float* grads, float* exp_avg, float* exp_inf, int64_t n, const float* norm_coef,
                       const float* sched_dev, float lr, float beta1, float beta2, float eps, int64_t step, void* stream);
/* Fills sched_dev = [noam_lr(step), 1-beta1^(step+1), sqrt(1-beta2^(step+1))] from the DEVICE counter *step_dev
 * (int64, completed iterations) and increments it: misc/lr_scheduler.py:18-37 (noam_decay; warmup_steps = 0 gives
 * the constant base_lr, min_lr < 0 = none) + the bias corrections of torch.optim.Adam, inside a captured graph. */
int glowk_optim_schedule(void* step_dev, float* sched_dev, float base_lr, int64_t warmup_steps, float min_lr,
                         float beta1, float beta2, void* stream);

/* =========================== pixel-major ("rows") flow state ================================
 * FlowModel.encode / decode (network/model.py:263-294) keep the flow state between the NCHW tensors of
 * the reference API as rows x[p][c] (p = (n*H + y)*W + x, c contiguous, fp32) -- the layout of every
 * coupling-network GEMM operand -- so each kernel below reads and writes whole contiguous pixels.
 * Element arithmetic is identical to the NCHW entry points above; only reduction orders differ.
 * All of them need C %% 4 == 0 and C <= glowk_rows_max_channels(). */
int glowk_rows_max_channels(void);

/* glowk_actnorm_mix on rows: x, z: [P][C] (model.py:94-103 fwd / 142-152 rev). */
int glowk_rows_actnorm_mix(const float* x, float* z, const float* w, const int64_t* idx, const float* bias,
                           const float* logs, float logscale_factor, int64_t P, int64_t C, int reverse, void* stream);

/* glowk_coupling + glowk_logdet_finish on rows (model.py:105-115 / 131-140; module.py:77-82,357-367):
 * z: [P][C], channels C/2.. updated in place; h_save (nullable): [P][Cout].  If ld_out is non-null:
 * ld_out[n] = ld_in[n] + sign*HW*(sum_c an_logscale_factor*an_logs[c] + logabsdet[0]) + sum log(scale)
 * (an_logs / logabsdet / ld_in nullable), reduced deterministically through `partials`
 * ([N][glowk_rows_coupling_nblk(HW, C)] floats) by the last CTA of each sample; `tickets` is N uint32
 * that must be zero on entry and are left zero. */
int64_t glowk_rows_coupling_nblk(int64_t HW, int64_t C);
int glowk_rows_coupling(const float* P3, int64_t ldp, const float* bias3, const float* logs3, float logscale_factor,
                        float* z, float* h_save, int64_t N, int64_t C, int64_t H, int64_t W, int affine, int reverse,
                        const float* ld_in, float* ld_out, const float* an_logs, float an_logscale_factor,
                        const float* logabsdet, float sign, float* partials, void* tickets, void* stream);

/* Reverse pass of a FlowStep behind the coupling network in one launch (model.py:131-152): inverse coupling
 * (glowk_rows_coupling, reverse = 1, no logdet) applied to the rows x while they are staged in shared memory, then
 * glowk_rows_actnorm_mix(reverse = 1) with w = W^-1 (or idx = inverse permutation); x is NOT modified, the result
 * goes to z.  Bit-identical to the two calls; C a multiple of 4, <= glowk_rows_max_channels(). */
int glowk_rows_coupling_rev_mix(const float* P3, int64_t ldp, const float* bias3, const float* logs3,
                                float logscale_factor3, const float* x, float* z, const float* w, const int64_t* idx,
                                const float* bias, const float* logs, float logscale_factor, int64_t N, int64_t C,
                                int64_t H, int64_t W, int affine, void* stream);

/* glowk_coupling_bwd on rows: y, dy, dz: [P][C]; hrows, du: [P][Cout]. */
int glowk_rows_coupling_bwd(const float* y, const float* hrows, const float* dy, const float* dld, const float* logs3,
                            float logscale_factor, float* dz, float* du, float* dlogs3, float* dbias3, int64_t N,
                            int64_t C, int64_t HW, int affine, void* stream);

/* glowk_actnorm_mix_bwd on rows, with the conv1 dgrad folded into the load of dz:
 * dz[p][c] += sum_tap dA1[nbr(p, 8-tap)][tap*Cin + c] for c < Cin (dA1: [P][ld_a1] fp32, the dgrad GEMM of the
 * coupling net's first conv in im2col form; nullable) -- i.e. glowk_tapsum_to_nchw(flip=1, accumulate=1) --
 * and, if dld ([N], nullable) is given, glowk_logdet_param_grad folded in (winv = W^-1, needed with w). */
int glowk_rows_actnorm_mix_bwd(const float* x, const float* dz, const float* dA1, int64_t ld_a1, int64_t Cin,
                               const float* w, const int64_t* idx, const float* bias, const float* logs,
                               float logscale_factor, float* dx, float* dw, float* dlogs, float* dbias, int64_t N,
                               int64_t C, int64_t H, int64_t W, const float* dld, const float* winv, void* stream);

/* Same with the element type of dA1 given (GLOWK_F32 or GLOWK_BF16): the bf16 training path lets the conv1 dgrad
 * GEMM store dA1 in bf16 (half the bytes of that GEMM's output and of this kernel's gather; the nine taps are
 * still summed in fp32).  A bf16 dA1 needs 16-byte aligned rows. */
int glowk_rows_actnorm_mix_bwd_ex(const float* x, const float* dz, const void* dA1, int da1_dtype, int64_t ld_a1,
                                  int64_t Cin, const float* w, const int64_t* idx, const float* bias, const float* logs,
                                  float logscale_factor, float* dx, float* dw, float* dlogs, float* dbias, int64_t N,
                                  int64_t C, int64_t H, int64_t W, const float* dld, const float* winv, void* stream);

/* Largest channel count of the rows coupling / Split2d kernels (384: levels 5-6 of the 256x256 L=6 model).  Above
 * glowk_rows_max_channels() (96) the ActNorm + 1x1 conv pair of a FlowStep runs as glowk_actnorm on the rows
 * (HW = 1) + glowk_gemm (fp32), and its adjoint as glowk_rows_tapsum + glowk_gemm / glowk_gemm_wgrad +
 * glowk_rows_actnorm_bwd + glowk_logdet_param_grad. */
int glowk_rows_max_channels_wide(void);
/* ActNorm adjoint on rows (module.py:34-84): dx = da*s, dbias += sum_p da*s, dlogs += f*sum_p da*(x+bias)*s with
 * s = exp(f*logs); da, x, dx: [P][C] (dx may alias da). */
int glowk_rows_actnorm_bwd(const float* da, const float* x, const float* bias, const float* logs, float logscale_factor,
                           float* dx, float* dlogs, float* dbias, int64_t P, int64_t C, void* stream);

/* glowk_gaussian_logp on rows: x: [P][ldx], channels c0..c0+Cz; h: [P][ldh] or null (N(0,I)). */
int glowk_rows_gaussian_logp(const float* h, int64_t ldh, const float* x, int64_t ldx, int64_t N, int64_t HW,
                             int64_t c0, int64_t Cz, const float* logdet_in, float* logdet_out, void* stream);
/* glowk_split2d_sample on rows: z1: [P][ldz1] (first Chalf channels), out: [P][2*Chalf]; eps stays NCHW
 * [N,Chalf,H,W] (drawn by torch's generator in the reference's element order, module.py:419-421). */
int glowk_rows_split2d_sample(const float* h, int64_t ldh, const float* z1, int64_t ldz1, const float* eps, float* out,
                              int64_t N, int64_t Chalf, int64_t HW, void* stream);
/* glowk_split2d_bwd on rows: writes channels C/2.. of dx ([P][C]) and du ([P][ldu]); channels 0..C/2-1 of dx
 * (the gradient of the returned z1) are written by glowk_rows_squeeze beforehand. */
int glowk_rows_split2d_bwd(const float* x, const float* hrows, int64_t ldh, const float* dld, const float* logs_p,
                           float logscale_factor, float* dx, float* du, int64_t ldu, float* dlogs_p, float* dbias_p,
                           int64_t N, int64_t C, int64_t HW, void* stream);

/* glowk_tapsum_to_nchw on rows: dst[p][c0 + c] (+)= sum_tap P[nbr(p, tap)][tap*C + c]; dst: [P][ld_dst]. */
int glowk_rows_tapsum(const float* P, int64_t ldp, float* dst, int64_t ld_dst, int64_t c0, int64_t C, int64_t N,
                      int64_t H, int64_t W, int flip, int accumulate, void* stream);

/* Squeeze2d / unsqueeze (module.py:551-591, bit-exact) between layouts.  `full` = [N,C,H,W],
 * `squeezed` = [N,C*f*f,H/f,W/f]; reverse=0 reads full (src) and writes squeezed (dst), reverse=1 the
 * other way.  Each side is NCHW (layout 0; ld = batch stride in elements) or rows (layout 1; ld = row
 * pitch in floats, >= that side's channel count: only its first channels are touched).  factor=1 is a
 * pure layout change. */
int glowk_rows_squeeze(const float* src, int src_layout, int64_t src_ld, float* dst, int dst_layout, int64_t dst_ld,
                       int64_t N, int64_t C, int64_t H, int64_t W, int factor, int reverse, void* stream);
/* The entry squeeze of Glow.normal_flow with the dequantisation folded in (network/model.py:419-423 + the first
 * Squeeze2d): dst = squeeze(src + add); `add` is the U(0, 1/n_bins) noise, laid out like src. */
int glowk_rows_squeeze_add(const float* src, const float* add, int src_layout, int64_t src_ld, float* dst,
                           int dst_layout, int64_t dst_ld, int64_t N, int64_t C, int64_t H, int64_t W, int factor,
                           void* stream);

/* One launch for many glowk_pack_conv_weight / glowk_unpack_weight_grad calls.  jobs: device array of
 *   struct { const float* w; void* packed; int32 O, I, ks, layout, rows, ld; int64 block0; }   (48 bytes)
 * sorted by block0 = index of the job's first CTA; total_blocks = sum over jobs.  Packing uses one CTA per
 * 32 x 32 (out x in channel) tile, i.e. ceil(O/32)*ceil(I/32) CTAs per job, and writes only the O*I*ks*ks data
 * elements (the padding of `packed` must already be zero); unpacking uses one CTA per 256 of the O*I*ks*ks
 * gradient elements and ACCUMULATES into w (a gradient). */
int glowk_pack_conv_weights_batched(const void* jobs, int64_t njobs, int64_t total_blocks, int act_dtype, void* stream);
int glowk_unpack_weight_grads_batched(const void* jobs, int64_t njobs, int64_t total_blocks, void* stream);

/* Gradient finish of the ActNorm that follows a coupling-net conv (Conv2d, module.py:188-260; ActNorm.forward
 * module.py:122-149), batched over layers.  glowk_gemm(GLOWK_EPI_RELU_BWD) on the bf16 path may be called with
 * dlogs == NULL and dbias == NULL or pointing at per-pass scratch `db`; this call then applies, per output channel n,
 *     dbias[n] += db[n];   dlogs[n] += f * ( <w[n,:], dw[n,:]> + bias[n]*db[n] )
 * which equals f * sum_m g*y of the direct epilogue reduction (y = (W a + b) s on the ReLU-active set).
 * jobs: device array of
 *   struct { const bf16* w; const float* dw; const float* bias; const float* db; float* dbias; float* dlogs;
 *            int32 N, K, ldw, lddw; float f; int32 db_stride; }                                      (72 bytes)
 * w = the GEMM-layout bf16 weight of the forward pass, dw = THIS pass's weight gradient in the same [N][K] order
 * (K, ldw, lddw even); db[n * db_stride] = this pass's bias gradient (db_stride = 1 for a vector the dgrad epilogue
 * reduced into, lddw when it is the ones column of dw, glowk_im2col_rows_ones); max_n = max over jobs of N. */
int glowk_conv_actnorm_finish_batched(const void* jobs, int64_t njobs, int64_t max_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLOWK_H_ */
