/*
 * cmfb200.h -- C ABI of libcmfb200.so: sm_100a kernels for the `cmfsm` stereo hot path.
 *
 * The reference (lidongyv/Explicit-Context-Mapping-for-Stereo-Matching) is pure PyTorch and has no
 * FFI of its own; each entry point below replaces the block of ATen calls cited next to it
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (plain cudaMalloc / torch allocations), 16-byte aligned;
 *   - tensors are dense, row-major in the layout named in the comment;
 *   - the last argument is the CUDA stream (cudaStream_t passed as void*), work is enqueued
 *     asynchronously: no allocation, no host synchronisation, no global mutable state;
 *   - return value: 0 = ok, <0 = error; `cmfb200_last_error()` returns a thread-local message.
 *   - fp32 arithmetic unless the name says otherwise.
 */
#ifndef CMFB200_H_
#define CMFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMFB200_OK 0
#define CMFB200_ERR_INVALID (-1) /* bad shape / null pointer / unsupported configuration */
#define CMFB200_ERR_CUDA (-2)    /* a CUDA runtime call failed (message has the cudaError string) */

#define CMFB200_ABI_VERSION 3

/* ABI version of the loaded library (== CMFB200_ABI_VERSION of the header it was built from). */
int cmfb200_abi_version(void);
/* Thread-local, NUL-terminated description of the last error returned on this thread. */
const char* cmfb200_last_error(void);
/* Number of kernel launches enqueued by this library since load (all threads); bench.py reports it. */
unsigned long long cmfb200_launch_count(void);

/* ---- K1: concat cost volume ----------------------------------------------------------------
 * Replaces cmf/models/cmfsm.py:667-682 (CPU zeros + H2D copy + 2*D slice copies).
 *   cost[b, c,   d, y, x] = L[b,c,y,x]     if x >= d else +0.0
 *   cost[b, C+c, d, y, x] = R[b,c,y,x-d]   if x >= d else +0.0
 * L, R: [B,C,h,w] fp32 NCHW.  cost: [B,2C,D,h,w] fp32 NCDHW (every element written). w % 4 == 0. */
int cmfb200_cost_volume_concat_fwd(const float* L, const float* R, float* cost,
                                   int B, int C, int h, int w, int D, void* stream);
/* Adjoint (autograd of the slice copies): dL[b,c,y,x] = sum_{d<=x} g[b,c,d,y,x],
 * dR[b,c,y,x] = sum_{d, x+d<w} g[b,C+c,d,y,x+d].  g: [B,2C,D,h,w]; dL,dR: [B,C,h,w] (overwritten). */
int cmfb200_cost_volume_concat_bwd(const float* g, float* dL, float* dR,
                                   int B, int C, int h, int w, int D, void* stream);

/* ---- K2: 3-D convolutions of the aggregation network ------------------------------------------
 * Replace nn.Conv3d(k=3,pad=1,bias=False) in convbn_3d (cmfsm.py:49-58; dres0/1 :604-613, hourglass
 * :244-259, classif :621-634) and nn.ConvTranspose3d(k=3,s=2,p=1,op=1) (hourglass conv5/conv6 :261-281).
 *
 * Weights are consumed in the packed layout  wp[Cin][27][Cout]  (tap = kd*9+kh*3+kw):
 *   conv   : wp[ci][t][co] = weight[co][ci][t]      (nn.Conv3d layout          [Cout,Cin,3,3,3])
 *   deconv : wp[ci][t][co] = weight[ci][co][t]      (nn.ConvTranspose3d layout [Cin,Cout,3,3,3]) */
int cmfb200_pack_conv3d_weight(const float* weight, float* packed, int Cout, int Cin,
                               int transposed, void* stream);

/* x: [B,Cin,D,H,W] NCDHW; y: [B,Cout,Do,Ho,Wo], Xo = (X-1)/stride + 1; stride in {1,2};
 * Cin % 8 == 0; Cout in {1, 32, 64}.
 * If gn_sums != NULL it must be a ZEROED [B,Cout,2] double buffer: the kernel accumulates per
 * (b,channel) sum and sum-of-squares of y into it (GroupNorm statistics fused into the epilogue). */
int cmfb200_conv3d_k3_fwd(const float* x, const float* packed_w, float* y, double* gn_sums,
                          int B, int Cin, int Cout, int D, int H, int W, int stride, void* stream);
/* Backward of the classifier tail conv nn.Conv3d(Cin, 1, 3, padding=1) (training): x [B,Cin,D,H,W], weight
 * [1,Cin,3,3,3] (unpacked), grad_y [B,1,D,H,W] -> dx [B,Cin,D,H,W], dw [1,Cin,3,3,3] (zeroed here).  Cin = 32. */
int cmfb200_conv3d_cout1_bwd(const float* x, const float* weight, const float* grad_y, float* dx, float* dw, int B,
                             int Cin, int D, int H, int W, void* stream);

/* Row-window variants for row-band sharding of ONE image pair (cmf_b200.parallel): the input tensor carries
 * halo rows received from the neighbouring ranks, and only the rows this rank owns are produced, so the GroupNorm
 * sums fused into the epilogue cover exactly the band (no crop copy, no separate statistics pass).
 *   conv:   x [B,Cin,D,H_in,W], y [B,Cout,Do,H_out,Wo]; output row m reads input rows m*stride - 1 + h_offset + kh
 *           (h_offset = 1 with one halo row on top for stride 1, 2 with two halo rows for stride 2); rows outside
 *           [0,H_in) read zero.  D and W are padded as usual.
 *   deconv: x [B,Cin,D,H_in,W] whose LAST row may be a halo row, y [B,Cout,2D,2*H_compute,2W]. */
int cmfb200_conv3d_k3_rows_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin,
                               int Cout, int D, int H_in, int W, int stride, int h_offset, int H_out, void* stream);
int cmfb200_deconv3d_k3s2_rows_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin,
                                   int Cout, int D, int H_in, int H_compute, int W, void* stream);
/* Transposed conv k3 s2 p1 op1: y: [B,Cout,2D,2H,2W].  Cout in {32,64}; same gn_sums contract. */
int cmfb200_deconv3d_k3s2_fwd(const float* x, const float* packed_w, float* y, double* gn_sums,
                              int B, int Cin, int Cout, int D, int H, int W, void* stream);

/* ---- K2 throughput path: bf16 tcgen05/TMEM/TMA implicit GEMM on the "C8" layout -------------------------
 * C8 layout: bf16 [B][C/8][D][H][W][8] (a voxel's 8-channel group is one 16-byte unit).  All *_c8 pointers
 * below are bf16 device buffers in that layout (void* in the ABI: no CUDA types in the header).
 * pack: nn.Conv3d [Cout,Cin,3,3,3] (transposed=0) fp32 -> bf16 [27][Cin/8][Cout][8]. */
int cmfb200_pack_igemm_weight_bf16(const float* weight, void* packed, int Cout, int Cin, int transposed,
                                   void* stream);
/* 3x3x3 stride-1 pad-1 conv, bf16 operands, fp32 accumulation (tcgen05.mma, accumulators in TMEM); output =
 * RAW conv result rounded to bf16, C8.  (Cin,Cout) in {(32,32),(64,32),(64,64)}.  gn_sums: ZEROED [B,Cout,2]
 * double buffer receiving sum / sum of squares of the stored (rounded) values, or NULL. */
int cmfb200_conv3d_igemm_bf16_fwd(const void* x_c8, const void* packed_w, void* y_c8, double* gn_sums,
                                  int B, int Cin, int Cout, int D, int H, int W, void* stream);
/* classifN.2 (32->1) on tensor cores (depth-stacked tcgen05 schedule of the 32->32 layers; kept as the fp32-output probe
 * of the tensor-core accumulator, tools/probe_tc_rounding.py -- the model runs this layer in fp32): packed_w32 is
 * cmfb200_pack_igemm_weight_bf16 of the [1,Cin,3,3,3] weight zero-padded to 32 output channels (so the weight is
 * rounded to bf16 like every other layer of the bf16 aggregation); y fp32 [B][D][H][W].  Cin = 32. */
int cmfb200_conv3d_igemm_cout1_bf16_fwd(const void* x_c8, const void* packed_w32, float* y, int B, int Cin, int D,
                                        int H, int W, void* stream);
/* Transposed conv k3 s2 p1 op1 on tensor cores: x_c8 [B][Cin/8][D][H][W][8] -> y_c8 [B][Cout/8][2D][2H][2W][8];
 * packed_w from cmfb200_pack_igemm_weight_bf16(transposed=1).  (Cin,Cout) in {(64,64),(64,32)}. */
int cmfb200_deconv3d_igemm_bf16_fwd(const void* x_c8, const void* packed_w, void* y_c8, double* gn_sums,
                                    int B, int Cin, int Cout, int D, int H, int W, void* stream);
/* Stride-2 conv k3 p1 on tensor cores.  The input is the PARITY-SPLIT copy of the C8 tensor,
 * [B][8][Cin/8][D/2][H/2][W/2][8] with parity index (d&1)*4+(h&1)*2+(w&1) (cmfb200_c8_parity_split); the output
 * is plain C8 [B][Cout/8][Do][Ho][Wo][8] with Do=D/2 etc.  (Cin,Cout) in {(32,64),(64,64)}. */
int cmfb200_conv3d_s2_igemm_bf16_fwd(const void* x_split_c8, const void* packed_w, void* y_c8, double* gn_sums,
                                     int B, int Cin, int Cout, int Do, int Ho, int Wo, void* stream);
int cmfb200_c8_parity_split(const void* x_c8, void* y_split_c8, int B, int C, int D, int H, int W, void* stream);
/* K1 in C8/bf16: same contract as cmfb200_cost_volume_concat_fwd, output [B][2C/8][D][h][w][8] bf16
 * (channel groups 0..C/8-1 = masked left features, C/8.. = shifted right features). */
int cmfb200_cost_volume_concat_c8_bf16(const float* L, const float* R, void* cost_c8,
                                       int B, int C, int h, int w, int D, void* stream);
/* K3 on C8/bf16: y = GroupNorm(x) (+residual) (ReLU), fp32 math, bf16 in/out; y may alias x.  If y_split_c8 != NULL
 * the result is ALSO written in the parity-split layout a stride-2 consumer reads (D,H,W even).  If y_f32 != NULL the
 * UN-ROUNDED fp32 result is also written as [B][C][D][H][W] (input of the fp32 classifier tail classifN.2, whose
 * logits are the rounding-sensitive spot of the bf16 mode); y_c8 may then be NULL. */
int cmfb200_gn_apply_c8_bf16(const void* x_c8, const double* gn_sums, const float* gamma, const float* beta,
                             const void* residual_c8, void* y_c8, void* y_split_c8, float* y_f32, int B, int C, int G,
                             int D, int H, int W, float eps, int relu, void* stream);
/* layout/dtype converters between C8/bf16 and dense fp32 [B,C,spatial] (NCDHW). */
int cmfb200_c8_bf16_to_f32(const void* x_c8, float* y, int B, int C, long long spatial, void* stream);
int cmfb200_f32_to_c8_bf16(const float* x, void* y_c8, int B, int C, long long spatial, void* stream);

/* ---- 2-D convolutions of the feature extractor ----------------------------------------------------
 * Replace nn.Conv2d(bias=False) in convbn / BasicBlock / feature_extraction (cmfsm.py:36-46, 61-85, 126-236).
 * Packed weights wp[Cin][k*k][Cout] from nn.Conv2d's [Cout,Cin,k,k].
 * x: [B,Cin,H,W] NCHW; y: [B,Cout,Ho,Wo]; padding = (k/2)*dilation (what convbn() computes);
 * (ksize,stride,dilation) in {(3,1,1),(3,2,1),(3,1,2),(3,1,4),(1,1,1),(1,2,1)}; Cin = 3 (stem, Cout 32) or a multiple
 * of 8; Cout = 32 or a multiple of 64.  gn_sums: same contract as cmfb200_conv3d_k3_fwd. */
int cmfb200_pack_conv2d_weight(const float* weight, float* packed, int Cout, int Cin, int ksize, void* stream);
int cmfb200_conv2d_fwd(const float* x, const float* packed_w, float* y, double* gn_sums,
                       int B, int Cin, int Cout, int H, int W, int ksize, int stride, int dilation, void* stream);
/* Row-window variant for row-band sharding (see cmfb200_conv3d_k3_rows_fwd): x [B,Cin,H_in,W] carries halo rows, y
 * [B,Cout,H_out,Wo]; output row m reads input rows m*stride - pad + h_offset + kh*dilation. */
int cmfb200_conv2d_rows_fwd(const float* x, const float* packed_w, float* y, double* gn_sums, int B, int Cin, int Cout,
                            int H_in, int W, int ksize, int stride, int dilation, int h_offset, int H_out, void* stream);

/* SPP tail of the feature extractor (cmfsm.py:152-170, 207-233).
 * pool: x [B,C,H,W] -> average pools with kernel = stride = 8/16/32/64 (floor mode): p8 [B,C,H/8,W/8] ... p64.
 * upsample_concat: cat [B,320,H,W] = [raw(64) | skip(128) | up(b4) | up(b3) | up(b2) | up(b1)], b4..b1 =
 * [B,32,H_full/8,W/8] .. [B,32,H_full/64,W/64], bilinear, align_corners=False (F.interpolate semantics);
 * raw/skip/cat hold rows [y_off, y_off+H) of an image of height H_full (H_full=H, y_off=0: whole image). */
int cmfb200_spp_pool_fwd(const float* x, float* p8, float* p16, float* p32, float* p64,
                         int B, int C, int H, int W, void* stream);
int cmfb200_spp_upsample_concat_fwd(const float* raw, const float* skip, const float* b4, const float* b3,
                                    const float* b2, const float* b1, float* cat, int B, int H, int W,
                                    int H_full, int y_off, void* stream);
/* The same op with explicit branch-map sizes [B,32,h*,w*] and raw-channel count raw_c in {64,128} (cmfsm_sub_8: SPP
 * pools 8/16/32/4 at 1/8 resolution, cmfsm_sub_8.py:152-170, 207-231; cmfsm_sub_16: 128 raw channels,
 * cmfsm_sub_16.py:207-236): cat [B, raw_c+256, H, W] = [raw | skip | up(ba) | up(bb) | up(bc) | up(bd)]. */
int cmfb200_spp_upsample_concat_sized_fwd(const float* raw, const float* skip, const float* ba, const float* bb,
                                          const float* bc, const float* bd, float* cat, int B, int raw_c, int H, int W,
                                          int ha, int wa, int hb, int wb, int hc, int wc, int hd, int wd, void* stream);

/* ---- K3: GroupNorm (+ residual add) (+ ReLU) ---------------------------------------------------
 * Replaces nn.GroupNorm(32,C) (cmfsm.py:58,269,280), the residual adds (:288,297,299,685,687,690,693)
 * and nn.ReLU / F.relu around them.
 * gn_stats: per-(b,channel) sum / sum of squares of x [B,C,spatial] into a ZEROED [B,C,2] double buffer
 * (only needed when the producer did not fuse them). */
int cmfb200_gn_stats(const float* x, double* gn_sums, int B, int C, long long spatial, void* stream);
/* y = ((x - mean_g) * rstd_g) * gamma[c] + beta[c]  (+ residual)  (then max(.,0) if relu)
 * mean/rstd per (b, group of C/G consecutive channels) from gn_sums; biased variance, eps inside sqrt.
 * y may alias x.  residual may be NULL. */
int cmfb200_gn_apply(const float* x, const double* gn_sums, const float* gamma, const float* beta,
                     const float* residual, float* y, int B, int C, int G, long long spatial,
                     float eps, int relu, void* stream);
/* GroupNorm backward (training).  grad_out = gradient of out = [relu](gamma*xhat + beta [+ residual]); x = the conv
 * output the forward normalised, gn_sums = its forward statistics; out_or_null = the forward output when a ReLU was
 * applied (mask out > 0), else NULL.  Writes dx (gradient of x), optionally dres (= masked grad_out, the gradient of
 * the residual input) and bwd_sums [B][C][2] doubles = { sum g', sum g'*xhat } per (b,channel) (zeroed here), from
 * which d_beta[c] = sum_b bwd_sums[b][c][0], d_gamma[c] = sum_b bwd_sums[b][c][1]. */
int cmfb200_gn_bwd(const float* grad_out, const float* x, const float* out_or_null, const double* gn_sums,
                   const float* gamma, double* bwd_sums, float* dx, float* dres_or_null, int B, int C, int G,
                   long long spatial, float eps, void* stream);

/* ---- K5: context-mapping weights -----------------------------------------------------------------
 * Replaces eight_related_context_mapping.forward + similarity_measure1 (cmfsm.py:443-593, 304-358).
 * lr: [B,32,h,w] (1/scale-res features), hr: [B,32,H,W] (full-res firstconv output), H=h*scale, W=w*scale,
 * scale in {4} (even).  w0:[32,66] w1:[16,32] w2:[8,16] w3:[1,8] (1x1 conv weights, no bias).
 * weights9: [B,9,H,W] softmax over the 9 neighbours (order c,l,r,t,b,lt,rt,lb,rb); out-of-image
 * neighbours take the constant logit -100.  [valid_y0, valid_y1) = low-res rows that are inside the IMAGE
 * (0,h for a whole image; a row band passes its halo rows in lr/hr and marks them here). */
int cmfb200_ctxmap_weights_fwd(const float* lr, const float* hr, const float* w0, const float* w1,
                               const float* w2, const float* w3, float* weights9,
                               int B, int h, int w, int scale, int valid_y0, int valid_y1, void* stream);
/* K5 backward (training; replaces autograd through cmfsm.py:443-593).  weights9 = the forward output, grad_weights9 =
 * its incoming gradient (both [B,9,H,W]).  With a0 = W0[:,0:32].lr(cell) + W0[:,32:64].hr(pixel) + W0[:,64:66].code
 * the kernel writes d_ahr [B,32,H,W] = dL/d(W0[:,32:64].hr), d_alr [B,32,h,w] = dL/d(W0[:,0:32].lr) and
 * d_wbuf[712] = { dW1[16][32], dW2[8][16], dW3[8], dW0[:,64:66] as [32][2] } (fp32, summed with atomics, zeroed
 * here).  The two remaining linear maps (d hr, d lr, dW0[:,0:64]) are 1x1 GEMMs left to the caller.  scale = 4. */
int cmfb200_ctxmap_weights_bwd(const float* lr, const float* hr, const float* w0, const float* w1, const float* w2,
                               const float* w3, const float* weights9, const float* grad_weights9, float* d_ahr,
                               float* d_alr, float* d_wbuf, int B, int h, int w, int scale, void* stream);
/* cmfsm_sub_8 variant (six_related_context_mapping, cmfsm_sub_8.py:440-572, reference-image half): five neighbours
 * centre, right, left, top, bottom; the MLP ends in a LeakyReLU; logit 0 where the neighbour cell is outside the
 * image; weights5 [B,5,H,W] = softmax(logits) * logits (NOT normalised).  Any even scale. */
int cmfb200_ctxmap_weights5_fwd(const float* lr, const float* hr, const float* w0, const float* w1, const float* w2,
                                const float* w3, float* weights5, int B, int h, int w, int scale, void* stream);
/* The TARGET-image half of the same module (three neighbours centre, right, left of the right image's own feature
 * maps; used by cmfsm_sub_16, cmfsm_sub_16.py:488-573): weights3 [B,3,H,W] = softmax(logits) * logits. */
int cmfb200_ctxmap_weights3_fwd(const float* lr, const float* hr, const float* w0, const float* w1, const float* w2,
                                const float* w3, float* weights3, int B, int h, int w, int scale, void* stream);
/* cmfsm_sub_16 epilogue (cmfsm_sub_16.py:760-850): the three classifier volumes c_n [B,D',h,w] are accumulated
 * (c2 += c1, c3 += c2), nearest-upsampled by `scale` along d, y, x, mixed over the five spatial neighbours with
 * weights5, then over the disparity axis with the target weights shifted by the disparity
 * (v[d,y,x] = weights3[.,y,x-d] for x >= d, 1 elsewhere; taps d, d+scale (left weight), d-scale (right weight)),
 * and regressed by a softmax over all maxdisp = D'*scale planes.  out_n: [B,H,W].  One launch, nothing of the
 * [B,maxdisp,H,W] volumes is materialised. */
int cmfb200_volume_mapping_fwd(const float* c1, const float* c2, const float* c3, const float* weights5,
                               const float* weights3, float* out1, float* out2, float* out3, int B, int Dl, int h,
                               int w, int scale, void* stream);
/* Epilogue of the bilinear_cmf / bilinear_cmf_sub_8 / bilinear_cmf_sub_16 baselines (bilinear_cmf.py:418-452): the
 * classifier volumes c_n [B,D',h,w] are accumulated (c2 += c1, c3 += c2), trilinearly upsampled (align_corners=False)
 * to [B,maxdisp,H,W], soft-maxed over maxdisp and regressed; out_n [B,H,W].  Nothing is materialised. */
int cmfb200_trilinear_softargmin_fwd(const float* c1, const float* c2, const float* c3, float* out1, float* out2,
                                     float* out3, int B, int Dl, int h, int w, int maxdisp, int H, int W, void* stream);

/* ---- K4: soft-argmin + x scale upsample + 9-neighbour context mapping ---------------------------
 * Replaces cmfsm.py:703-769 (3x softmax, disparityregression :111-123, ~60 slice kernels).
 * c1,c2,c3: raw classifier volumes [B,D,h,w]; cost1=c1, cost2=c2+cost1, cost3=c3+cost2.
 * p_i[b,cy,cx] = sum_d d*softmax_d(cost_i);  out_i[b,0,y,x] = sum_k w_k[b,y,x]*scale*p_i[b,y/s+dy_k,x/s+dx_k]
 * over in-image neighbours.  weights9: [B,9,H,W]; out1..3: [B,1,H,W]; pred_lr (optional, may be NULL):
 * [3,B,h,w] receives p_i.  scale % 4 == 0. */
int cmfb200_softargmin_ctxmap_fwd(const float* c1, const float* c2, const float* c3,
                                  const float* weights9, float* out1, float* out2, float* out3,
                                  float* pred_lr, int B, int D, int h, int w, int scale, void* stream);
/* K4 backward (training): g_n = gradients of the three outputs [B,1,H,W]; writes dc_n [B,D,h,w] (gradients of the RAW
 * classifier volumes, i.e. the cumulative sums of the forward are back-propagated too) and dweights9 [B,9,H,W].
 * One launch, no atomics, every output element written. */
int cmfb200_softargmin_ctxmap_bwd(const float* c1, const float* c2, const float* c3, const float* weights9,
                                  const float* g1, const float* g2, const float* g3, float* dc1, float* dc2, float* dc3,
                                  float* dweights9, int B, int D, int h, int w, int scale, void* stream);
/* cmfsm_sub_8 variant (cmfsm_sub_8.py:757-802): the three volumes are regressed independently (no cumulative sums)
 * and mapped with the five weights of cmfb200_ctxmap_weights5_fwd (order c,r,l,t,b). */
int cmfb200_softargmin_ctxmap5_fwd(const float* c1, const float* c2, const float* c3, const float* weights5,
                                   float* out1, float* out2, float* out3, float* pred_lr, int B, int D, int h, int w,
                                   int scale, void* stream);

/* ---- fp32-accurate tensor-core convolution ("tc3": three-term bf16 split on tcgen05) ------------------------------
 * Replaces the stride-1 nn.Conv2d layers of feature_extraction (cmf/models/cmfsm.py:126-236: convbn / BasicBlock /
 * firstconv / lastconv) and the stride-1 nn.Conv3d layers of the aggregation network (cmfsm.py:49-58, 240-303,
 * 604-634) in the fp32 parity mode: every fp32 operand is the exact sum of three bf16 terms, six bf16 MMAs per
 * product accumulate in fp32 (csrc/conv_tc3.cu).  2-D images are volumes with D = 1.
 *   activations "C8S3": bf16 [B][C/8][3][D][H][W][8];  raw conv output "C8F": fp32 [B][C/8][D][H][W][8]. */
/* weight: nn.Conv2d [Cout][Cin][k][k] (KD = 1) or nn.Conv3d [Cout][Cin][3][3][3] (KD = 3), fp32 ->
 * packed bf16 [KD*Cin/16][k][k][2][3][Cout][8].  Cin % 16 == 0, Cout % 8 == 0, k in {1,3}. */
int cmfb200_pack_tc3_weight(const float* weight, void* packed, int Cout, int Cin, int KD, int KHW, void* stream);
/* Stride-1 "same" convolution (padding = dilation*(k/2), depth padding 1 when KD = 3).  y: C8F, or plain
 * [B][Cout][D][H][W] fp32 when out_nchw != 0; gn_sums (optional): [B][Cout][2] doubles (sum, sum of squares of y),
 * must be zero on entry.  Supported: 3x3 d1 Cout 32/64/128, 3x3 d2 Cout 128, 1x1 Cout 32/128. */
int cmfb200_conv_tc3_fwd(const void* x_c8s3, const void* packed_w, float* y, double* gn_sums, int B, int Cin, int Cout,
                         int D, int H, int W, int KD, int KHW, int dilation, int out_nchw, void* stream);
/* Row-window form for row-band sharding: the input has H rows (the band plus halo rows received from the neighbour
 * ranks, zeros at the image border); only H_out rows are produced -- output row h is centred on input row h + row_off --
 * and only they enter gn_sums, so the epilogue's statistics are the band's statistics. */
int cmfb200_conv_tc3_rows_fwd(const void* x_c8s3, const void* packed_w, float* y, double* gn_sums, int B, int Cin, int Cout,
                              int D, int H, int W, int KD, int KHW, int dilation, int out_nchw, int row_off, int H_out,
                              void* stream);
/* y = GroupNorm(raw) (+residual) (ReLU) (nn.GroupNorm + the adds / ReLUs of convbn, BasicBlock, hourglass):
 * raw is C8F (raw_is_c8f) or [B][C][spatial] fp32; residual as C8S3 and/or [B][C][spatial] fp32 (either may be NULL);
 * the result is written split into three bf16 terms (y_c8s3) and/or as [B][C][spatial] fp32 (y_nchw).
 * gn_sums == NULL: no normalisation (layout conversion / split only). */
int cmfb200_gn_apply_tc3(const float* raw, int raw_is_c8f, const double* gn_sums, const float* gamma, const float* beta,
                         const void* residual_c8s3, const float* residual_nchw, void* y_c8s3, float* y_nchw, int B,
                         int C, int groups, long long spatial, float eps, int relu, void* stream);
/* push_up / push_dn (optional, row bands): DEVICE POINTERS INTO THE NEIGHBOUR RANKS' MEMORY (peer-mapped over NVLink): the
 * first / last push_rows rows of the C8S3 result are also stored there, dense [B*C/8][3][D][push_rows][W][8] -- the halo
 * exchange of the next conv is fused into this kernel's epilogue (the consumer only waits on a signal and copies locally).
 * y_split_c8s3 (optional): the result ALSO as the parity-split copy [B][8][C/8][3][D/2][H/2][W/2][8] that the stride-2
 * tensor-core conv reads (D, H, W even; H and W must then be passed).
 * Row-band forms: y_c8s3 / residual_c8s3 (and the K1 output) carry `pad` extra rows above and below the H rows of every
 * depth plane ([...][D][H + 2 pad][W][8]); the kernels fill the H interior rows, the halo rows are filled by the halo
 * exchange -- no re-copy of the activation to attach them.  pad = 0 is the dense form.  nchw_padded != 0: y_nchw
 * carries the same `pad` rows ([B][C][D][H + 2 pad][W]; the row-window 32->1 tail reads its halo rows in place). */
int cmfb200_gn_apply_tc3_padded(const float* raw, int raw_is_c8f, const double* gn_sums, const float* gamma,
                                const float* beta, const void* residual_c8s3, const float* residual_nchw, void* y_c8s3,
                                float* y_nchw, int B, int C, int groups, long long spatial, float eps, int relu, int pad,
                                int H, int W, void* y_split_c8s3, void* push_up, void* push_dn, int push_rows,
                                int nchw_padded, void* stream);
int cmfb200_cost_volume_concat_c8s3_padded(const float* L, const float* R, void* cost_c8s3, int B, int C, int h, int w,
                                           int D, int pad, void* stream);
/* K1 written directly as C8S3 (cmfsm.py:667-682): cost [B][2C/8][3][D][h][w][8] bf16; the three terms of an element
 * sum to the fp32 feature value exactly (or to +0.0 in the masked triangle). */
int cmfb200_cost_volume_concat_c8s3(const float* L, const float* R, void* cost_c8s3, int B, int C, int h, int w, int D,
                                    void* stream);

/* ---- K1, correlation form: corr[b,d,y,x] = (1/C) sum_c L[b,c,y,x] R[b,c,y,x-d] for x >= d, +0.0 otherwise; with
 * normalize != 0 the cosine similarity sum_c L R / max(|L| |R|, 1e-8) instead.  The reference has no live correlation
 * path (its models build the concat volume); this is the cosine matching of the dead file
 * "cmf/models/rstereo # dense volume match.py":309-311 with K1's shift / mask convention.  corr: [B][D][h][w]. */
int cmfb200_cost_volume_corr_fwd(const float* L, const float* R, float* corr, int B, int C, int h, int w, int D,
                                 int normalize, void* stream);

/* ---- fused masked smooth-L1 training loss (train.py:162-174) --------------------------------------------------------
 * valid = (disp < maxdisp) & (disp > 0).  fwd: sums4 (ZEROED, doubles) += [sum_valid smooth_l1(out_i - disp) for i = 1..3,
 * number of valid pixels]; the caller forms loss = sum_i weight_i * sums[i] / count (count all-reduced over ranks in
 * data-parallel training).  bwd: g_i = scale3[i] * clamp(out_i - disp, -1, 1) on valid pixels, 0 elsewhere, with
 * scale3 (DEVICE, 3 floats) = weight_i * dLoss / count.  n = number of pixels (B*H*W). */
int cmfb200_masked_smooth_l1_fwd(const float* out1, const float* out2, const float* out3, const float* disp, double* sums4,
                                 long long n, float maxdisp, void* stream);
int cmfb200_masked_smooth_l1_bwd(const float* out1, const float* out2, const float* out3, const float* disp,
                                 const float* scale3, float* g1, float* g2, float* g3, long long n, float maxdisp,
                                 void* stream);

/* ---- pitched block copy for the row-band halo exchange (cmf_b200/parallel.py) ---------------------------
 * `height` blocks of `width` contiguous bytes, dst_pitch / src_pitch bytes apart; everything a multiple of 16 bytes.
 * dst or src may be a peer GPU's memory mapped through the symmetric-memory mailbox (NVLink loads / stores). */
int cmfb200_copy_2d(void* dst, long long dst_pitch, const void* src, long long src_pitch, long long width, long long height,
                    void* stream);
/* GroupNorm statistics of a row-band forward: dst[i] = scale * sum_r srcs[r][i] (doubles), r = 0..n-1 in rank order on every
 * rank; srcs = HOST array of n <= 16 device pointers (the ranks' mailbox buffers, peer-mapped). */
int cmfb200_sum_peers_f64(double* dst, const void* const* srcs, int n, int count, double scale, void* stream);

/* ---- weight gradient of every convolution (training backward; replaces aten::convolution_backward) ------------------
 * dW[co][ci][kd][kh][kw] = sum_{b,o} dy[b][co][o] * x[b][ci][o*stride - pad + k*dilation], "same" padding (dilation*(k/2),
 * depth padding 1 when KD = 3).  x: [B][Cin][D][H][W] (D = 1 with KD = 1), dy: [B][Cout][Do][Ho][Wo] with
 * o = (i - 1)/stride + 1 per axis; dw: [Cout][Cin][KD][k][k], MUST BE ZERO on entry (partial sums are added with fp32
 * atomics).  Cout % 32 == 0.  Supported: 3x3x3 s1/s2; 3x3 s1 d1/d2, 3x3 s2, 1x1 s1/s2.
 * Transposed conv (weight [Cin_t][Cout_t][27], k3 s2 p1 op1): call with x := grad of the OUTPUT, dy := the layer INPUT,
 * Cin := Cout_t, Cout := Cin_t, stride 2 -- the result has the ConvTranspose3d weight layout.
 * Stages are fetched by TMA when W and Wo are multiples of 4 and x, dy are 16-byte aligned (every layer of the SceneFlow /
 * KITTI-padded shapes down to 1/8 resolution); otherwise by 4-byte cp.async -- same arithmetic, same result layout. */
int cmfb200_conv_wgrad(const float* x, const float* dy, float* dw, int B, int Cin, int Cout, int D, int H, int W, int KD,
                       int KHW, int stride, int dilation, void* stream);

/* ---- fp32-accurate stride-2 / transposed 3x3x3 convs of the hourglasses on the tensor cores (csrc/conv_tc3_s2.cu) ----
 * Same arithmetic and layouts as cmfb200_conv_tc3_fwd.  Stride 2 (hourglass.conv1 / conv3, cmfsm.py:244-254): the input is
 * the PARITY-SPLIT C8S3 copy written by cmfb200_gn_apply_tc3_padded (y_split_c8s3); (Do,Ho,Wo) = output = cell grid;
 * (Cin,Cout) in {(32,64),(64,64)}.  Transposed (conv5 / conv6, :261-281, k3 s2 p1 op1): x_c8s3 [B][Cin/8][3][D][H][W][8] ->
 * y_c8f [B][Cout/8][2D][2H][2W][8], eight launches (one per output parity class); (Cin,Cout) in {(64,64),(64,32)}.
 * packed_w: bf16, three terms, in the K-step order of the kernels (cmf_b200.ops.pack_tc3_s2_weight /
 * pack_tc3_deconv_weight).  gn_sums as for cmfb200_conv_tc3_fwd. */
int cmfb200_conv_tc3_s2_fwd(const void* x_split_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B, int Cin,
                            int Cout, int Do, int Ho, int Wo, void* stream);
int cmfb200_deconv_tc3_fwd(const void* x_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B, int Cin, int Cout,
                           int D, int H, int W, void* stream);
/* Row-band forms: the inputs carry `pad` spare rows (cell rows for the parity-split input) above and below the H (Ho)
 * rows; the halo exchange fills the one row each layer needs (stride 2: the cell row above; transposed: the row below). */
int cmfb200_conv_tc3_s2_rows_fwd(const void* x_split_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B,
                                 int Cin, int Cout, int Do, int Ho, int Wo, int pad, void* stream);
int cmfb200_deconv_tc3_rows_fwd(const void* x_c8s3, const void* packed_w, float* y_c8f, double* gn_sums, int B, int Cin,
                                int Cout, int D, int H, int W, int pad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMFB200_H_ */
