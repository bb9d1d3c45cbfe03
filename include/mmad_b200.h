/* mmad_b200 — C-ABI of the B200-native hot path of dongzj56/Multimodal_AD.
 *
 * Drop-in boundary.  Everything crosses as plain pointers and sizes; no
 * torch / C++ types.  Device pointers are CUDA device addresses of the
 * current device; `stream` is a cudaStream_t passed as void* (NULL = legacy
 * default stream).  Every function returns 0 on success or a negative
 * MMAD_E* code; mmad_last_error() gives the message of the calling thread's
 * last failure.  There is no CPU fallback anywhere behind this header.
 *
 * Part 1  — atlas ROI pooling
 *   replaces /root/reference/image_features.py:80-82 (one-hot atlas mask),
 *   :111-114 (masked sum / clamp_min(count,1e-6) -> (B,R,C) ROI means) and
 *   the empty /root/reference/models/ROI_pol.py the north star names.
 * Part 2  — Conv3d / BatchNorm3d / ReLU stacks of the 3D-CNN image branch
 *   replaces the torch.nn.Conv3d / BatchNorm3d / ReLU / MaxPool3d calls of
 *   /root/reference/models/resnet.py:14-23,40-109,112-215 (and resnet18.py,
 *   ImageEncoder.py which repeat them) and of /root/reference/models/unet3d.py:14-46,
 *   51-84,116-157 (Conv3d with bias, MaxPool3d(2,2), ConvTranspose3d(2,2), channel
 *   concatenation, pad to 96x112x96 / crop back), forward and backward.
 */
#ifndef MMAD_B200_H
#define MMAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMAD_OK            0
#define MMAD_EINVAL       -1   /* bad argument (shape, range, alignment) */
#define MMAD_ECUDA        -2   /* a CUDA runtime / driver call failed */
#define MMAD_ENOMEM       -3
#define MMAD_EUNSUPPORTED -4   /* valid request this build has no kernel for */

/* Message for the calling thread's last non-zero return ("" if none). */
const char* mmad_last_error(void);
/* ABI version of this header (bumped on incompatible change). */
int mmad_abi_version(void);
/* Number of kernel launches issued through this library by the calling
 * process since load (bench.py reports it as gpu_launches). */
int64_t mmad_launch_count(void);
/* Tensor-core FLOPs actually ISSUED (2*M*N*16 per tcgen05.mma, counted by the issuing loops) by the convolution kernels
 * of this library on the current device since load; synchronises with the device.  Differs from the algorithmic FLOPs
 * of a layer by the zero-padding taps the kernels skip (fewer) and by partially filled 128-voxel tiles (more). */
int64_t mmad_executed_mma_flops(void);

/* ------------------------------------------------------------------ */
/* Part 1: atlas ROI pooling                                           */
/* ------------------------------------------------------------------ */

typedef struct mmad_roi_plan mmad_roi_plan;

/* Build the pooling plan of one atlas.  `labels_host`: n_voxels int32 on the
 * HOST, C-order (D,H,W) flattened, values in [0, n_rois]; 0 = background
 * (image_features.py:67-69).  n_rois = the atlas' max label, 1..255
 * (image_features.py:80-81 sizes the one-hot by max label, so labels with no
 * voxel still own an output column).  The plan run-length encodes the label
 * map into the per-tile run programme the kernel executes, uploads it, and
 * counts voxels per ROI on the GPU. */
int mmad_roi_plan_create(const int32_t* labels_host, int64_t n_voxels,
                         int32_t n_rois, mmad_roi_plan** plan_out);
int mmad_roi_plan_destroy(mmad_roi_plan* plan);

/* Per-ROI voxel counts (image_features.py:113 `den` before the clamp),
 * computed on the GPU at plan creation.  counts_host: int32[n_rois]. */
int mmad_roi_plan_counts(const mmad_roi_plan* plan, int32_t* counts_host);
/* Same counts as a device pointer owned by the plan (int32[n_rois]). */
int mmad_roi_plan_counts_dev(const mmad_roi_plan* plan, const int32_t** counts_dev);

/* Pool n_vols volumes.  vols_dev: n_vols x n_voxels float32, contiguous
 * (the reference's (B,C,D,H,W) tensor viewed as (B*C, D*H*W)); 4-byte
 * alignment suffices.  Outputs (device, any may be NULL to skip):
 *   mean_dev   float32[n_vols x n_rois]  sum / clamp_min(count, 1e-6)
 *   max_dev    float32[n_vols x n_rois]  max over member voxels (0 if empty)
 *   argmax_dev int32  [n_vols x n_rois]  flat voxel index of the first max
 *                                        (-1 if the ROI is empty)
 * With all three NULL only the streaming kernel runs (partials stay in the
 * plan's workspace); bench.py uses that to time the dominant kernel alone.
 * Asynchronous on `stream`.  A plan serves one in-flight call at a time. */
int mmad_roi_pool_f32(mmad_roi_plan* plan, const float* vols_dev, int64_t n_vols,
                      float* mean_dev, float* max_dev, int32_t* argmax_dev,
                      void* stream);

/* Same, HOST buffers in and out (what a caller holding numpy / pinned
 * tensors uses; bench.py's e2e leg).  Copies are chunked and overlapped with
 * the kernel on internal streams; returns after the results are on the host.
 * Pinned host memory gives full PCIe rate, pageable memory works. */
int mmad_roi_pool_host_f32(mmad_roi_plan* plan, const float* vols_host, int64_t n_vols,
                           float* mean_host, float* max_host, int32_t* argmax_host);

/* Backward of the mean (autograd of image_features.py:111-114):
 * grad_vols[n, v] = label[v] ? grad_mean[n, label[v]-1] / clamp_min(count,1e-6) : 0.
 * grad_mean_dev float32[n_vols x n_rois], grad_vols_dev float32[n_vols x n_voxels]. */
int mmad_roi_pool_mean_backward_f32(mmad_roi_plan* plan, const float* grad_mean_dev,
                                    int64_t n_vols, float* grad_vols_dev, void* stream);

/* Tuning / introspection entry points of the plan (the Python host code and the CPU test-suite use them).
 * mmad_roi_plan_create_ex: as mmad_roi_plan_create with an explicit tile (128/256/512 voxels), ring depth (0 = as many
 * stages as fit, <= 4), consumer warps (0 = 16; 8 or 16) and host_only != 0 to build the run programme without touching
 * a GPU.  mmad_roi_plan_programme copies the run programme (any pointer may be NULL; sizes come back through n_words /
 * n_tiles): words = per tile a 4-word header {records, 0, 0, 0} + records `label<<24 | start<<12 | len-1`, offs =
 * n_tiles+1 offsets in 16-byte units.  mmad_roi_plan_binding returns the work-item / partial-slot layout the plan uses
 * for n_vols volumes on a GPU with `sms` SMs (host arithmetic only; arrays may be NULL to query the sizes first). */
int mmad_roi_plan_create_ex(const int32_t* labels_host, int64_t n_voxels, int32_t n_rois, int32_t tile,
                            int32_t stages, int32_t consumer_warps, int32_t host_only, mmad_roi_plan** plan_out);
int mmad_roi_plan_programme(const mmad_roi_plan* plan, uint32_t* words, int64_t* n_words, int32_t* offs,
                            int32_t* n_tiles, int32_t* stages, int64_t* smem_bytes, int32_t* consumer_warps);
int mmad_roi_plan_binding(mmad_roi_plan* plan, int64_t n_vols, int32_t sms, int32_t* n_items, int32_t* n_slots,
                          int32_t* grid, int32_t* item_group, int32_t* item_t0, int32_t* item_t1,
                          int32_t* item_slot_ptr, uint8_t* slot_label, int32_t* slot_dst, int32_t* fin_ptr);

/* ROI means of a channels-last feature map that never left the GPU: feats_dev fp32 (n_vols, Dp, Hp, Wp, 64) NDHWC - the fp32
 * side output of mmad_conv3d_fwd_ex_bf16 for s_block1.conv2, the tensor image_features.py:58-60 hooks - pooled over the
 * plan's atlas (D, H, W) <= (Dp, Hp, Wp) (the crop of image_features.py:104 folded in).  mean_dev float32[n_vols][n_rois][64]
 * = image_features.py:114's roi_feat (B, R, C).  Only the rows of labelled voxels are read. */
int mmad_roi_pool_ndhwc_f32(mmad_roi_plan* plan, const float* feats_dev, int64_t n_vols, int Dp, int Hp, int Wp,
                            int D, int H, int W, int C, float* mean_dev, void* stream);

/* Algorithmic HBM bytes one mmad_roi_pool_f32 launch over n_vols volumes must
 * move (volume bytes + run programme + outputs); bench.py's roofline uses it. */
int64_t mmad_roi_pool_algorithmic_bytes(const mmad_roi_plan* plan, int64_t n_vols);

/* ------------------------------------------------------------------ */
/* Part 2: Conv3d / BatchNorm3d / ReLU stacks (NDHWC bf16 activations)  */
/* ------------------------------------------------------------------ */

/* y = conv3d(x, w): cubic kernel k, same stride / padding / dilation on the
 * three axes, no bias (resnet.py:14-23, :126-132, :188-194 all use
 * bias=False).  x: (N,D,H,W,Cin) bf16, w: [Cout][k*k*k][Cin] bf16 (tap index
 * = (kd*k + kh)*k + kw), y: (N,Do,Ho,Wo,Cout) bf16, all device, 16-byte
 * aligned.  Cin % 64 == 0; Cout in {64,128,256} or a multiple of 256.
 * If stats_partials != NULL it receives, per CTA, the per-channel sum and sum
 * of squares of the stored outputs: float[mmad_conv3d_stats_partials(...)]
 * [Cout][2] (inputs of mmad_bn_finalize).  Implicit GEMM on tcgen05/TMEM with
 * TMA-staged operands; dgrad of a stride-1 convolution is the same call on the
 * flipped / transposed weights (mmad_conv3d_prep_weights). */
int mmad_conv3d_fwd_bf16(const void* x, const void* w, void* y, float* stats_partials,
                         int N, int D, int H, int W, int Cin, int Cout,
                         int k, int stride, int pad, int dil, void* stream);
int mmad_conv3d_stats_partials(int N, int D, int H, int W, int Cout,
                               int k, int stride, int pad, int dil);

/* Weight gradient of the same convolution (the dW half of Conv3d.backward):
 * partials[split][Cout][taps][Cin] fp32 = split-K partial sums over output
 * voxels of dy[v][co] * x[v*stride + tap*dil - pad][ci]; x (N,D,H,W,Cin) and
 * dy (N,Do,Ho,Wo,Cout) bf16 NDHWC.  mmad_conv3d_wgrad_workspace returns the
 * fp32 element count of `partials` (and the split count); mmad_wgrad_reduce
 * sums the splits into torch's (Cout, Cin, k,k,k) fp32 layout.
 * Cin in {64} or a multiple of 128; Cout in {64,128} or a multiple of 256. */
int64_t mmad_conv3d_wgrad_workspace(int N, int D, int H, int W, int Cin, int Cout,
                                    int k, int stride, int pad, int dil, int* nsplit_out);
int mmad_conv3d_wgrad_bf16(const void* x, const void* dy, float* partials,
                           int N, int D, int H, int W, int Cin, int Cout,
                           int k, int stride, int pad, int dil, void* stream);
int mmad_wgrad_reduce(const float* partials, int nsplit, float* dw,
                      int Cout, int Cin, int taps, void* stream);

/* torch weights (Cout, Cin, k,k,k) fp32 -> forward layout [Cout][taps][Cin]
 * bf16 (w_fwd) and dgrad layout [Cin][taps reversed][Cout] bf16 (w_dgrad);
 * either output may be NULL. */
int mmad_conv3d_prep_weights(const float* w, void* w_fwd, void* w_dgrad,
                             int Cout, int Cin, int taps, void* stream);

/* The same for n layers at once (two launches per 64 layers instead of two per layer: a training step re-lays every
 * convolution's weights).  w / w_fwd / w_dgrad: HOST arrays of n DEVICE pointers (w_dgrad[i] NULL = no dgrad layout for layer
 * i); cout / cin / taps: host arrays of n ints. */
int mmad_conv3d_prep_weights_batched(int n, const void* const* w, void* const* w_fwd, void* const* w_dgrad,
                                     const int* cout, const int* cin, const int* taps, void* stream);

/* Data gradient of a 3x3x3, stride-2, padding-1 convolution (resnet.py:18-23 with
 * stride 2: layer2.0.conv1) without zero insertion: dx (N,D,H,W,Cdx) bf16 from
 * dy (N,(D-1)/2+1,..,Cdy) bf16 as 8 interleaved stride-1 phase convolutions.
 * w_phases (27*Cdx*Cdy bf16) comes from mmad_conv3d_prep_weights_s2 on the
 * torch weight (Cdy, Cdx, 3,3,3) fp32.  Cdx: 64/128/256/k*256, Cdy % 64 == 0. */
int mmad_conv3d_prep_weights_s2(const float* w, void* w_phases, int Cdy, int Cdx, void* stream);
int mmad_conv3d_dgrad_s2_bf16(const void* dy, const void* w_phases, void* dx,
                              int N, int D, int H, int W, int Cdx, int Cdy, void* stream);

/* Stem without a materialised im2col matrix (the path the model uses).  The
 * stride-2 7^3 convolution of one channel is rewritten as a stride-1 4^3
 * convolution of the 8 space-to-depth phase channels (K = 512):
 *   xs bf16 [N][Do+3][Ho+3][Wo+3][8],  xs[..][jd][jh][jw][pd*4+ph*2+pw] =
 *   x[2jd+pd-3][2jh+ph-3][2jw+pw-3]   (mmad_stem_s2d_elems elements),
 *   wk bf16 [64][512], K = ((kd*4+kh)*4+kw)*8 + phase, zero where a tap is 7.
 * mmad_stem_s2d_fwd writes y (N,Do,Ho,Wo,64) bf16 and, unless NULL, the same
 * per-CTA BatchNorm partials as mmad_conv3d_fwd_bf16
 * ([mmad_stem_s2d_stats_partials][64][2]).  mmad_stem_s2d_wgrad replaces the
 * weight-gradient half of conv1's backward: fp32 partials
 * [nsplit][64][512] (mmad_stem_s2d_wgrad_workspace), summed into the torch
 * layout (64,1,7,7,7) by mmad_stem_s2d_wgrad_reduce. */
int64_t mmad_stem_s2d_elems(int N, int D, int H, int W);
int mmad_stem_s2d_pack(const float* x, void* xs, int N, int D, int H, int W, void* stream);
int mmad_stem_s2d_prep_weights(const float* w, void* wk, void* stream);
int mmad_stem_s2d_stats_partials(int N, int D, int H, int W);
int mmad_stem_s2d_fwd(const void* xs, const void* wk, void* y, float* stats_partials,
                      int N, int D, int H, int W, void* stream);
int64_t mmad_stem_s2d_wgrad_workspace(int N, int D, int H, int W, int* nsplit_out);
int mmad_stem_s2d_wgrad(const void* xs, const void* dy, float* partials,
                        int N, int D, int H, int W, void* stream);
int mmad_stem_s2d_wgrad_reduce(const float* partials, int nsplit, float* dw, void* stream);

/* ---- UNet3D pieces (/root/reference/models/unet3d.py) --------------------------------------------------------------- */

/* mmad_conv3d_fwd_bf16 with (a) output rows `ldy` elements apart (0 = dense): y may be a channel slice of a wider NDHWC
 * tensor, so the producers of a concatenation (unet3d.py:77 torch.cat((upconv, residual), 1)) write it in place; (b) a
 * per-channel epilogue on the fp32 accumulators, stored = act(acc * ep_scale[c] + ep_shift[c]) (NULL scale = 1, NULL shift =
 * 0, ep_relu != 0 = ReLU): a convolution bias (unet3d.py:37-40), or in eval mode the BatchNorm3d + ReLU that follows the
 * convolution folded into the producing kernel; (c) an fp32 side output out_f32[voxel][Cout] = acc + f32_bias[c] (NULL bias
 * = 0), dense NDHWC: the raw convolution output image_features.py:58-60 hooks (s_block1.conv2), kept at accumulator
 * precision for the ROI pooling.  stats_partials are the statistics of the STORED bf16 values.  (d) cin_tensor (0 = Cin): the
 * input tensor has only cin_tensor < Cin = 64 channels per voxel (unet3d.py's 32-channel a_block1.conv1 output): x is
 * (N,D,H,W,cin_tensor), the weights keep Cin = 64 with zero columns, and TMA zero-fills the rest of every 128-byte K slice in
 * flight - the padding is neither stored nor fetched. */
int mmad_conv3d_fwd_ex_bf16(const void* x, const void* w, void* y, int64_t ldy, float* stats_partials,
                            const float* ep_scale, const float* ep_shift, int ep_relu, float* out_f32, const float* f32_bias,
                            int N, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dil,
                            int cin_tensor, void* stream);
/* The tail of unet3d.py's eval-mode forward as one kernel: 3x3x3 convolution (padding 1) to 64 channels, the BatchNorm3d + ReLU
 * that follow it (ep_scale / ep_shift), the 1x1x1 head (unet3d.py:72 conv3: head_w float[K][64], head_b float[K], K <= 8) and the
 * crop back to (Dc,Hc,Wc) (unet3d.py:126-135) - the 64-channel activation is never written.  head_out fp32 (N,K,Dc,Hc,Wc);
 * out_f32 (optional) receives the raw convolution output + f32_bias as fp32 NDHWC (N,D,H,W,64), the tensor
 * image_features.py:58-60 hooks. */
int mmad_conv3d_fwd_head_bf16(const void* x, const void* w, const float* ep_scale, const float* ep_shift,
                              float* out_f32, const float* f32_bias, const float* head_w, const float* head_b, float* head_out,
                              int K, int N, int D, int H, int W, int Dc, int Hc, int Wc, int Cin, void* stream);
/* mmad_conv3d_wgrad_bf16 for an x that is the channel slice [c0, c0+Cin) of a wider NDHWC tensor (rows ldx elements apart, x
 * points at channel c0), and the matching reduction into dw[:, ci_off : ci_off+Cin] of a (Cout, Cin_total, k,k,k) tensor:
 * the weight gradient of a convolution over a concatenated input, computed per source. */
int mmad_conv3d_wgrad_ex_bf16(const void* x, int64_t ldx, const void* dy, float* partials,
                              int N, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dil, void* stream);
int mmad_wgrad_reduce_ex(const float* partials, int nsplit, float* dw, int Cout, int Cin, int taps, int Cin_total, int ci_off,
                         void* stream);

/* ConvTranspose3d(kernel 2, stride 2) (unet3d.py:68,75): y[n][2v+p][co] = bias[co] + sum_ci x[n][v][ci] * w[ci][co][p] as eight
 * 1x1x1 implicit GEMMs writing interleaved through strided tensor maps, rows ldy elements apart (0 = dense).  x (N,D,H,W,Cin),
 * y (N,2D,2H,2W,[ldy]) bf16; w_phases [8][Cout][Cin] bf16 from mmad_convtranspose3d_prep_weights on the torch weight
 * (Cin, Cout, 2,2,2) fp32.  Its data gradient is mmad_conv3d_fwd_bf16 with k = 2, stride 2, pad 0 on
 * mmad_conv3d_prep_weights(w viewed as (Cout' = Cin, Cin' = Cout, 8 taps)); its weight gradient is mmad_conv3d_wgrad_bf16 with
 * x := dy (fine grid), dy := x (coarse grid), k = 2, stride 2, pad 0 - the result is already in the (Cin, Cout, 2,2,2) layout. */
int mmad_convtranspose3d_prep_weights(const float* w, void* w_phases, int Cin, int Cout, void* stream);
int mmad_convtranspose3d_k2s2_fwd_bf16(const void* x, const void* w_phases, const float* bias, void* y, int64_t ldy,
                                       int N, int D, int H, int W, int Cin, int Cout, void* stream);

/* First UNet layer (unet3d.py:37 a_block1.conv1 = Conv3d(1, 32, 3, padding 1) after the zero extension to 96x112x96 of
 * unet3d.py:116-123): direct convolution.  x fp32 (N,1,D,H,W); w fp32 (32,1,3,3,3); y bf16 (N,Do,Ho,Wo,64), (Do,Ho,Wo) >= (D,H,W),
 * channels 32..63 written as zeros (the next convolution reads 64-channel rows); no bias (add it through the BatchNorm
 * shift).  Training: stats_partials float[mmad_conv3d_c1_blocks][64][2] like mmad_conv3d_fwd_bf16 (ep_* NULL).  Eval:
 * ep_scale / ep_shift float[32] fold BatchNorm3d + ReLU (and the bias) in, stored = relu(acc * scale[c] + shift[c])
 * (stats_partials NULL).  out_channels = 64: y (N,Do,Ho,Wo,64) as above; 32: y (N,Do,Ho,Wo,32), no padding channels stored - the
 * next convolution then reads it with cin_tensor = 32 (mmad_conv3d_fwd_ex_bf16).  Weight gradient: partials float[mmad_conv3d_c1_wgrad_blocks][32][27], summed by
 * mmad_wgrad_reduce(partials, blocks, dw, 32, 1, 27). */
int mmad_conv3d_c1_blocks(int N, int Do, int Ho, int Wo);
int mmad_conv3d_c1_wgrad_blocks(int N, int Do, int Ho, int Wo);
int mmad_conv3d_c1_fwd(const float* x, const float* w, void* y, float* stats_partials, const float* ep_scale, const float* ep_shift,
                       int N, int D, int H, int W, int Do, int Ho, int Wo, int out_channels, void* stream);
int mmad_conv3d_c1_wgrad(const float* x, const void* dy, float* partials,
                         int N, int D, int H, int W, int Do, int Ho, int Wo, void* stream);

/* MaxPool3d(kernel 2, stride 2) (unet3d.py:31,44; floor mode), NDHWC bf16.  Forward reads x with rows ldx elements apart
 * (0 = dense; the skip tensor is pooled in place from the concatenation buffer), writes y dense (N,D/2,H/2,W/2,C) and idx
 * (uint8 per element: winning window position, first maximum in scan order).  Backward writes dx dense (N,D,H,W,C). */
int mmad_maxpool3d_k2_fwd(const void* x, int64_t ldx, void* y, void* idx, int N, int D, int H, int W, int C, void* stream);
int mmad_maxpool3d_k2_bwd(const void* dy, const void* idx, void* dx, int N, int D, int H, int W, int C, void* stream);

/* Head (unet3d.py:72 conv3 = Conv3d(64, K, 1) with bias, fused with the crop back of unet3d.py:126-135): x bf16
 * (N,Dp,Hp,Wp,64) -> out fp32 (N,K,D,H,W) for the (D,H,W) <= (Dp,Hp,Wp) corner; K <= 8.  Backward: dx bf16 (N,Dp,Hp,Wp,64)
 * (zero in the padded margin), dw float[K][64], db float[K]; partials: float[mmad_head1x1_bwd_blocks()][K][65] scratch. */
int mmad_head1x1_fwd(const void* x, const float* w, const float* bias, float* out,
                     int N, int Dp, int Hp, int Wp, int D, int H, int W, int C, int K, void* stream);
int mmad_head1x1_bwd_blocks(void);
int mmad_head1x1_bwd(const void* x, const float* w, const float* dout, void* dx, float* partials, float* dw, float* db,
                     int N, int Dp, int Hp, int Wp, int D, int H, int W, int C, int K, void* stream);

/* mmad_bn_apply with the bf16 output's rows out_ld elements apart (0 = dense): the activation is written straight into its
 * channel slice of a concatenation buffer. */
int mmad_bn_apply_ex(const void* x, const float* scale, const float* shift,
                     const void* res, const float* rscale, const float* rshift, int relu,
                     void* out_bf16, int64_t out_ld, float* out_f32, int64_t rows, int C, void* stream);

/* BatchNorm3d (resnet.py:46,49,134), training statistics from the conv
 * epilogue's partials: mean, invstd, scale = gamma*invstd, shift = beta -
 * mean*scale (all float[C]); running_mean/var updated like nn.BatchNorm3d
 * (momentum, unbiased variance) unless NULL.  `count` = voxels per channel. */
int mmad_bn_finalize(const float* partials, int nparts, int C, double count,
                     const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var,
                     float* mean, float* invstd, float* scale, float* shift, void* stream);
/* eval mode: the same four vectors from the running statistics */
int mmad_bn_eval_params(int C, const float* gamma, const float* beta,
                        const float* running_mean, const float* running_var, float eps,
                        float* mean, float* invstd, float* scale, float* shift, void* stream);
/* y = act(x*scale + shift [+ res*rscale + rshift | + res]); x, res: rows x C
 * bf16; relu != 0 applies ReLU (resnet.py:57-67); out_bf16 and/or out_f32. */
int mmad_bn_apply(const void* x, const float* scale, const float* shift,
                  const void* res, const float* rscale, const float* rshift, int relu,
                  void* out_bf16, float* out_f32, int64_t rows, int C, void* stream);
/* Backward of ReLU + BatchNorm3d.  reduce: g = (dy [+ dy2]) * (mask > 0),
 * written to g_out (bf16, may be NULL) with per-block partial sums of g and
 * g*xhat; finalize: dgamma, dbeta and coef = float[3][C] with
 * dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) = coef0*g + coef1*x + coef2
 * (training == 0: eval-mode BatchNorm is affine, dx = gamma*invstd*g); apply:
 * that map.  dy is bf16 (dy_bf16) or fp32 (dy_f32), both rows x C in NDHWC
 * order; mask NULL = no ReLU, unless mask_scale/mask_shift (float[C]) are given:
 * then the mask is recomputed as relu(x*mask_scale + mask_shift) > 0. */
int mmad_bn_bwd_partials(int64_t rows);
int mmad_bn_bwd_reduce(const void* dy_bf16, const float* dy_f32, const void* dy2,
                       const void* mask, const void* x, const float* mean, const float* invstd,
                       const float* mask_scale, const float* mask_shift,
                       void* g_out, float* partials, int64_t rows, int C, void* stream);
int mmad_bn_bwd_finalize(const float* partials, int nparts, int C, double count,
                         const float* gamma, const float* mean, const float* invstd, int training,
                         float* dgamma, float* dbeta, float* coef, void* stream);
int mmad_bn_bwd_apply(const void* g, const void* x, const float* coef,
                      void* dx, int64_t rows, int C, void* stream);
/* The same with the ReLU mask recomputed in THIS pass: g is the unmasked upstream gradient (bf16), the mask is
 * relu(x*mask_scale + mask_shift) > 0.  With mmad_bn_bwd_reduce(g_out = NULL, mask_scale / mask_shift given) a BatchNorm + ReLU
 * backward costs four tensor passes instead of six: no masked gradient is written, no stored activation is read as the mask. */
int mmad_bn_bwd_apply_ex(const void* g, const void* x, const float* coef, const float* mask_scale, const float* mask_shift,
                         void* dx, int64_t rows, int C, void* stream);

/* MaxPool3d(kernel 3, stride 2, padding 1) (resnet.py:136), NDHWC bf16; idx
 * keeps the winning tap per element (uint8, rows x C) for the backward. */
int mmad_maxpool3d_fwd(const void* x, void* y, void* idx, int N, int D, int H, int W, int C, void* stream);
int mmad_maxpool3d_bwd(const void* dy, const void* idx, void* dx, int N, int D, int H, int W, int C, void* stream);

/* Fused stem forward (resnet.py:205-208 after conv1): p = maxpool3d(relu(bn(c)), 3, 2, 1)
 * in one pass over the conv output c (N,D,H,W,C) bf16; the post-ReLU tensor is
 * never stored (its ReLU mask is recomputed in the backward from c, see
 * mask_scale / mask_shift of mmad_bn_bwd_reduce). */
int mmad_stem_bn_relu_maxpool_fwd(const void* c, const float* scale, const float* shift,
                                  void* y, void* idx, int N, int D, int H, int W, int C, void* stream);

/* y[2*o] = x[o], zero elsewhere: dgrad of a stride-2 convolution is the
 * unit-stride convolution of this with the flipped kernel. */
int mmad_upsample_zero2(const void* x, void* y, int N, int Dx, int Hx, int Wx,
                        int Dy, int Hy, int Wy, int C, void* stream);
/* (N, C, S) fp32 [torch NCDHW] -> (N, S, C) bf16 [NDHWC] */
int mmad_ncs_f32_to_nsc_bf16(const float* x, void* y, int N, int C, int64_t S, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMAD_B200_H */
