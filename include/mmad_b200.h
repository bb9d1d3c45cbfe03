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
 *   /root/reference/models/resnet.py:14-23,40-69,112-215 (and resnet18.py,
 *   ImageEncoder.py, unet3d.py which repeat them), forward and backward.
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

#ifdef __cplusplus
}
#endif
#endif /* MMAD_B200_H */
