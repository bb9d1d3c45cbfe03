// UNet3D companions of the Conv3d implicit GEMM (sm_100a): the pieces of /root/reference/models/unet3d.py that are not a
// 64-channel-aligned convolution.
//   * a_block1.conv1 (unet3d.py:37, Conv3d(1, 32, 3, padding 1) on the volume zero-extended to 96x112x96, unet3d.py:116-123):
//     direct convolution on CUDA cores (K = 27: nothing for a tensor core to do), forward + weight gradient
//   * MaxPool3d(2, 2) forward / backward (unet3d.py:44), reading its input in place from a concatenation buffer
//   * ConvTranspose3d(2, 2) weight re-layout for the phase GEMMs of mmad_convtranspose3d_k2s2_fwd_bf16 (unet3d.py:68)
//   * s_block1.conv3 (unet3d.py:72, Conv3d(64, num_classes, 1)) fused with the crop back to the input size (unet3d.py:126-135),
//     forward + backward
// All activations NDHWC bf16.  HBM-bound kernels: 16-byte vectors per thread, grids sized in multiples of the SM count.
#include "common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace mmad {

__device__ __forceinline__ void u_unpack8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 u_pack8(const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// First layer: y[n][d][h][w][co] = sum_taps x_ext[n][d+a-1][h+b-1][w+c-1] * wgt[co][tap], co < 32; channels 32..63 of the
// 64-channel output row are written as zeros (the next convolution's K slices are 64 channels wide).  x_ext = the fp32 input
// (N,1,D,H,W) zero-extended to the output grid (Do,Ho,Wo) >= (D,H,W) (F.pad on the right, unet3d.py:121-122) and zero padded by
// one voxel for the 3x3x3 kernel.  A block owns 8 x 8 x 4 output voxels (one per thread): the 10 x 10 x 6 input halo and the
// 27 x 32 weights are staged in shared memory; per-channel sum / sum of squares of the STORED bf16 values go to
// stats_partials[block][64][2] (BatchNorm statistics, the same contract as mmad_conv3d_fwd_bf16).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kC1Tw = 8, kC1Th = 8, kC1Td = 4;            // weight-gradient tile (one voxel per thread)
constexpr int kF1Tw = 16, kF1Th = 8, kF1Td = 4;           // forward tile: 512 voxels, two W-neighbours per thread
constexpr int kF1Sx = (kF1Td + 2) * (kF1Th + 2) * (kF1Tw + 2);

// STATS: per-channel sum / sum of squares of the stored values (training).  EPI: stored = relu(acc * scale[c] + shift[c]) - in
// eval mode the BatchNorm3d + ReLU that follow the convolution (and its bias) are folded in, the pre-activation is never written.
// A thread computes 8 W-consecutive voxels x 8 channels (one 16-byte chunk of the output row): the four lanes of a voxel group
// write 64 contiguous bytes per store (full sectors - one voxel's 32 channels), every broadcast weight vector feeds 16 FMAs and
// every staged input value 24, so the kernel is bound by the FP32 pipe (864 FMAs per voxel) and by the 128-byte-per-voxel
// output stream, not by shared-memory traffic or partial-sector writes.
template <bool STATS, bool EPI>
__global__ void __launch_bounds__(256) conv3d_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wgt, uint4* __restrict__ y,
                                                            float* __restrict__ stats_partials, const float* __restrict__ ep_scale,
                                                            const float* __restrict__ ep_shift, int N, int D, int H, int W, int Do, int Ho,
                                                            int Wo, int tiles_w, int tiles_h, int tiles_d, int out_c8) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float sw[27 * 32];
    __shared__ float sx[kF1Sx];
    __shared__ float sred[64 * 2];
    __shared__ __align__(16) float sep[64];                 // scale[32], shift[32]
    for (int i = threadIdx.x; i < 27 * 32; i += 256) sw[i] = wgt[(i & 31) * 27 + (i >> 5)];      // [tap][co]
    if (threadIdx.x < 128) sred[threadIdx.x] = 0.f;
    if (EPI && threadIdx.x < 64) sep[threadIdx.x] = threadIdx.x < 32 ? ep_scale[threadIdx.x] : ep_shift[threadIdx.x - 32];
    const int cq = threadIdx.x & 3;                         // channel chunk: channels 8 cq .. 8 cq + 7
    const int vg = threadIdx.x >> 2;                        // voxel group: 8 consecutive voxels along W
    const int wg = vg & 1, th = (vg >> 1) & 7, td = vg >> 4;
    float ssum[STATS ? 8 : 1], ssq[STATS ? 8 : 1];
#pragma unroll
    for (int c = 0; c < (STATS ? 8 : 1); ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
    const long long total = (long long)N * tiles_d * tiles_h * tiles_w;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        long long r = t;
        const int wt = (int)(r % tiles_w); r /= tiles_w;
        const int ht = (int)(r % tiles_h); r /= tiles_h;
        const int dt = (int)(r % tiles_d); r /= tiles_d;
        const int n = (int)r;
        const int w0 = wt * kF1Tw, h0 = ht * kF1Th, d0 = dt * kF1Td;
        __syncthreads();                                    // previous tile's readers are done with sx (and sw / sep are complete)
        for (int i = threadIdx.x; i < kF1Sx; i += 256) {
            const int iw = w0 - 1 + i % (kF1Tw + 2), ih = h0 - 1 + (i / (kF1Tw + 2)) % (kF1Th + 2), id = d0 - 1 + i / ((kF1Tw + 2) * (kF1Th + 2));
            sx[i] = ((unsigned)iw < (unsigned)W && (unsigned)ih < (unsigned)H && (unsigned)id < (unsigned)D)
                        ? __ldg(x + (((long long)n * D + id) * H + ih) * W + iw) : 0.f;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int v = 0; v < 8; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const float* row = sx + ((td + a) * (kF1Th + 2) + th + b) * (kF1Tw + 2) + wg * 8;
                float xr[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) xr[k] = row[k];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4* wp = reinterpret_cast<const float4*>(sw + ((a * 3 + b) * 3 + c) * 32 + cq * 8);
                    const float4 wa = wp[0], wb = wp[1];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float xv = xr[v + c];
                        acc[v][0] = fmaf(xv, wa.x, acc[v][0]); acc[v][1] = fmaf(xv, wa.y, acc[v][1]);
                        acc[v][2] = fmaf(xv, wa.z, acc[v][2]); acc[v][3] = fmaf(xv, wa.w, acc[v][3]);
                        acc[v][4] = fmaf(xv, wb.x, acc[v][4]); acc[v][5] = fmaf(xv, wb.y, acc[v][5]);
                        acc[v][6] = fmaf(xv, wb.z, acc[v][6]); acc[v][7] = fmaf(xv, wb.w, acc[v][7]);
                    }
                }
            }
        const int oh = h0 + th, od = d0 + td;
        if (oh < Ho && od < Do) {
            uint4* line = y + (((long long)n * Do + od) * Ho + oh) * Wo * out_c8;   // out_c8 = 8: 64-channel rows (upper half zero), 4: 32-channel rows
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const int ow = w0 + wg * 8 + v;
                if (ow < Wo) {
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[j] = acc[v][j];
                        if (EPI) f[j] = fmaxf(fmaf(f[j], sep[cq * 8 + j], sep[32 + cq * 8 + j]), 0.f);
                    }
                    const uint4 pk = u_pack8(f);
                    line[(long long)ow * out_c8 + cq] = pk;
                    if (out_c8 == 8) line[(long long)ow * 8 + 4 + cq] = make_uint4(0u, 0u, 0u, 0u);   // padding channels 32..63
                    if (STATS) {
                        u_unpack8(pk, f);                   // statistics of the rounded values, as the tensor-core epilogue does
#pragma unroll
                        for (int j = 0; j < 8; ++j) { ssum[j] += f[j]; ssq[j] += f[j] * f[j]; }
                    }
                }
            }
        }
    }
    if (STATS && stats_partials) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = ssum[j], b = ssq[j];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if ((threadIdx.x & 31) < 4) { atomicAdd(&sred[2 * (cq * 8 + j)], a); atomicAdd(&sred[2 * (cq * 8 + j) + 1], b); }
        }
        __syncthreads();
        if (threadIdx.x < 128) stats_partials[(size_t)blockIdx.x * 128 + threadIdx.x] = threadIdx.x < 64 ? sred[threadIdx.x] : 0.f;
    }
}

// Weight gradient of the first layer: dW[co][tap] = sum_v dy[v][co] * x_ext[v + tap - 1], co < 32 (dy: 64-channel rows, the upper
// 32 channels are ignored).  8 x 8 x 4 voxel tiles staged in shared memory (input halo + the 32 dy channels as fp32).  Thread =
// (4 output channels, 7 taps, one of 8 voxel slices): per voxel one 16-byte load of dy and 7 input loads feed 28 FMAs, all 32
// lanes of a warp walk the same voxels (the warp IS the voxel slice), so every load is a broadcast / conflict-free.  The 28
// accumulators live in registers across all tiles of the block; the 8 slices are summed through shared memory at the end and the
// block writes partials[block][co][tap] (summed over blocks by mmad_wgrad_reduce).
__global__ void __launch_bounds__(256) conv3d_c1_wgrad_kernel(const float* __restrict__ x, const uint4* __restrict__ dy, float* __restrict__ partials,
                                                              int N, int D, int H, int W, int Do, int Ho, int Wo, int tiles_w, int tiles_h,
                                                              int tiles_d) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sx[6 * 10 * 10];
    __shared__ __align__(16) float sdy[256 * 36];           // [voxel][co], rows 36 floats apart (16-byte aligned, bank-staggered)
    __shared__ float sred[32 * 28];
    const int cg = threadIdx.x & 7, tg = (threadIdx.x >> 3) & 3, vs = threadIdx.x >> 5;
    int toff[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int tap = min(tg * 7 + j, 26);                 // tap 27 (tg == 3, j == 6) does not exist: computed on tap 26, never stored
        toff[j] = ((tap / 9) * 10 + (tap / 3) % 3) * 10 + tap % 3;
    }
    float acc[4][7];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[c][j] = 0.f;
    for (int i = threadIdx.x; i < 32 * 28; i += 256) sred[i] = 0.f;
    const long long total = (long long)N * tiles_d * tiles_h * tiles_w;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        long long r = t;
        const int wt = (int)(r % tiles_w); r /= tiles_w;
        const int ht = (int)(r % tiles_h); r /= tiles_h;
        const int dt = (int)(r % tiles_d); r /= tiles_d;
        const int n = (int)r;
        const int w0 = wt * kC1Tw, h0 = ht * kC1Th, d0 = dt * kC1Td;
        __syncthreads();
        for (int i = threadIdx.x; i < 600; i += 256) {
            const int iw = w0 - 1 + i % 10, ih = h0 - 1 + (i / 10) % 10, id = d0 - 1 + i / 100;
            sx[i] = ((unsigned)iw < (unsigned)W && (unsigned)ih < (unsigned)H && (unsigned)id < (unsigned)D)
                        ? __ldg(x + (((long long)n * D + id) * H + ih) * W + iw) : 0.f;
        }
        // dy tile: 256 voxels x 32 channels = 4 vectors per voxel; thread i loads vector (i & 3) of voxel (i >> 2) + 64 * pass
        for (int pass = 0; pass < 4; ++pass) {
            const int v = (threadIdx.x >> 2) + 64 * pass, q = threadIdx.x & 3;
            const int ow = w0 + (v & 7), oh = h0 + ((v >> 3) & 7), od = d0 + (v >> 6);
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ow < Wo && oh < Ho && od < Do) u_unpack8(__ldg(dy + ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * 8 + q), f);
            float4* dst = reinterpret_cast<float4*>(sdy + v * 36 + 8 * q);
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int v = vs * 32 + k;
            const float4 g = *reinterpret_cast<const float4*>(sdy + v * 36 + cg * 4);
            const int base = ((v >> 6) * 10 + ((v >> 3) & 7)) * 10 + (v & 7);
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float xv = sx[base + toff[j]];
                acc[0][j] = fmaf(g.x, xv, acc[0][j]); acc[1][j] = fmaf(g.y, xv, acc[1][j]);
                acc[2][j] = fmaf(g.z, xv, acc[2][j]); acc[3][j] = fmaf(g.w, xv, acc[3][j]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 7; ++j)
            atomicAdd(&sred[(cg * 4 + c) * 28 + tg * 7 + j], acc[c][j]);     // 8 voxel slices (warps) meet here, once per kernel
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 27; i += 256) partials[(size_t)blockIdx.x * 32 * 27 + i] = sred[(i / 27) * 28 + i % 27];
}

// ---------------------------------------------------------------------------------------------------------------------
// MaxPool3d(kernel 2, stride 2) (unet3d.py:31,44; floor mode).  x: (N,D,H,W,C) bf16 whose rows are ldx elements apart (the skip
// half of a concatenation buffer is pooled in place), y dense (N,D/2,H/2,W/2,C), idx uint8 = winning window element
// ((a*2+b)*2+c, first maximum in scan order like torch).  Thread = (output voxel, 8-channel vector).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const uint4* __restrict__ x, int ldx8, uint4* __restrict__ y, uint2* __restrict__ idx,
                                                           int D, int H, int W, int cv, int Do, int Ho, int Wo, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int v = (int)(r % cv); r /= cv;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho); r /= Ho;
        const int od = (int)(r % Do); r /= Do;
        const long long n = r;
        float best[8];
        unsigned char bi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int a = k >> 2, b = (k >> 1) & 1, c = k & 1;
            const long long vox = ((n * D + 2 * od + a) * H + 2 * oh + b) * W + 2 * ow + c;
            float f[8];
            u_unpack8(__ldg(x + vox * ldx8 + v), f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k == 0 || f[j] > best[j]) { best[j] = f[j]; bi[j] = (unsigned char)k; }
        }
        y[i] = u_pack8(best);
        if (idx) {
            uint2 o;
            o.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | ((unsigned)bi[3] << 24);
            o.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | ((unsigned)bi[7] << 24);
            idx[i] = o;
        }
    }
}

// dx (N,D,H,W,C) dense bf16: every window element receives dy if it won, else 0 (windows do not overlap: a pure scatter,
// written as 8 full vectors per thread).  Voxels of an odd tail (not covered by any window) must be pre-zeroed by the caller.
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx, uint4* __restrict__ dx,
                                                           int D, int H, int W, int cv, int Do, int Ho, int Wo, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int v = (int)(r % cv); r /= cv;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho); r /= Ho;
        const int od = (int)(r % Do); r /= Do;
        const long long n = r;
        const uint4 g = __ldg(dy + i);
        const uint2 ix = __ldg(idx + i);
        const unsigned short* gs = reinterpret_cast<const unsigned short*>(&g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int a = k >> 2, b = (k >> 1) & 1, c = k & 1;
            unsigned short o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned w = ((j < 4 ? ix.x : ix.y) >> (8 * (j & 3))) & 0xffu;
                o[j] = w == (unsigned)k ? gs[j] : (unsigned short)0;
            }
            uint4 ov;
            ov.x = o[0] | ((unsigned)o[1] << 16); ov.y = o[2] | ((unsigned)o[3] << 16);
            ov.z = o[4] | ((unsigned)o[5] << 16); ov.w = o[6] | ((unsigned)o[7] << 16);
            dx[(((n * D + 2 * od + a) * H + 2 * oh + b) * W + 2 * ow + c) * cv + v] = ov;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// ConvTranspose3d(2, 2) weights: torch (Cin, Cout, 2,2,2) fp32 -> [8 phases][Cout][Cin] bf16 (phase p = (pd*2+ph)*2+pw).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_convtranspose_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cin,
                                                                         int Cout) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = 8ll * Cin * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout);
        const int p = (int)(i / ((long long)Cin * Cout));
        wp[i] = __float2bfloat16_rn(w[((long long)ci * Cout + co) * 8 + p]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Head: out[n][k][d][h][w] = bias[k] + sum_c x[n][d][h][w][c] * wgt[k][c] for (d,h,w) inside the crop (D,H,W) of the padded grid
// (Dp,Hp,Wp) (unet3d.py:72 conv3 + :126-135 _crop_back).  C = 64: eight lanes per voxel (one 16-byte vector each), a warp covers
// four consecutive voxels; K (num_classes) <= 8.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kHeadMaxK = 8;

__global__ void __launch_bounds__(256) head_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                                                       float* __restrict__ out, int N, int Dp, int Hp, int Wp, int D, int H, int W, int K) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sw[kHeadMaxK * 64];
    for (int i = threadIdx.x; i < K * 64; i += 256) sw[i] = wgt[i];
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const long long total = (long long)N * D * H * W;
    for (long long v = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; v < ((total + 31) & ~31ll);
         v += ((long long)gridDim.x * blockDim.x) >> 3) {
        const bool ok = v < total;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        long long r = ok ? v : 0;
        const int w_ = (int)(r % W); r /= W;
        const int h_ = (int)(r % H); r /= H;
        const int d_ = (int)(r % D); r /= D;
        const long long n = r;
        if (ok) u_unpack8(__ldg(x + (((n * Dp + d_) * Hp + h_) * Wp + w_) * 8 + sub), f);
        for (int k = 0; k < K; ++k) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(f[j], sw[k * 64 + sub * 8 + j], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (ok && sub == 0) out[((n * K + k) * D + d_) * (long long)H * W + (long long)h_ * W + w_] = s + bias[k];
        }
    }
}

// Backward of the head: dx[n][d][h][w][c] = sum_k dout[n][k][d][h][w] * wgt[k][c] inside the crop, 0 in the padded margin
// (the gradient of a crop is a zero pad); dW[k][c] = sum_v dout[k][v] * x[v][c] and db[k] = sum_v dout[k][v] as per-block
// partials [block][K][65] (column 64 = db), summed by head_bwd_reduce_kernel.
__global__ void __launch_bounds__(256) head_bwd_kernel(const uint4* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ dout,
                                                       uint4* __restrict__ dx, float* __restrict__ partials, int N, int Dp, int Hp, int Wp, int D,
                                                       int H, int W, int K) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sw[kHeadMaxK * 64];
    __shared__ float sred[kHeadMaxK * 65];
    for (int i = threadIdx.x; i < K * 64; i += 256) sw[i] = wgt[i];
    for (int i = threadIdx.x; i < K * 65; i += 256) sred[i] = 0.f;
    __syncthreads();
    const int sub = threadIdx.x & 7;
    float aw[kHeadMaxK][8], ab[kHeadMaxK];
#pragma unroll
    for (int k = 0; k < kHeadMaxK; ++k) {
        ab[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) aw[k][j] = 0.f;
    }
    const long long total = (long long)N * Dp * Hp * Wp;     // every voxel of the PADDED grid gets a gradient row
    for (long long v = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; v < total; v += ((long long)gridDim.x * blockDim.x) >> 3) {
        long long r = v;
        const int w_ = (int)(r % Wp); r /= Wp;
        const int h_ = (int)(r % Hp); r /= Hp;
        const int d_ = (int)(r % Dp); r /= Dp;
        const long long n = r;
        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (w_ < W && h_ < H && d_ < D) {
            float f[8];
            u_unpack8(__ldg(x + v * 8 + sub), f);
#pragma unroll
            for (int k = 0; k < kHeadMaxK; ++k) {
                if (k >= K) break;
                const float go = __ldg(dout + ((n * K + k) * D + d_) * (long long)H * W + (long long)h_ * W + w_);
                if (sub == 0) ab[k] += go;
#pragma unroll
                for (int j = 0; j < 8; ++j) { g[j] = fmaf(go, sw[k * 64 + sub * 8 + j], g[j]); aw[k][j] = fmaf(go, f[j], aw[k][j]); }
            }
        }
        dx[v * 8 + sub] = u_pack8(g);
    }
#pragma unroll
    for (int k = 0; k < kHeadMaxK; ++k) {
        if (k >= K) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = aw[k][j];
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            if ((threadIdx.x & 31) < 8) atomicAdd(&sred[k * 65 + sub * 8 + j], a);
        }
        float b = ab[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sred[k * 65 + 64], b);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * 65; i += 256) partials[(size_t)blockIdx.x * K * 65 + i] = sred[i];
}
__global__ void __launch_bounds__(128) head_bwd_reduce_kernel(const float* __restrict__ partials, int nparts, int K, float* __restrict__ dw,
                                                              float* __restrict__ db) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * 65) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * K * 65 + i];
    const int k = i / 65, c = i % 65;
    if (c < 64) dw[k * 64 + c] = (float)s; else db[k] = (float)s;
}

static inline int grid_cap(long long work, int block, int cap) { return (int)std::max<long long>(1, std::min<long long>((work + block - 1) / block, cap)); }

}  // namespace mmad

using namespace mmad;
#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() do { MMAD_CUDA(cudaGetLastError()); count_launch(); return MMAD_OK; } while (0)

extern "C" {

static int c1_blocks(int N, int Do, int Ho, int Wo, int tw, int th, int td) {
    const long long tiles = (long long)N * ((Do + td - 1) / td) * ((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
    return (int)std::max<long long>(1, std::min<long long>(tiles, (long long)sm_count() * 4));
}
// blocks (== statistic partials) of the forward kernel / of the weight-gradient kernel
int mmad_conv3d_c1_blocks(int N, int Do, int Ho, int Wo) { return c1_blocks(N, Do, Ho, Wo, kF1Tw, kF1Th, kF1Td); }
int mmad_conv3d_c1_wgrad_blocks(int N, int Do, int Ho, int Wo) { return c1_blocks(N, Do, Ho, Wo, kC1Tw, kC1Th, kC1Td); }

int mmad_conv3d_c1_fwd(const float* x, const float* w, void* y, float* stats_partials, const float* ep_scale, const float* ep_shift, int N,
                       int D, int H, int W, int Do, int Ho, int Wo, int out_channels, void* stream) {
    MMAD_CHECK_ARG(out_channels == 64 || out_channels == 32, "conv3d_c1_fwd: y has 64-channel rows (upper half zero) or 32-channel rows");
    const int oc8 = out_channels / 8;
    MMAD_CHECK_ARG(x && w && y && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_c1_fwd: bad argument");
    MMAD_CHECK_ARG(Do >= D && Ho >= H && Wo >= W, "conv3d_c1_fwd: the output grid is the input grid zero-extended on the right");
    MMAD_CHECK_ARG((ep_scale == nullptr) == (ep_shift == nullptr), "conv3d_c1_fwd: epilogue scale and shift come together");
    MMAD_CHECK_ARG(!(ep_scale && stats_partials), "conv3d_c1_fwd: statistics are those of the raw output (no epilogue in training mode)");
    const dim3 grid(mmad_conv3d_c1_blocks(N, Do, Ho, Wo));
    const int tws = (Wo + kF1Tw - 1) / kF1Tw, ths = (Ho + kF1Th - 1) / kF1Th, tds = (Do + kF1Td - 1) / kF1Td;
    if (ep_scale) launch_pdl(conv3d_c1_fwd_kernel<false, true>, grid, dim3(256), 0, ST, x, w, (uint4*)y, stats_partials, ep_scale, ep_shift, N, D, H, W, Do, Ho, Wo, tws, ths, tds, oc8);
    else if (stats_partials) launch_pdl(conv3d_c1_fwd_kernel<true, false>, grid, dim3(256), 0, ST, x, w, (uint4*)y, stats_partials, ep_scale, ep_shift, N, D, H, W, Do, Ho, Wo, tws, ths, tds, oc8);
    else launch_pdl(conv3d_c1_fwd_kernel<false, false>, grid, dim3(256), 0, ST, x, w, (uint4*)y, stats_partials, ep_scale, ep_shift, N, D, H, W, Do, Ho, Wo, tws, ths, tds, oc8);
    LAUNCH_OK();
}

// partials: float[mmad_conv3d_c1_wgrad_blocks][32][27]; sum with mmad_wgrad_reduce(partials, blocks, dw, 32, 1, 27)
int mmad_conv3d_c1_wgrad(const float* x, const void* dy, float* partials, int N, int D, int H, int W, int Do, int Ho, int Wo, void* stream) {
    MMAD_CHECK_ARG(x && dy && partials && N > 0 && Do >= D && Ho >= H && Wo >= W, "conv3d_c1_wgrad: bad argument");
    launch_pdl(conv3d_c1_wgrad_kernel, dim3(mmad_conv3d_c1_wgrad_blocks(N, Do, Ho, Wo)), dim3(256), 0, ST, x, (const uint4*)dy, partials, N, D, H, W,
               Do, Ho, Wo, (Wo + kC1Tw - 1) / kC1Tw, (Ho + kC1Th - 1) / kC1Th, (Do + kC1Td - 1) / kC1Td);
    LAUNCH_OK();
}

int mmad_maxpool3d_k2_fwd(const void* x, int64_t ldx, void* y, void* idx, int N, int D, int H, int W, int C, void* stream) {
    MMAD_CHECK_ARG(x && y && C % 8 == 0 && D >= 2 && H >= 2 && W >= 2 && N > 0, "maxpool3d_k2_fwd: bad argument");
    MMAD_CHECK_ARG(ldx == 0 || (ldx >= C && ldx % 8 == 0), "maxpool3d_k2_fwd: ldx must be 0 (dense) or >= C and a multiple of 8");
    const int Do = D / 2, Ho = H / 2, Wo = W / 2, cv = C / 8;
    const long long total = (long long)N * Do * Ho * Wo * cv;
    launch_pdl(maxpool2_fwd_kernel, dim3(grid_cap(total, 256, sm_count() * 16)), dim3(256), 0, ST, (const uint4*)x, (int)((ldx ? ldx : C) / 8),
               (uint4*)y, (uint2*)idx, D, H, W, cv, Do, Ho, Wo, total);
    LAUNCH_OK();
}

int mmad_maxpool3d_k2_bwd(const void* dy, const void* idx, void* dx, int N, int D, int H, int W, int C, void* stream) {
    MMAD_CHECK_ARG(dy && idx && dx && C % 8 == 0 && D >= 2 && H >= 2 && W >= 2 && N > 0, "maxpool3d_k2_bwd: bad argument");
    const int Do = D / 2, Ho = H / 2, Wo = W / 2, cv = C / 8;
    if ((D | H | W) & 1) MMAD_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * D * H * W * C * 2, ST));    // odd tail: no window covers it
    const long long total = (long long)N * Do * Ho * Wo * cv;
    launch_pdl(maxpool2_bwd_kernel, dim3(grid_cap(total, 256, sm_count() * 16)), dim3(256), 0, ST, (const uint4*)dy, (const uint2*)idx, (uint4*)dx,
               D, H, W, cv, Do, Ho, Wo, total);
    LAUNCH_OK();
}

int mmad_convtranspose3d_prep_weights(const float* w, void* w_phases, int Cin, int Cout, void* stream) {
    MMAD_CHECK_ARG(w && w_phases && Cin > 0 && Cout > 0, "convtranspose3d_prep_weights: bad argument");
    launch_pdl(prep_convtranspose_weights_kernel, dim3(grid_cap(8ll * Cin * Cout, 256, sm_count() * 8)), dim3(256), 0, ST, w,
               (__nv_bfloat16*)w_phases, Cin, Cout);
    LAUNCH_OK();
}

int mmad_head1x1_fwd(const void* x, const float* w, const float* bias, float* out, int N, int Dp, int Hp, int Wp, int D, int H, int W, int C,
                     int K, void* stream) {
    MMAD_CHECK_ARG(x && w && bias && out && N > 0 && C == 64 && K >= 1 && K <= kHeadMaxK, "head1x1_fwd: C must be 64 and 1 <= K <= 8");
    MMAD_CHECK_ARG(D <= Dp && H <= Hp && W <= Wp && D > 0 && H > 0 && W > 0, "head1x1_fwd: the crop must lie inside the padded grid");
    const long long total = (long long)N * D * H * W * 8;
    launch_pdl(head_fwd_kernel, dim3(grid_cap(total, 256, sm_count() * 8)), dim3(256), 0, ST, (const uint4*)x, w, bias, out, N, Dp, Hp, Wp, D, H, W,
               K);
    LAUNCH_OK();
}

int mmad_head1x1_bwd_blocks(void) { return sm_count() * 4; }

// partials: float[mmad_head1x1_bwd_blocks()][K][65] scratch; dw float[K][64], db float[K]
int mmad_head1x1_bwd(const void* x, const float* w, const float* dout, void* dx, float* partials, float* dw, float* db, int N, int Dp, int Hp,
                     int Wp, int D, int H, int W, int C, int K, void* stream) {
    MMAD_CHECK_ARG(x && w && dout && dx && partials && dw && db && N > 0 && C == 64 && K >= 1 && K <= kHeadMaxK, "head1x1_bwd: bad argument");
    MMAD_CHECK_ARG(D <= Dp && H <= Hp && W <= Wp, "head1x1_bwd: the crop must lie inside the padded grid");
    const int blocks = mmad_head1x1_bwd_blocks();
    launch_pdl(head_bwd_kernel, dim3(blocks), dim3(256), 0, ST, (const uint4*)x, w, dout, (uint4*)dx, partials, N, Dp, Hp, Wp, D, H, W, K);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    launch_pdl(head_bwd_reduce_kernel, dim3((K * 65 + 127) / 128), dim3(128), 0, ST, (const float*)partials, blocks, K, dw, db);
    LAUNCH_OK();
}

}  // extern "C"
