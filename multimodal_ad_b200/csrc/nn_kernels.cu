// Bandwidth-bound companions of the Conv3d implicit GEMM (sm_100a): weight re-layout, BatchNorm3d
// statistics / apply / backward, ReLU, residual add, MaxPool3d, zero-insertion upsampling, layout conversion.
// Replaces nn.BatchNorm3d / nn.ReLU / nn.MaxPool3d / `out += residual` of /root/reference/models/resnet.py:46-69,
// :134-136, :204-213 (forward and backward).  All activations are NDHWC bf16 with C % 8 == 0; every kernel moves
// 16-byte vectors (8 channels) per thread.
#include "common.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>

namespace mmad {

struct bf16x8 { uint4 v; };
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}
static inline int grid_for(long long work, int block, int cap) { return (int)std::max<long long>(1, std::min<long long>((work + block - 1) / block, cap)); }
// grid for the per-channel-vector streaming kernels: grid * 256 must be a multiple of cv = C/8 so that a thread's grid-stride
// loop stays on one channel vector (its coefficients live in registers)
static inline int grid_for_channels(long long nvec, int cv, int cap) {
    int g = grid_for(nvec, 256, cap);
    int a = 256, b = cv;
    while (b) { const int t = a % b; a = b; b = t; }            // a = gcd(256, cv)
    const int m = cv / a;
    return std::max(m, g / m * m);
}

// ---- weights: torch (Cout, Cin, k,k,k) fp32 -> forward layout [Cout][taps][Cin] bf16 (pass 1: per (co, 64-ci chunk) transpose
//      of the [ci][tap] matrix through shared memory) and dgrad layout [Cin][taps reversed][Cout] bf16 (pass 2: per tap, a
//      32x32 tiled [co][ci] -> [ci][co] transpose of the forward layout).  Both passes read and write coalesced.
__global__ void __launch_bounds__(64) prep_weights_fwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int Cin, int taps) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    extern __shared__ float tile[];                  // [64][taps + 1]
    const int co = blockIdx.y, c0 = blockIdx.x * 64;
    const int nci = min(64, Cin - c0);
    const float* src = w + ((long long)co * Cin + c0) * taps;            // nci * taps contiguous floats
    for (int i = threadIdx.x; i < nci * taps; i += 64) tile[(i / taps) * (taps + 1) + (i % taps)] = src[i];
    __syncthreads();
    if (threadIdx.x < nci)
        for (int t = 0; t < taps; ++t)
            wf[((long long)co * taps + t) * Cin + c0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x * (taps + 1) + t]);
}
__global__ void __launch_bounds__(256) prep_weights_dgrad_kernel(const __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wt, int Cout,
                                                                 int Cin, int taps) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    __shared__ __nv_bfloat16 tile[32][34];
    const int tap = blockIdx.z, co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int co = co0 + j, ci = ci0 + tx;
        if (co < Cout && ci < Cin) tile[j][tx] = wf[((long long)co * taps + tap) * Cin + ci];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int ci = ci0 + j, co = co0 + tx;
        if (co < Cout && ci < Cin) wt[((long long)ci * taps + (taps - 1 - tap)) * Cout + co] = tile[tx][j];
    }
}
// ---- the same two passes for MANY layers in one launch each: a training step re-lays every convolution's weights (they change
//      every step), and ~40 launches of 5-10 us kernels cost more in launch gaps than in bytes.  The layer table travels in the
//      kernel parameter space (no upload); a block finds its layer by scanning the block-offset prefix.
constexpr int kPrepBatch = 64;
struct PrepDesc { const float* w; __nv_bfloat16* wf; __nv_bfloat16* wt; int Cout, Cin, taps; };
struct PrepBatch { int n; int start[kPrepBatch + 1]; PrepDesc d[kPrepBatch]; };
__device__ __forceinline__ int prep_find_layer(const PrepBatch& b, int blk) {
    int l = 0;
    while (l + 1 < b.n && blk >= b.start[l + 1]) ++l;
    return l;
}
__global__ void __launch_bounds__(64) prep_weights_fwd_batched_kernel(const __grid_constant__ PrepBatch b) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float tile[];                  // [64][taps + 1]
    const int l = prep_find_layer(b, blockIdx.x);
    const PrepDesc& d = b.d[l];
    const int local = blockIdx.x - b.start[l], chunks = (d.Cin + 63) / 64;
    const int co = local / chunks, c0 = (local % chunks) * 64, taps = d.taps;
    const int nci = min(64, d.Cin - c0);
    const float* src = d.w + ((long long)co * d.Cin + c0) * taps;
    for (int i = threadIdx.x; i < nci * taps; i += 64) tile[(i / taps) * (taps + 1) + (i % taps)] = src[i];
    __syncthreads();
    if (threadIdx.x < nci)
        for (int t = 0; t < taps; ++t)
            d.wf[((long long)co * taps + t) * d.Cin + c0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x * (taps + 1) + t]);
}
__global__ void __launch_bounds__(256) prep_weights_dgrad_batched_kernel(const __grid_constant__ PrepBatch b) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __nv_bfloat16 tile[32][34];
    const int l = prep_find_layer(b, blockIdx.x);
    const PrepDesc& d = b.d[l];
    int local = blockIdx.x - b.start[l];
    const int nx = (d.Cin + 31) / 32, ny = (d.Cout + 31) / 32;
    const int ci0 = (local % nx) * 32; local /= nx;
    const int co0 = (local % ny) * 32;
    const int tap = local / ny, taps = d.taps;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int co = co0 + j, ci = ci0 + tx;
        if (co < d.Cout && ci < d.Cin) tile[j][tx] = d.wf[((long long)co * taps + tap) * d.Cin + ci];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int ci = ci0 + j, co = co0 + tx;
        if (co < d.Cout && ci < d.Cin) d.wt[((long long)ci * taps + (taps - 1 - tap)) * d.Cout + co] = tile[tx][j];
    }
}
// ---- weights of the phase convolutions of mmad_conv3d_dgrad_s2_bf16: torch (Cdy, Cdx, 3,3,3) fp32 -> for phase p = pd*4+ph*2+pw
//      a block [Cdx][taps_p][Cdy] bf16, tap (a, b, c) of the phase = original tap t with t = 1 on an even axis, t = 2 - 2*a on an
//      odd one (dx[2j+1] = dy[j] w[2] + dy[j+1] w[0]).  Blocks are concatenated in phase order.
__global__ void __launch_bounds__(256) prep_weights_s2_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Cdy, int Cdx) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = 27ll * Cdx * Cdy;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i, base = 0;
        int p = 0, taps = 1;
        for (; p < 8; ++p) {                                   // locate the phase block
            taps = (1 + (p >> 2)) * (1 + ((p >> 1) & 1)) * (1 + (p & 1));
            const long long sz = (long long)taps * Cdx * Cdy;
            if (r < base + sz) break;
            base += sz;
        }
        r -= base;
        const int cdy = (int)(r % Cdy); r /= Cdy;
        const int tap = (int)(r % taps); r /= taps;
        const int cdx = (int)r;
        const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        const int kwp = 1 + pw, khp = 1 + ph;
        const int c = tap % kwp, b = (tap / kwp) % khp, a = tap / (kwp * khp);
        const int td = pd ? 2 - 2 * a : 1, th = ph ? 2 - 2 * b : 1, tw = pw ? 2 - 2 * c : 1;
        wp[i] = __float2bfloat16(w[((size_t)cdy * Cdx + cdx) * 27 + (td * 3 + th) * 3 + tw]);
    }
}

// ---- BatchNorm statistics: per-CTA partial (sum, sum of squares) -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale;
//      running statistics updated like nn.BatchNorm3d (momentum, unbiased variance).  One thread per channel, double accumulation.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(128) bn_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, float* __restrict__ scale_out, float* __restrict__ shift_out) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        // one warp per channel
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < nparts; p += 32) {
        const float2 v = *reinterpret_cast<const float2*>(partials + ((size_t)p * C + c) * 2);
        s += v.x; q += v.y;
    }
    s = warp_sum_d(s); q = warp_sum_d(q);
    if (lane) return;
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    mean_out[c] = (float)mean; invstd_out[c] = invstd; scale_out[c] = sc; shift_out[c] = beta[c] - (float)mean * sc;
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
// eval mode: scale / shift from the running statistics
__global__ void bn_eval_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ running_mean,
                               const float* __restrict__ running_var, float eps, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                               float* __restrict__ scale_out, float* __restrict__ shift_out) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float invstd = rsqrtf(running_var[c] + eps);
    const float sc = gamma[c] * invstd;
    mean_out[c] = running_mean[c]; invstd_out[c] = invstd; scale_out[c] = sc; shift_out[c] = beta[c] - running_mean[c] * sc;
}

// ---- y = act(x*scale + shift  [+ res*rscale + rshift | + res]);  x, res bf16 NDHWC;  out bf16 and/or fp32 (same NDHWC order)
template <bool RELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                                       const uint4* __restrict__ res, const float* __restrict__ rscale,
                                                       const float* __restrict__ rshift, uint4* __restrict__ out_bf16,
                                                       float4* __restrict__ out_f32, long long nvec, int C, int out_ld8) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    // the grid stride is a multiple of C/8 (see the launcher): a thread keeps one channel vector, coefficients in registers
    const int cv = C >> 3;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c0 = (int)(i0 % cv) * 8;
    float sc[8], sh[8], rs[8], rb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j];
        rs[j] = rscale ? rscale[c0 + j] : 1.f; rb[j] = rscale ? rshift[c0 + j] : 0.f;
    }
    for (long long i = i0; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float f[8];
        unpack8(x[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
        if (res) {
            float r[8];
            unpack8(res[i], r);
            if (rscale) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += fmaf(r[j], rs[j], rb[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += r[j];
            }
        }
        if (RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        // out_ld8: distance between output rows in 16-byte vectors (== cv when dense); the bf16 output may be a channel slice of a
        // wider NDHWC tensor (the skip half of a concatenation buffer, unet3d.py:77)
        if (out_bf16) out_bf16[out_ld8 == cv ? i : (i / cv) * out_ld8 + (i % cv)] = pack8(f);
        if (out_f32) { out_f32[2 * i] = make_float4(f[0], f[1], f[2], f[3]); out_f32[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]); }
    }
}

// ---- BatchNorm / ReLU backward, pass 1: g = (dy [+ dy2]) * (mask > 0) ; per-channel partial sums of g and g*xhat.
//      dy may be bf16 (dy_bf16) or fp32 (dy_f32, same NDHWC order).  Writes g (bf16) when g_out != NULL.
//      Block = 256 threads = 8 row-lanes x 32 channel-vectors... generic: thread owns channel vector (i % cv); block partials in smem.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const uint4* __restrict__ dy_bf16, const float4* __restrict__ dy_f32,
                                                            const uint4* __restrict__ dy2, const uint4* __restrict__ mask,
                                                            const uint4* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ mscale,
                                                            const float* __restrict__ mshift, uint4* __restrict__ g_out,
                                                            float* __restrict__ partials, long long rows, int C) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    // each block walks rows [r0, r1); thread t handles channel vector (t % cv) of rows r0 + t / cv, + 256/cv, ...
    extern __shared__ float red[];   // [256][16]
    const int cv = C >> 3;           // vectors per row: 8, 16, 32 or 64
    const int rpb = 256 / cv;        // rows per block step (>= 4)
    const int tv = threadIdx.x % cv, tr = threadIdx.x / cv;
    const int c0 = tv * 8;
    float mu[8], is[8], msc[8], msh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j];
        msc[j] = mscale ? mscale[c0 + j] : 0.f; msh[j] = mscale ? mshift[c0 + j] : 0.f;
    }
    float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sgx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long per = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = per * blockIdx.x, r1 = min(rows, r0 + per);
    for (long long r = r0 + tr; r < r1; r += rpb) {
        const long long i = r * cv + tv;
        float g[8];
        if (dy_bf16) unpack8(dy_bf16[i], g);
        else { const float4 a = dy_f32[2 * i], b = dy_f32[2 * i + 1]; g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w; }
        if (dy2) { float h[8]; unpack8(dy2[i], h);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] += h[j]; }
        if (mask) { float m[8]; unpack8(mask[i], m);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = m[j] > 0.f ? g[j] : 0.f; }
        float xv[8];
        unpack8(x[i], xv);
        if (mscale) {                 // ReLU mask recomputed from the pre-BN tensor: relu(x*scale + shift) > 0
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = fmaf(xv[j], msc[j], msh[j]) > 0.f ? g[j] : 0.f;
        }
        if (g_out) {
            const uint4 gp = pack8(g);
            g_out[i] = gp;
            unpack8(gp, g);          // statistics of the ROUNDED g, the values pass 2 will read
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { sg[j] += g[j]; sgx[j] += g[j] * (xv[j] - mu[j]) * is[j]; }
    }
    float* my = red + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my[j] = sg[j]; my[8 + j] = sgx[j]; }
    __syncthreads();
    if (threadIdx.x < cv) {          // thread tv sums the rpb row-lanes of its channel vector
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = 0.f;
        for (int q = 0; q < rpb; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] += red[(q * cv + threadIdx.x) * 16 + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            partials[((size_t)blockIdx.x * C + c0 + j) * 2] = a[j];
            partials[((size_t)blockIdx.x * C + c0 + j) * 2 + 1] = a[8 + j];
        }
    }
}

// ---- the same pass with the inputs staged through shared memory by the TMA engine (bf16 dy).  The register version above
//      keeps only 2-4 x 16 bytes per thread in flight (80 registers -> 3 blocks per SM) and reaches ~1.9 TB/s; here a producer
//      warp streams tiles of 512 vectors (8 KB per input tensor) into a 6-deep ring with 1-D bulk copies, so ~190 KB per SM
//      are in flight whatever the consumers' register count.  512 consumer threads; thread t always owns channel vector
//      t % cv (cv divides 512), so the statistics stay in registers.
constexpr int kBnTileVec = 512;                  // 16-byte vectors per tile and input tensor
constexpr int kBnStages = 6;
constexpr int kBnConsumers = 512;

__global__ void __launch_bounds__(kBnConsumers + 32, 1)
bn_bwd_reduce_tma_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ dy2, const uint4* __restrict__ mask,
                         const uint4* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ mscale, const float* __restrict__ mshift, uint4* __restrict__ g_out,
                         float* __restrict__ partials, long long nvec, int C) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    extern __shared__ __align__(128) unsigned char bsm[];
    const int nin = 2 + (dy2 ? 1 : 0) + (mask ? 1 : 0);
    const uint32_t stage_bytes = (uint32_t)nin * kBnTileVec * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + (size_t)kBnStages * 4 * kBnTileVec * 16);   // full[S], empty[S]
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kBnStages;
    const int warp = threadIdx.x >> 5;
    const long long tiles = (nvec + kBnTileVec - 1) / kBnTileVec;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kBnStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kBnConsumers / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kBnConsumers / 32) {
        // ---------------- producer warp ----------------
        uint32_t s = 0, ph = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (elect_one()) {
                const long long v0 = t * kBnTileVec;
                const uint32_t bytes = (uint32_t)min((long long)kBnTileVec, nvec - v0) * 16u;
                mbar_arrive_expect_tx(full0 + 8 * s, bytes * (uint32_t)nin);
                const uint32_t dst = smem_u32(bsm) + s * stage_bytes;
                bulk_g2s(dst, dy + v0, bytes, full0 + 8 * s);
                bulk_g2s(dst + kBnTileVec * 16, x + v0, bytes, full0 + 8 * s);
                uint32_t o = 2 * kBnTileVec * 16;
                if (dy2) { bulk_g2s(dst + o, dy2 + v0, bytes, full0 + 8 * s); o += kBnTileVec * 16; }
                if (mask) bulk_g2s(dst + o, mask + v0, bytes, full0 + 8 * s);
            }
            __syncwarp();
            if (++s == kBnStages) { s = 0; ph ^= 1; }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int cv = C >> 3;
    const int tv = threadIdx.x % cv;
    const int c0 = tv * 8;
    float mu[8], is[8], msc[8], msh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j];
        msc[j] = mscale ? mscale[c0 + j] : 0.f; msh[j] = mscale ? mshift[c0 + j] : 0.f;
    }
    float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sgx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t s = 0, ph = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        mbar_wait(full0 + 8 * s, ph);
        const uint4* st = reinterpret_cast<const uint4*>(bsm + (size_t)s * stage_bytes);
        const long long i = t * kBnTileVec + threadIdx.x;
        if (i < nvec) {
            float g[8], xv[8];
            unpack8(st[threadIdx.x], g);
            unpack8(st[kBnTileVec + threadIdx.x], xv);
            int o = 2 * kBnTileVec;
            if (dy2) { float h[8]; unpack8(st[o + threadIdx.x], h); o += kBnTileVec;
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] += h[j]; }
            if (mask) { float m[8]; unpack8(st[o + threadIdx.x], m);
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = m[j] > 0.f ? g[j] : 0.f; }
            if (mscale) {
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = fmaf(xv[j], msc[j], msh[j]) > 0.f ? g[j] : 0.f;
            }
            if (g_out) {
                const uint4 gp = pack8(g);
                g_out[i] = gp;
                unpack8(gp, g);          // statistics of the ROUNDED g, the values pass 2 will read
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { sg[j] += g[j]; sgx[j] += g[j] * (xv[j] - mu[j]) * is[j]; }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == kBnStages) { s = 0; ph ^= 1; }
    }
    // block reduction through the (now idle) staging memory: [512][16] floats
    asm volatile("bar.sync 1, %0;" ::"n"(kBnConsumers) : "memory");
    float* red = reinterpret_cast<float*>(bsm);
    float* my = red + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my[j] = sg[j]; my[8 + j] = sgx[j]; }
    asm volatile("bar.sync 1, %0;" ::"n"(kBnConsumers) : "memory");
    if (threadIdx.x < cv) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = 0.f;
        for (int q = 0; q < kBnConsumers / cv; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] += red[(q * cv + threadIdx.x) * 16 + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            partials[((size_t)blockIdx.x * C + c0 + j) * 2] = a[j];
            partials[((size_t)blockIdx.x * C + c0 + j) * 2 + 1] = a[8 + j];
        }
    }
}
// sums the block partials (one warp per channel); writes dgamma, dbeta and the three per-channel coefficients of pass 2:
//   dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) = A*g + B*x + Cc,  A = gamma*invstd, B = -A*invstd*mgx,
//   Cc = A*(mean*invstd*mgx - mg).   Eval mode (training == 0): BatchNorm is affine, dx = A*g.
__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                                       int training, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < nparts; p += 32) {
        const float2 v = *reinterpret_cast<const float2*>(partials + ((size_t)p * C + c) * 2);
        s += v.x; q += v.y;
    }
    s = warp_sum_d(s); q = warp_sum_d(q);
    if (lane) return;
    if (dbeta) dbeta[c] = (float)s;
    if (dgamma) dgamma[c] = (float)q;
    const float mg = training ? (float)(s / count) : 0.f, mgx = training ? (float)(q / count) : 0.f;
    const float a = gamma[c] * invstd[c];
    coef[c] = a;
    coef[C + c] = -a * invstd[c] * mgx;
    coef[2 * C + c] = a * (mean[c] * invstd[c] * mgx - mg);
}
// pass 2: dx = A*g + B*x + Cc
// mscale / mshift (optional): g is the UNMASKED upstream gradient and the ReLU mask relu(x * mscale + mshift) > 0 is recomputed here
// from the pre-BatchNorm tensor - pass 1 then never writes g (and never reads a stored activation as the mask): four tensor passes
// per BatchNorm backward instead of six.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ x, const float* __restrict__ coef,
                                                           const float* __restrict__ mscale, const float* __restrict__ mshift,
                                                           uint4* __restrict__ dx, long long nvec, int C) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    // the grid stride (gridDim.x * 256) is a multiple of C/8, so a thread keeps ONE channel vector: its 24 coefficients are
    // loaded once instead of per element (they cost three times the tensor bytes in L1 traffic)
    const int cv = C >> 3;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c0 = (int)(i0 % cv) * 8;
    float A[8], B[8], Cc[8], ms[8], mb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        A[j] = coef[c0 + j]; B[j] = coef[C + c0 + j]; Cc[j] = coef[2 * C + c0 + j];
        ms[j] = mscale ? mscale[c0 + j] : 0.f; mb[j] = mscale ? mshift[c0 + j] : 1.f;       // no mask: 0 * x + 1 > 0 always
    }
    for (long long i = i0; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float gv[8], xv[8], o[8];
        unpack8(g[i], gv);
        unpack8(x[i], xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gm = fmaf(xv[j], ms[j], mb[j]) > 0.f ? gv[j] : 0.f;
            o[j] = fmaf(A[j], gm, fmaf(B[j], xv[j], Cc[j]));
        }
        dx[i] = pack8(o);
    }
}

// ---- MaxPool3d k3 s2 p1 (resnet.py:136), NDHWC bf16; the winning tap index (0..26, first maximum in (kd,kh,kw) order) is kept
//      for the backward pass.
__global__ void __launch_bounds__(256) maxpool3d_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint2* __restrict__ idx, int N,
                                                            int D, int H, int W, int C, int Do, int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int cv = C >> 3;
    const long long total = (long long)N * Do * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int v = (int)(t % cv); t /= cv;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do); t /= Do;
        const int n = (int)t;
        float best[8];
        int bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
        int tap = 0;
        for (int kd = 0; kd < 3; ++kd)
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw, ++tap) {
                    const int id = od * 2 + kd - 1, ih = oh * 2 + kh - 1, iw = ow * 2 + kw - 1;
                    if (id < 0 || id >= D || ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                    float f[8];
                    unpack8(x[((((long long)n * D + id) * H + ih) * W + iw) * cv + v], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (f[j] > best[j]) { best[j] = f[j]; bi[j] = tap; }
                }
        y[i] = pack8(best);
        uint2 p;
        p.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
        p.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
        idx[i] = p;
    }
}
// gradient of the max-pool w.r.t. one input voxel vector (8 channels): sum of (dp [+ dp2]) over the <= 8 windows that
// contain the voxel and picked it.  id, ih are block-uniform in the callers, so only the w loop can diverge.
__device__ __forceinline__ void stem_pool_gather(const uint4* __restrict__ dp, const uint4* __restrict__ dp2, const uint2* __restrict__ idx,
                                                 int n, int id, int ih, int iw, int v, int cv, int Do, int Ho, int Wo, float (&g)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    const int od1 = ((id & 1) && ((id + 1) >> 1) < Do) ? (id + 1) >> 1 : id >> 1;
    const int oh1 = ((ih & 1) && ((ih + 1) >> 1) < Ho) ? (ih + 1) >> 1 : ih >> 1;
    const int ow1 = ((iw & 1) && ((iw + 1) >> 1) < Wo) ? (iw + 1) >> 1 : iw >> 1;
    for (int od = id >> 1; od <= od1; ++od)
        for (int oh = ih >> 1; oh <= oh1; ++oh)
            for (int ow = iw >> 1; ow <= ow1; ++ow) {
                const uint32_t tap = (uint32_t)(((id - 2 * od + 1) * 3 + (ih - 2 * oh + 1)) * 3 + (iw - 2 * ow + 1));
                const long long o = ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * cv + v;
                const uint2 p = idx[o];
                uint4 gy = dp[o];
                const uint32_t t4 = tap * 0x01010101u;
                const uint32_t mlo = __vcmpeq4(p.x, t4), mhi = __vcmpeq4(p.y, t4);      // 0xff per matching channel byte
                if (!(mlo | mhi)) continue;
                float a[8];
                if (dp2) {
                    float b[8];
                    unpack8(gy, a);
                    unpack8(dp2[o], b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] += b[j];
                } else {
                    unpack8(gy, a);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] += (((j < 4 ? mlo : mhi) >> (8 * (j & 3))) & 1u) ? a[j] : 0.f;
            }
}

// gather form of the backward: every input voxel collects from the (<= 2 per axis) windows that contain it and picked it.
// blockIdx.x = (n, id), blockIdx.y = group of `hrows` rows; threads cover (iw, channel vector): no per-thread divisions and
// the d / h window loops are block-uniform.
__global__ void __launch_bounds__(512) maxpool3d_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx, uint4* __restrict__ dx,
                                                            int N, int D, int H, int W, int C, int Do, int Ho, int Wo, int hrows) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int cv = C >> 3;
    const int v = threadIdx.x % cv, wl = threadIdx.x / cv, wstep = blockDim.x / cv;
    const int n = blockIdx.x / D, id = blockIdx.x % D;
    for (int ih = blockIdx.y * hrows; ih < min(H, (int)(blockIdx.y + 1) * hrows); ++ih)
        for (int iw = wl; iw < W; iw += wstep) {
            float acc[8];
            stem_pool_gather(dy, nullptr, idx, n, id, ih, iw, v, cv, Do, Ho, Wo, acc);
            dx[((((long long)n * D + id) * H + ih) * W + iw) * cv + v] = pack8(acc);
        }
}


// ---- the same gather, marching along D with the pooled gradient staged in shared memory.  A block owns 4 input rows
//      (all of W) of one sample and walks a range of input slices; the <= 3 pooled rows x Wo of gradient + argmax bytes that
//      a pooled slice contributes are streamed into a 3-slot ring with 1-D bulk copies (each pooled slice serves 3 input
//      slices), so the per-voxel window tests read shared memory instead of issuing dependent global loads.
constexpr int kPoolSlots = 3;
__global__ void __launch_bounds__(512) maxpool3d_bwd_march_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                                  uint4* __restrict__ dx, int N, int D, int H, int W, int C, int Do,
                                                                  int Ho, int Wo, int dsplit) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    extern __shared__ __align__(128) unsigned char psm[];
    const int cv = C >> 3;
    const int hgroups = (H + 3) >> 2;
    int b = blockIdx.x;
    const int ds = b % dsplit; b /= dsplit;
    const int hg = b % hgroups; b /= hgroups;
    const int n = b;
    const int ih0 = hg * 4, ih1 = min(H, ih0 + 4);
    const int oh_lo = ih0 >> 1, nrow = min(Ho, oh_lo + 3) - oh_lo;        // pooled rows touching input rows ih0..ih0+3
    const int d_lo = (int)((long long)D * ds / dsplit), d_hi = (int)((long long)D * (ds + 1) / dsplit);   // [d_lo, d_hi)
    const int od_start = d_lo >> 1, od_last = min(Do - 1, d_hi >> 1);      // pooled slices needed: od_start..od_last
    const uint32_t dy_bytes = (uint32_t)nrow * Wo * cv * 16, ix_bytes = dy_bytes >> 1;
    const uint32_t slot_bytes = (uint32_t)3 * Wo * cv * 24;               // dy tile, then idx tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(psm + kPoolSlots * slot_bytes);
    const uint32_t full0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int q = 0; q < kPoolSlots; ++q) mbar_init(full0 + 8 * q, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto load = [&](int od) {                                              // thread 0 only
        const int q = od - od_start, slot = q % kPoolSlots;
        const size_t e = (((size_t)n * Do + od) * Ho + oh_lo) * Wo * cv;
        const uint32_t dst = smem_u32(psm) + slot * slot_bytes;
        mbar_arrive_expect_tx(full0 + 8 * slot, dy_bytes + ix_bytes);
        bulk_g2s(dst, dy + e, dy_bytes, full0 + 8 * slot);
        bulk_g2s(dst + 3 * Wo * cv * 16, idx + e, ix_bytes, full0 + 8 * slot);
    };
    if (threadIdx.x == 0)
        for (int od = od_start; od <= min(od_last, od_start + kPoolSlots - 1); ++od) load(od);

    const int v = threadIdx.x % cv, wstep = blockDim.x / cv;
    // A warp covers 32 / cv..4 voxels of W; odd voxels look at two pooled columns, even ones at one.  Give every warp voxels
    // of ONE parity (0,2,4,6 | 1,3,5,7 | 8,10,...) so that the w loop does not diverge.
    int wl = threadIdx.x / cv;
    if ((wstep & 7) == 0) wl = ((wl >> 3) << 3) + ((wl & 3) << 1) + ((wl >> 2) & 1);
    int waited = od_start - 1;                                             // pooled slices whose arrival this thread has seen
    for (int id = d_lo; id < d_hi; ++id) {
        const int od0 = id >> 1;
        const int od1 = ((id & 1) && od0 + 1 < Do) ? od0 + 1 : od0;
        while (waited < od1) {
            ++waited;
            const int q = waited - od_start;
            mbar_wait(full0 + 8 * (q % kPoolSlots), (q / kPoolSlots) & 1);
        }
        for (int ih = ih0; ih < ih1; ++ih) {
            const int oha = ih >> 1;
            const int ohb = ((ih & 1) && oha + 1 < Ho) ? oha + 1 : oha;
            for (int iw = wl; iw < W; iw += wstep) {
                const int owa = iw >> 1;
                const int owb = ((iw & 1) && owa + 1 < Wo) ? owa + 1 : owa;
                float g[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = 0.f;
                for (int od = od0; od <= od1; ++od) {
                    const unsigned char* sl = psm + ((od - od_start) % kPoolSlots) * slot_bytes;
                    const uint4* sdy = reinterpret_cast<const uint4*>(sl);
                    const uint2* six = reinterpret_cast<const uint2*>(sl + 3 * Wo * cv * 16);
                    for (int oh = oha; oh <= ohb; ++oh)
                        for (int ow = owa; ow <= owb; ++ow) {
                            const uint32_t tap = (uint32_t)(((id - 2 * od + 1) * 3 + (ih - 2 * oh + 1)) * 3 + (iw - 2 * ow + 1));
                            const int o = ((oh - oh_lo) * Wo + ow) * cv + v;
                            const uint2 p = six[o];
                            uint4 gy = sdy[o];
                            const uint32_t t4 = tap * 0x01010101u;
                            const uint32_t mlo = __vcmpeq4(p.x, t4), mhi = __vcmpeq4(p.y, t4);      // 0xff per matching channel byte
                            // widen the byte masks to the bf16 lanes and clear the channels this window did not pick
                            gy.x &= __byte_perm(mlo, 0, 0x1100); gy.y &= __byte_perm(mlo, 0, 0x3322);
                            gy.z &= __byte_perm(mhi, 0, 0x1100); gy.w &= __byte_perm(mhi, 0, 0x3322);
                            float a[8];
                            unpack8(gy, a);
#pragma unroll
                            for (int j = 0; j < 8; ++j) g[j] += a[j];
                        }
                }
                dx[((((size_t)n * D + id) * H + ih) * W + iw) * cv + v] = pack8(g);
            }
        }
        if (id & 1) {                                                      // pooled slice od0 is done: refill its slot
            __syncthreads();
            if (threadIdx.x == 0 && od0 + kPoolSlots <= od_last) load(od0 + kPoolSlots);
        }
    }
}


// ---- the same fused stem pass, marching along D with the conv output staged in shared memory.  A block owns two pooled rows
//      (all of W) of one sample and a range of pooled slices; the five input rows x W x C of every input slice are streamed once
//      into a 5-slot ring (1-D bulk copy), turned IN PLACE into the stored activation bf16(relu(bn(c))) - once per element
//      instead of once per window that contains it (3.4x) - and pooled with packed 16-bit compares (the activations are
//      non-negative, so their bf16 bit patterns order like the values; +1 per lane makes "no tap yet" = 0).
constexpr int kSpSlots = 5;
__global__ void __launch_bounds__(512, 1)
stem_pool_march_kernel(const uint4* __restrict__ c, const float* __restrict__ scale, const float* __restrict__ shift, uint4* __restrict__ y,
                       uint2* __restrict__ idx, int N, int D, int H, int W, int C, int Do, int Ho, int Wo, int dsplit) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char ssm[];
    const int cv = C >> 3;
    const int hgroups = (Ho + 1) >> 1;
    int b = blockIdx.x;
    const int ds = b % dsplit; b /= dsplit;
    const int hg = b % hgroups; b /= hgroups;
    const int n = b;
    const int oh0 = hg * 2, nro = min(2, Ho - oh0);                       // pooled rows of this block
    const int ihb = 2 * oh0 - 1;                                          // input row of tile row 0 (may be -1)
    const int ih_lo = max(0, ihb), ih_hi = min(H - 1, 2 * (oh0 + nro - 1) + 1);
    const int od_lo = (int)((long long)Do * ds / dsplit), od_hi = (int)((long long)Do * (ds + 1) / dsplit);
    if (od_lo >= od_hi) return;
    const int id_lo = max(0, 2 * od_lo - 1), id_hi = min(D - 1, 2 * (od_hi - 1) + 1);   // input slices needed
    const uint32_t row_bytes = (uint32_t)W * cv * 16, slot_bytes = 5u * row_bytes;
    const uint32_t copy_bytes = (uint32_t)(ih_hi - ih_lo + 1) * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ssm + kSpSlots * slot_bytes);
    const uint32_t full0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int q = 0; q < kSpSlots; ++q) mbar_init(full0 + 8 * q, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto load = [&](int id) {                                             // thread 0 only
        const int slot = (id - id_lo) % kSpSlots;
        mbar_arrive_expect_tx(full0 + 8 * slot, copy_bytes);
        bulk_g2s(smem_u32(ssm) + slot * slot_bytes + (uint32_t)(ih_lo - ihb) * row_bytes,
                 c + (((size_t)n * D + id) * H + ih_lo) * W * cv, copy_bytes, full0 + 8 * slot);
    };
    if (threadIdx.x == 0)
        for (int id = id_lo; id <= min(id_hi, id_lo + kSpSlots - 1); ++id) load(id);

    const int tv = threadIdx.x % cv;                                      // 512 % cv == 0: one channel vector per thread
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[tv * 8 + j]; sh[j] = shift[tv * 8 + j]; }
    const int row_vec = W * cv, n_in = (ih_hi - ih_lo + 1) * row_vec, in_off = (ih_lo - ihb) * row_vec;

    int ready = id_lo - 1;                                                // input slices already transformed
    for (int od = od_lo; od < od_hi; ++od) {
        const int s_hi = min(D - 1, 2 * od + 1);
        // 1. transform the slices that arrived since the last step
        for (int id = ready + 1; id <= s_hi; ++id) {
            const int q = id - id_lo;
            mbar_wait(full0 + 8 * (q % kSpSlots), (q / kSpSlots) & 1);
            uint4* sl = reinterpret_cast<uint4*>(ssm + (q % kSpSlots) * slot_bytes) + in_off;
            for (int i = threadIdx.x; i < n_in; i += 512) {
                float f[8];
                unpack8(sl[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                uint4 a = pack8(f);
                a.x &= 0x7fff7fffu; a.y &= 0x7fff7fffu; a.z &= 0x7fff7fffu; a.w &= 0x7fff7fffu;     // -0 -> +0
                sl[i] = a;
            }
        }
        ready = s_hi;
        __syncthreads();
        // 2. pool: one pooled vector (8 channels) per thread and iteration
        for (int o = threadIdx.x; o < nro * Wo * cv; o += 512) {
            const int v = o % cv, ow = (o / cv) % Wo, ohl = o / (cv * Wo);
            uint32_t best[4] = {0u, 0u, 0u, 0u}, bi[4] = {0u, 0u, 0u, 0u};
            uint32_t tap = 0;
            for (int kd = 0; kd < 3; ++kd) {
                const int id = 2 * od + kd - 1;
                if ((unsigned)id >= (unsigned)D) { tap += 9; continue; }
                const uint4* sl = reinterpret_cast<const uint4*>(ssm + ((id - id_lo) % kSpSlots) * slot_bytes);
                for (int kh = 0; kh < 3; ++kh) {
                    const int r = 2 * ohl + kh;                           // tile row; input row ihb + r
                    if ((unsigned)(ihb + r) >= (unsigned)H) { tap += 3; continue; }
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw, ++tap) {
                        const int iw = 2 * ow + kw - 1;
                        if ((unsigned)iw >= (unsigned)W) continue;
                        const uint4 a = sl[(r * W + iw) * cv + v];
                        const uint32_t t2 = tap * 0x00010001u;
                        const uint32_t av[4] = {a.x + 0x00010001u, a.y + 0x00010001u, a.z + 0x00010001u, a.w + 0x00010001u};
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            const uint32_t m = __vcmpgtu2(av[w4], best[w4]);   // strictly greater: the first maximum wins
                            best[w4] = __vmaxu2(av[w4], best[w4]);
                            bi[w4] = (bi[w4] & ~m) | (t2 & m);
                        }
                    }
                }
            }
            const size_t oi = ((((size_t)n * Do + od) * Ho + oh0 + ohl) * Wo + ow) * cv + v;
            y[oi] = make_uint4(best[0] - 0x00010001u, best[1] - 0x00010001u, best[2] - 0x00010001u, best[3] - 0x00010001u);
            idx[oi] = make_uint2(__byte_perm(bi[0], bi[1], 0x6420), __byte_perm(bi[2], bi[3], 0x6420));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // in-place (generic) writes before the refilling bulk copies
        __syncthreads();
        // 3. slices 2od-1 and 2od are dead: refill their slots with the slices two steps ahead
        if (threadIdx.x == 0)
            for (int id = 2 * od - 1; id <= 2 * od; ++id)
                if (id >= id_lo && id + kSpSlots <= id_hi) load(id + kSpSlots);
    }
}

// ---- fused stem (resnet.py:206-208): p = maxpool3d(relu(bn(c)), k3 s2 p1) in one pass over the conv output c; the
//      post-ReLU tensor is never stored.  idx keeps the winning tap (first maximum in (kd,kh,kw) order).
__global__ void __launch_bounds__(256) stem_bn_relu_maxpool_fwd_kernel(const uint4* __restrict__ c, const float* __restrict__ scale,
                                                                       const float* __restrict__ shift, uint4* __restrict__ y,
                                                                       uint2* __restrict__ idx, int N, int D, int H, int W, int C, int Do,
                                                                       int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int cv = C >> 3;
    const long long total = (long long)N * Do * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        unsigned t = (unsigned)i;
        const int v = (int)(t % (unsigned)cv); t /= (unsigned)cv;
        const int ow = (int)(t % (unsigned)Wo); t /= (unsigned)Wo;
        const int oh = (int)(t % (unsigned)Ho); t /= (unsigned)Ho;
        const int od = (int)(t % (unsigned)Do); t /= (unsigned)Do;
        const int n = (int)t;
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[v * 8 + j]; sh[j] = shift[v * 8 + j]; }
        float best[8];
        int bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
        int tap = 0;
        for (int kd = 0; kd < 3; ++kd)
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw, ++tap) {
                    const int id = od * 2 + kd - 1, ih = oh * 2 + kh - 1, iw = ow * 2 + kw - 1;
                    if ((unsigned)id >= (unsigned)D || (unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) continue;
                    float f[8];
                    unpack8(c[((((long long)n * D + id) * H + ih) * W + iw) * cv + v], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // the activation the unfused path would have STORED: bf16(relu(bn(c)))
                        const float a = __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f)));
                        if (a > best[j]) { best[j] = a; bi[j] = tap; }
                    }
                }
        y[i] = pack8(best);
        uint2 p;
        p.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
        p.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
        idx[i] = p;
    }
}

// ---- zero insertion: y (N, Dy,Hy,Wy, C), y[2*o] = x[o], zero elsewhere (dgrad of a stride-2 convolution = unit-stride
//      convolution of the zero-upsampled gradient with the flipped kernel)
__global__ void __launch_bounds__(256) upsample_zero2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int Dx, int Hx, int Wx,
                                                             int Dy, int Hy, int Wy, int C) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int cv = C >> 3;
    const long long total = (long long)N * Dy * Hy * Wy * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int v = (int)(t % cv); t /= cv;
        const int w = (int)(t % Wy); t /= Wy;
        const int h = (int)(t % Hy); t /= Hy;
        const int d = (int)(t % Dy); t /= Dy;
        const int n = (int)t;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (!((w | h | d) & 1) && (w >> 1) < Wx && (h >> 1) < Hx && (d >> 1) < Dx)
            o = x[((((long long)n * Dx + (d >> 1)) * Hx + (h >> 1)) * Wx + (w >> 1)) * cv + v];
        y[i] = o;
    }
}

// ---- layout: (N, C, S) fp32 (torch NCDHW, S = D*H*W) -> (N, S, C) bf16, 32x32 shared-memory transpose
__global__ void __launch_bounds__(256) ncs_f32_to_nsc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, long long S) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long s0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long s = s0 + tx;
        tile[j][tx] = (c < C && s < S) ? x[((long long)n * C + c) * S + s] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const long long s = s0 + j;
        const int c = c0 + tx;
        if (c < C && s < S) y[((long long)n * S + s) * C + c] = __float2bfloat16_rn(tile[tx][j]);
    }
}

// ---- wgrad epilogue: sum the split-K partials [nsplit][Cout][taps][Cin] fp32 -> torch layout (Cout, Cin, taps) fp32.
//      Block = (co, 64-ci chunk), 256 threads: coalesced reads of the [tap][ci] rows of every split (4 independent partial
//      sums per thread), shared-memory transpose, coalesced [ci][tap] writes.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, float* __restrict__ dw, int Cout, int Cin,
                                                           int taps, int Cin_total, int ci_off) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    extern __shared__ float tile[];                  // [taps][65]
    const int co = blockIdx.y, c0 = blockIdx.x * 64;
    const int nci = min(64, Cin - c0);
    const size_t plane = (size_t)Cout * taps * Cin;
    for (int e = threadIdx.x; e < taps * 64; e += 256) {
        const int t = e >> 6, c = e & 63;
        if (c >= nci) continue;
        const float* src = part + ((size_t)co * taps + t) * Cin + c0 + c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int p = 0;
        for (; p + 4 <= nsplit; p += 4) {
            s0 += src[(size_t)p * plane]; s1 += src[(size_t)(p + 1) * plane];
            s2 += src[(size_t)(p + 2) * plane]; s3 += src[(size_t)(p + 3) * plane];
        }
        for (; p < nsplit; ++p) s0 += src[(size_t)p * plane];
        tile[t * 65 + c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    float* dst = dw + ((size_t)co * Cin_total + ci_off + c0) * taps;     // nci * taps contiguous floats
    for (int i = threadIdx.x; i < nci * taps; i += 256) dst[i] = tile[(i % taps) * 65 + (i / taps)];
}

// small weight tensors (few (co, ci-chunk) blocks): one thread per element, strided write into the torch layout
__global__ void __launch_bounds__(256) wgrad_reduce_small_kernel(const float* __restrict__ part, int nsplit, float* __restrict__ dw, int Cout,
                                                                 int Cin, int taps, int Cin_total, int ci_off) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const long long plane = (long long)Cout * taps * Cin;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= plane) return;
    const int ci = (int)(i % Cin);
    const int tap = (int)((i / Cin) % taps);
    const int co = (int)(i / ((long long)Cin * taps));
    const float* src = part + i;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 4 <= nsplit; p += 4) {
        s0 += src[(size_t)p * plane]; s1 += src[(size_t)(p + 1) * plane];
        s2 += src[(size_t)(p + 2) * plane]; s3 += src[(size_t)(p + 3) * plane];
    }
    for (; p < nsplit; ++p) s0 += src[(size_t)p * plane];
    dw[((long long)co * Cin_total + ci_off + ci) * taps + tap] = (s0 + s1) + (s2 + s3);
}

}  // namespace mmad

using namespace mmad;
#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() do { MMAD_CUDA(cudaGetLastError()); count_launch(); return MMAD_OK; } while (0)

extern "C" {

int mmad_conv3d_prep_weights(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int taps, void* stream) {
    MMAD_CHECK_ARG(w && (w_fwd || w_dgrad) && Cout > 0 && Cin > 0 && taps > 0, "prep_weights: bad argument");
    MMAD_CHECK_ARG(w_fwd, "prep_weights: the forward layout is required (the dgrad layout is derived from it)");
    launch_pdl(prep_weights_fwd_kernel, dim3((Cin + 63) / 64, Cout), dim3(64), 64 * (taps + 1) * sizeof(float), ST, w, (__nv_bfloat16*)w_fwd, Cin, taps);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    if (w_dgrad) {
        launch_pdl(prep_weights_dgrad_kernel, dim3((Cin + 31) / 32, (Cout + 31) / 32, taps), dim3(256), 0, ST, (const __nv_bfloat16*)w_fwd,
                                                                                              (__nv_bfloat16*)w_dgrad, Cout, Cin, taps);
        MMAD_CUDA(cudaGetLastError());
        count_launch();
    }
    return MMAD_OK;
}
// Batched mmad_conv3d_prep_weights: n layers, two launches per 64 layers.  w / w_fwd / w_dgrad: host arrays of n device pointers
// (w_dgrad[i] NULL = no dgrad layout for layer i); cout / cin / taps: host arrays of n ints.
int mmad_conv3d_prep_weights_batched(int n, const void* const* w, void* const* w_fwd, void* const* w_dgrad, const int* cout,
                                     const int* cin, const int* taps, void* stream) {
    MMAD_CHECK_ARG(n > 0 && w && w_fwd && w_dgrad && cout && cin && taps, "prep_weights_batched: bad argument");
    for (int base = 0; base < n; base += kPrepBatch) {
        const int m = std::min(kPrepBatch, n - base);
        PrepBatch fb = {}, db = {};
        int max_taps = 1, nd = 0;
        fb.n = m;
        for (int i = 0; i < m; ++i) {
            const int k = base + i;
            MMAD_CHECK_ARG(w[k] && w_fwd[k] && cout[k] > 0 && cin[k] > 0 && taps[k] > 0 && taps[k] <= 343, "prep_weights_batched: bad layer");
            fb.d[i] = PrepDesc{(const float*)w[k], (__nv_bfloat16*)w_fwd[k], (__nv_bfloat16*)w_dgrad[k], cout[k], cin[k], taps[k]};
            fb.start[i + 1] = fb.start[i] + ((cin[k] + 63) / 64) * cout[k];
            max_taps = std::max(max_taps, taps[k]);
            if (w_dgrad[k]) {
                db.d[nd] = fb.d[i];
                db.start[nd + 1] = db.start[nd] + ((cin[k] + 31) / 32) * ((cout[k] + 31) / 32) * taps[k];
                ++nd;
            }
        }
        db.n = nd;
        const size_t smem = 64 * (size_t)(max_taps + 1) * sizeof(float);
        if (smem > 48 * 1024) {
            static DevOnce attr_done;
            if (attr_done.need()) MMAD_CUDA(cudaFuncSetAttribute(prep_weights_fwd_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
        launch_pdl(prep_weights_fwd_batched_kernel, dim3(fb.start[m]), dim3(64), smem, ST, fb);
        MMAD_CUDA(cudaGetLastError());
        count_launch();
        if (nd) {
            launch_pdl(prep_weights_dgrad_batched_kernel, dim3(db.start[nd]), dim3(256), 0, ST, db);
            MMAD_CUDA(cudaGetLastError());
            count_launch();
        }
    }
    return MMAD_OK;
}
int mmad_conv3d_prep_weights_s2(const float* w, void* w_phases, int Cdy, int Cdx, void* stream) {
    MMAD_CHECK_ARG(w && w_phases && Cdy > 0 && Cdx > 0, "prep_weights_s2: bad argument");
    launch_pdl(prep_weights_s2_kernel, dim3(grid_for(27ll * Cdx * Cdy, 256, 1184)), dim3(256), 0, ST, w, (__nv_bfloat16*)w_phases, Cdy, Cdx);
    LAUNCH_OK();
}
int mmad_bn_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                     float* shift, void* stream) {
    MMAD_CHECK_ARG(partials && gamma && beta && mean && invstd && scale && shift && C > 0 && count > 0, "bn_finalize: bad argument");
    launch_pdl(bn_finalize_kernel, dim3((C + 3) / 4), dim3(128), 0, ST, partials, nparts, C, count, gamma, beta, eps, momentum, running_mean, running_var,
                                                   mean, invstd, scale, shift);
    LAUNCH_OK();
}
int mmad_bn_eval_params(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                        float* mean, float* invstd, float* scale, float* shift, void* stream) {
    MMAD_CHECK_ARG(gamma && beta && running_mean && running_var && mean && invstd && scale && shift, "bn_eval_params: null pointer");
    launch_pdl(bn_eval_kernel, dim3((C + 127) / 128), dim3(128), 0, ST, C, gamma, beta, running_mean, running_var, eps, mean, invstd, scale, shift);
    LAUNCH_OK();
}
int mmad_bn_apply_ex(const void* x, const float* scale, const float* shift, const void* res, const float* rscale, const float* rshift,
                     int relu, void* out_bf16, int64_t out_ld, float* out_f32, int64_t rows, int C, void* stream) {
    MMAD_CHECK_ARG(x && scale && shift && (out_bf16 || out_f32) && C % 8 == 0 && rows > 0, "bn_apply: bad argument");
    MMAD_CHECK_ARG(out_ld == 0 || (out_ld >= C && out_ld % 8 == 0), "bn_apply: out_ld must be 0 (dense) or >= C and a multiple of 8");
    const long long nvec = rows * (C / 8);
    const int ld8 = (int)((out_ld ? out_ld : C) / 8);
    const int grid = grid_for_channels(nvec, C / 8, sm_count() * 8);
    if (relu) launch_pdl(bn_apply_kernel<true>, dim3(grid), dim3(256), 0, ST, (const uint4*)x, scale, shift, (const uint4*)res, rscale, rshift, (uint4*)out_bf16, (float4*)out_f32, nvec, C, ld8);
    else launch_pdl(bn_apply_kernel<false>, dim3(grid), dim3(256), 0, ST, (const uint4*)x, scale, shift, (const uint4*)res, rscale, rshift, (uint4*)out_bf16, (float4*)out_f32, nvec, C, ld8);
    LAUNCH_OK();
}
int mmad_bn_apply(const void* x, const float* scale, const float* shift, const void* res, const float* rscale, const float* rshift,
                  int relu, void* out_bf16, float* out_f32, int64_t rows, int C, void* stream) {
    return mmad_bn_apply_ex(x, scale, shift, res, rscale, rshift, relu, out_bf16, 0, out_f32, rows, C, stream);
}
// number of block partials mmad_bn_bwd_reduce writes: float[n][C][2]
int mmad_bn_bwd_partials(int64_t rows) { return (int)std::max<long long>(1, std::min<long long>(rows / 64, sm_count())); }
int mmad_bn_bwd_reduce(const void* dy_bf16, const float* dy_f32, const void* dy2, const void* mask, const void* x, const float* mean,
                       const float* invstd, const float* mask_scale, const float* mask_shift, void* g_out, float* partials, int64_t rows,
                       int C, void* stream) {
    MMAD_CHECK_ARG((dy_bf16 || dy_f32) && x && mean && invstd && partials && rows > 0, "bn_bwd_reduce: bad argument");
    MMAD_CHECK_ARG(C % 64 == 0 && C <= 2048 && 256 % (C / 8) == 0, "bn_bwd_reduce: C must be 64, 128, 256, 512, 1024 or 2048");
    const int grid = mmad_bn_bwd_partials(rows);
    static int tma_mode = -1;                          // TMA-staged kernel for bf16 gradients: on unless MMAD_BN_TMA=0
    if (tma_mode < 0) { const char* e = getenv("MMAD_BN_TMA"); tma_mode = e ? atoi(e) : 1; }
    if (dy_bf16 && tma_mode) {
        const int smem = kBnStages * 4 * kBnTileVec * 16 + 2 * kBnStages * 8;
        static DevOnce attr_done;
        if (attr_done.need()) {
            MMAD_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        }
        launch_pdl(bn_bwd_reduce_tma_kernel, dim3(grid), dim3(kBnConsumers + 32), smem, ST, (const uint4*)dy_bf16, (const uint4*)dy2, (const uint4*)mask,
                                                                         (const uint4*)x, mean, invstd, mask_scale, mask_shift,
                                                                         (uint4*)g_out, partials, (long long)rows * (C / 8), C);
        LAUNCH_OK();
    }
    launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(256), 256 * 16 * sizeof(float), ST, (const uint4*)dy_bf16, (const float4*)dy_f32, (const uint4*)dy2,
                                                                     (const uint4*)mask, (const uint4*)x, mean, invstd, mask_scale, mask_shift,
                                                                     (uint4*)g_out, partials, rows, C);
    LAUNCH_OK();
}
int mmad_bn_bwd_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* mean,
                         const float* invstd, int training, float* dgamma, float* dbeta, float* coef, void* stream) {
    MMAD_CHECK_ARG(partials && gamma && mean && invstd && coef && C > 0, "bn_bwd_finalize: bad argument");
    launch_pdl(bn_bwd_finalize_kernel, dim3((C + 3) / 4), dim3(128), 0, ST, partials, nparts, C, count, gamma, mean, invstd, training, dgamma, dbeta, coef);
    LAUNCH_OK();
}
int mmad_bn_bwd_apply_ex(const void* g, const void* x, const float* coef, const float* mask_scale, const float* mask_shift, void* dx,
                         int64_t rows, int C, void* stream) {
    MMAD_CHECK_ARG(g && x && coef && dx && C % 8 == 0, "bn_bwd_apply: bad argument");
    MMAD_CHECK_ARG((mask_scale == nullptr) == (mask_shift == nullptr), "bn_bwd_apply: mask scale and shift come together");
    const long long nvec = rows * (C / 8);
    launch_pdl(bn_bwd_apply_kernel, dim3(grid_for_channels(nvec, C / 8, sm_count() * 8)), dim3(256), 0, ST, (const uint4*)g, (const uint4*)x, coef,
               mask_scale, mask_shift, (uint4*)dx, nvec, C);
    LAUNCH_OK();
}
int mmad_bn_bwd_apply(const void* g, const void* x, const float* coef, void* dx, int64_t rows, int C, void* stream) {
    return mmad_bn_bwd_apply_ex(g, x, coef, nullptr, nullptr, dx, rows, C, stream);
}
int mmad_maxpool3d_fwd(const void* x, void* y, void* idx, int N, int D, int H, int W, int C, void* stream) {
    MMAD_CHECK_ARG(x && y && idx && C % 8 == 0, "maxpool3d_fwd: bad argument");
    const int Do = (D - 1) / 2 + 1, Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)N * Do * Ho * Wo * (C / 8);
    launch_pdl(maxpool3d_fwd_kernel, dim3(grid_for(total, 256, sm_count() * 16)), dim3(256), 0, ST, (const uint4*)x, (uint4*)y, (uint2*)idx, N, D, H, W, C, Do, Ho, Wo);
    LAUNCH_OK();
}
int mmad_maxpool3d_bwd(const void* dy, const void* idx, void* dx, int N, int D, int H, int W, int C, void* stream) {
    MMAD_CHECK_ARG(dy && dx && idx && C % 8 == 0, "maxpool3d_bwd: bad argument");
    const int Do = (D - 1) / 2 + 1, Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int cv = C / 8;
    MMAD_CHECK_ARG(cv <= 512 && (long long)N * D < (1ll << 31), "maxpool3d_bwd: unsupported size");
    int threads = 512;
    while (threads > 32 && (threads % cv != 0 || threads / cv > W)) threads >>= 1;
    MMAD_CHECK_ARG(threads % cv == 0, "maxpool3d_bwd: C/8 must divide a power of two <= 512");
    static int march_mode = -1;                        // marching kernel: on unless MMAD_POOL_MARCH=0
    if (march_mode < 0) { const char* e = getenv("MMAD_POOL_MARCH"); march_mode = e ? atoi(e) : 1; }
    const long long slot = 3ll * Wo * cv * 24;
    if (march_mode && 512 % cv == 0 && kPoolSlots * slot + 64 <= 72 * 1024 && D >= 4) {
        const int smem = (int)(kPoolSlots * slot) + 64;
        static DevOnce attr_done;
        if (attr_done.need()) {
            MMAD_CUDA(cudaFuncSetAttribute(maxpool3d_bwd_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        }
        const int hgroups = (H + 3) / 4;
        int dsplit = 1;                                // enough blocks for ~3 per SM
        while ((long long)N * hgroups * dsplit < 444 && D / (dsplit * 2) >= 4) dsplit *= 2;
        launch_pdl(maxpool3d_bwd_march_kernel, dim3((unsigned)(N * hgroups * dsplit)), dim3(512), smem, ST, (const uint4*)dy, (const uint2*)idx, (uint4*)dx, N, D,
                                                                                        H, W, C, Do, Ho, Wo, dsplit);
        LAUNCH_OK();
    }
    const int hrows = 4;
    launch_pdl(maxpool3d_bwd_kernel, dim3((unsigned)(N * D), (unsigned)((H + hrows - 1) / hrows)), dim3(threads), 0, ST, 
        (const uint4*)dy, (const uint2*)idx, (uint4*)dx, N, D, H, W, C, Do, Ho, Wo, hrows);
    LAUNCH_OK();
}
int mmad_stem_bn_relu_maxpool_fwd(const void* c, const float* scale, const float* shift, void* y, void* idx, int N, int D, int H, int W,
                                  int C, void* stream) {
    MMAD_CHECK_ARG(c && scale && shift && y && idx && C % 8 == 0, "stem_bn_relu_maxpool_fwd: bad argument");
    const int Do = (D - 1) / 2 + 1, Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)N * Do * Ho * Wo * (C / 8);
    MMAD_CHECK_ARG(total < (1ll << 32), "stem_bn_relu_maxpool_fwd: tensor too large for 32-bit indexing");
    static int march_mode = -1;                        // marching kernel: on unless MMAD_STEM_POOL_MARCH=0
    if (march_mode < 0) { const char* e = getenv("MMAD_STEM_POOL_MARCH"); march_mode = e ? atoi(e) : 1; }
    const int cv = C / 8;
    const long long slot = 5ll * W * cv * 16;
    if (march_mode && 512 % cv == 0 && kSpSlots * slot + 64 <= 220 * 1024 && Do >= 2) {
        const int smem = (int)(kSpSlots * slot) + 64;
        static DevOnce attr_done;
        if (attr_done.need()) {
            MMAD_CUDA(cudaFuncSetAttribute(stem_pool_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        }
        const int hgroups = (Ho + 1) / 2;
        int dsplit = 1;                                // several blocks per SM over the run, at least 4 pooled slices each
        while ((long long)N * hgroups * dsplit < sm_count() * 6 && Do / (dsplit * 2) >= 4) dsplit *= 2;
        launch_pdl(stem_pool_march_kernel, dim3((unsigned)(N * hgroups * dsplit)), dim3(512), smem, ST, (const uint4*)c, scale, shift, (uint4*)y,
                   (uint2*)idx, N, D, H, W, C, Do, Ho, Wo, dsplit);
        LAUNCH_OK();
    }
    launch_pdl(stem_bn_relu_maxpool_fwd_kernel, dim3(grid_for(total, 256, sm_count() * 16)), dim3(256), 0, ST, (const uint4*)c, scale, shift, (uint4*)y, (uint2*)idx, N, D, H,
                                                                                    W, C, Do, Ho, Wo);
    LAUNCH_OK();
}
int mmad_upsample_zero2(const void* x, void* y, int N, int Dx, int Hx, int Wx, int Dy, int Hy, int Wy, int C, void* stream) {
    MMAD_CHECK_ARG(x && y && C % 8 == 0, "upsample_zero2: bad argument");
    const long long total = (long long)N * Dy * Hy * Wy * (C / 8);
    launch_pdl(upsample_zero2_kernel, dim3(grid_for(total, 256, sm_count() * 16)), dim3(256), 0, ST, (const uint4*)x, (uint4*)y, N, Dx, Hx, Wx, Dy, Hy, Wy, C);
    LAUNCH_OK();
}
int mmad_ncs_f32_to_nsc_bf16(const float* x, void* y, int N, int C, int64_t S, void* stream) {
    MMAD_CHECK_ARG(x && y && N > 0 && C > 0 && S > 0, "ncs_f32_to_nsc_bf16: bad argument");
    dim3 grid((unsigned)((S + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
    launch_pdl(ncs_f32_to_nsc_bf16_kernel, dim3(grid), dim3(256), 0, ST, x, (__nv_bfloat16*)y, C, S);
    LAUNCH_OK();
}
// partials [nsplit][Cout][taps][Cin] -> dw[:, ci_off : ci_off + Cin] of a torch-layout (Cout, Cin_total, taps) fp32 tensor
int mmad_wgrad_reduce_ex(const float* partials, int nsplit, float* dw, int Cout, int Cin, int taps, int Cin_total, int ci_off, void* stream) {
    MMAD_CHECK_ARG(partials && dw && nsplit > 0 && ci_off >= 0 && ci_off + Cin <= Cin_total, "wgrad_reduce: bad argument");
    if ((long long)((Cin + 63) / 64) * Cout < 1024) {
        const long long plane = (long long)Cout * taps * Cin;
        launch_pdl(wgrad_reduce_small_kernel, dim3((unsigned)((plane + 255) / 256)), dim3(256), 0, ST, partials, nsplit, dw, Cout, Cin, taps, Cin_total, ci_off);
    } else {
        launch_pdl(wgrad_reduce_kernel, dim3((Cin + 63) / 64, Cout), dim3(256), taps * 65 * sizeof(float), ST, partials, nsplit, dw, Cout, Cin, taps, Cin_total, ci_off);
    }
    LAUNCH_OK();
}
int mmad_wgrad_reduce(const float* partials, int nsplit, float* dw, int Cout, int Cin, int taps, void* stream) {
    return mmad_wgrad_reduce_ex(partials, nsplit, dw, Cout, Cin, taps, Cin, 0, stream);
}

}  // extern "C"
