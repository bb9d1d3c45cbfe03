// Stem convolution conv1 (1 -> 64 channels, 7x7x7, stride 2, padding 3: /root/reference/models/resnet.py:126-132) as an
// implicit GEMM on tcgen05 WITHOUT a materialised im2col matrix, forward and weight gradient.
//
// Space-to-depth: a stride-2 7^3 convolution of one channel is a stride-1 4^3 convolution of the 8 "phase" channels
//   xs[n][jd][jh][jw][pd*4+ph*2+pw] = x[n][2jd+pd-3][2jh+ph-3][2jw+pw-3]          (zero outside the volume)
//   y[n][do][ho][wo][co] = sum_{kd,kh,kw,p} xs[n][do+kd][ho+kh][wo+kw][p] * wk[co][kd][kh][kw][p],   t = 2k+p, wk = 0 where t = 7
// so K = 4*4*4*8 = 512 and one im2col row segment (kd, kh fixed; kw, p running) is 64 CONTIGUOUS bytes of xs starting
// at voxel (do+kd, ho+kh, wo).  A tensor map with OVERLAPPING rows (row stride 16 bytes = one voxel, row length 64
// bytes) lets TMA write those segments straight into the SWIZZLE_64B K-major operand layout:
//   box {32 elements, 8 wo, 4 ho, 7 d, 4 kh}  ->  smem [kh][d 0..6][ho][wo] x 64 B   (56 KB, ONE box per 128-voxel tile)
// The operand of tap (kd, kh) is the 128 rows starting at slice d = kd of atom kh: uniform 8-row group stride, so the same
// box serves all 16 (kd, kh) K-atoms.  Forward: A = that box (K-major), B = the packed weights, resident in smem.
// Weight gradient: the SAME box read MN-major (rows = voxels = the reduction index) against dY; the four kh atoms of a
// kd form one 128-row accumulator block, so 4 blocks x 64 output channels live in 256 TMEM columns.
#include "tc_common.cuh"

#include <algorithm>

namespace mmad {

constexpr int kStemThreads = 256;
constexpr int kStemK = 512;                       // 4*4*4 taps x 8 phases
constexpr int kStemSlab = 32 * 64;                // one d slice of an atom: 32 (ho, wo) rows x 64 B
constexpr int kStemAtom = 7 * kStemSlab;          // one kh atom: 7 d slices
constexpr int kStemABox = 4 * kStemAtom;          // 57344 B
constexpr int kStemBBytes = 64 * kStemK * 2;      // packed weights: 16 atoms x [64 co x 64 B]
constexpr int kStemOutBytes = 128 * 128;          // epilogue staging: 128 voxels x 64 channels bf16
constexpr int kStemDyBox = 128 * 128;             // wgrad: 128 voxels x 64 channels of dY

__device__ unsigned long long g_mma_flops_stem;            // executed tensor-core flops of this file's kernels (tc_common.cuh)
long long mma_flops_stem() {
    unsigned long long v = 0;
    return cudaMemcpyFromSymbol(&v, g_mma_flops_stem, sizeof(v)) == cudaSuccess ? (long long)v : -1;
}

struct StemGeom {
    int N, D, H, W;               // input volume (one channel)
    int Do, Ho, Wo;               // conv1 output
    int Ds, Hs, Ws;               // space-to-depth grid: Do + 3 etc.
    int tiles_w, tiles_h, tiles_d, m_tiles;   // 8 x 4 x 4 output tiles
};

__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (4ull << 61);
}

// ---------------------------------------------------------------------------------------------------------------
// pack kernels
// ---------------------------------------------------------------------------------------------------------------
// x fp32 [N][D][H][W] -> xs bf16 [N][Ds][Hs][Ws][8]; one thread per s2d voxel (one 16-byte store)
__global__ void __launch_bounds__(256) stem_s2d_pack_kernel(const float* __restrict__ x, uint4* __restrict__ xs, StemGeom g) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const long long total = (long long)g.N * g.Ds * g.Hs * g.Ws;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int jw = (int)(r % g.Ws); r /= g.Ws;
        const int jh = (int)(r % g.Hs); r /= g.Hs;
        const int jd = (int)(r % g.Ds); r /= g.Ds;
        const float* xn = x + (size_t)r * g.D * g.H * g.W;
        float v[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int d = 2 * jd + (p >> 2) - 3, h = 2 * jh + ((p >> 1) & 1) - 3, w = 2 * jw + (p & 1) - 3;
            v[p] = (d >= 0 && d < g.D && h >= 0 && h < g.H && w >= 0 && w < g.W) ? __ldg(xn + ((size_t)d * g.H + h) * g.W + w) : 0.f;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        xs[i] = o;
    }
}

// K index of the packed weight / gradient layouts for filter tap (td, th, tw)
__host__ __device__ __forceinline__ int stem_k_index(int td, int th, int tw) {
    return ((((td >> 1) * 4 + (th >> 1)) * 4 + (tw >> 1)) << 3) + ((td & 1) << 2) + ((th & 1) << 1) + (tw & 1);
}

// w fp32 [64][343] -> wk bf16 [64][512]
__global__ void __launch_bounds__(512) stem_s2d_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wk) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int co = blockIdx.x, K = threadIdx.x;
    const int p = K & 7, kw = (K >> 3) & 3, kh = (K >> 5) & 3, kd = K >> 7;
    const int td = 2 * kd + (p >> 2), th = 2 * kh + ((p >> 1) & 1), tw = 2 * kw + (p & 1);
    const float v = (td < 7 && th < 7 && tw < 7) ? w[co * 343 + (td * 7 + th) * 7 + tw] : 0.f;
    wk[co * kStemK + K] = __float2bfloat16(v);
}

// partials fp32 [nsplit][64][512] -> dw fp32 [64][343] (torch layout)
__global__ void __launch_bounds__(384) stem_s2d_wgrad_reduce_kernel(const float* __restrict__ partials, int nsplit, float* __restrict__ dw) {
    pdl_launch_dependents();
    pdl_wait();                                        // see launch_pdl (common.cuh)
    const int co = blockIdx.x, t = threadIdx.x;
    if (t >= 343) return;
    const int K = stem_k_index(t / 49, (t / 7) % 7, t % 7);
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partials[((size_t)sp * 64 + co) * kStemK + K];
    dw[co * 343 + t] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// forward: y = conv1(x) as bf16 NDHWC (+ per-CTA BatchNorm statistic partials), persistent CTAs over 8x4x4 tiles
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStemThreads, 1)
stem_conv_s2d_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const StemGeom g, float* __restrict__ stats_partials) {
    pdl_launch_dependents();
    constexpr int S = 2;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 0, 0);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const uint32_t b0 = base;                                 // weights
    const uint32_t a0 = base + kStemBBytes;                   // S input boxes
    const uint32_t out0 = a0 + S * kStemABox;                 // epilogue staging
    unsigned char* tail = sm + kStemBBytes + S * kStemABox + kStemOutBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);       // full[S], empty[S], bfull, tfull[2], tempty[2]
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, bfull = empty0 + 8 * S, tfull0 = bfull + 8, tempty0 = tfull0 + 16;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 5);
    float* st_sum = reinterpret_cast<float*>(tmem_ptr_s + 4);   // [2][64] (two row halves)
    float* st_sq = st_sum + 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(bfull, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        mbar_fence_init();
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_s), 128);
    for (int i = threadIdx.x; i < 256; i += kStemThreads) st_sum[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    if (warp == 0) {
        // ============================ TMA producer: the weights once, then one input box per tile ============================
        {
            if (elect_one()) {
                mbar_arrive_expect_tx(bfull, kStemBBytes);
                tma_load_3d(b0, &tmB, bfull, 0, 0, 0);
            }
            __syncwarp();
            uint32_t s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
                int r = tile;
                const int wt = r % g.tiles_w; r /= g.tiles_w;
                const int ht = r % g.tiles_h; r /= g.tiles_h;
                const int dt = r % g.tiles_d; r /= g.tiles_d;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full0 + 8 * s, kStemABox);
                    tma_load_5d(a0 + s * kStemABox, &tmA, full0 + 8 * s, 0, wt * 8, ht * 4, r * g.Ds + dt * 4, 0);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        {
            mbar_wait(bfull, 0);
            uint32_t s = 0, ph = 0, it = 0, nmma = 0;
            for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(tempty0 + 8 * acc, aph ^ 1);
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t sa = a0 + s * kStemABox;
                if (elect_one()) {
#pragma unroll
                    for (int kd = 0; kd < 4; ++kd)
#pragma unroll
                        for (int kh = 0; kh < 4; ++kh) {
                            const uint64_t adesc = umma_desc_sw64(sa + kh * kStemAtom + kd * kStemSlab, 16, 512);
                            const uint64_t bdesc = umma_desc_sw64(b0 + (kd * 4 + kh) * 4096, 16, 512);
#pragma unroll
                            for (int j = 0; j < 2; ++j)           // 2 x K16 inside the 32-wide (64-byte) swizzled row
                                umma_bf16(tmem_base + acc * 64, adesc + 2 * j, bdesc + 2 * j, IDESC, (kd | kh | j) ? 1u : 0u);
                        }
                    nmma += 32;
                    umma_commit(empty0 + 8 * s);
                    umma_commit(tfull0 + 8 * acc);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1; }
            }
            if (elect_one()) mma_count_flush(&g_mma_flops_stem, nmma, 2u * 128u * 64u * 16u);
        }
    } else if (warp >= 4) {
        // ============================ epilogue: TMEM -> bf16 -> smem -> TMA store (+ BN statistics) ============================
        const int ew = warp - 4;
        const int et = threadIdx.x - 128;
        const int row = ew * 32 + lane;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            int r = tile;
            const int wt = r % g.tiles_w; r /= g.tiles_w;
            const int ht = r % g.tiles_h; r /= g.tiles_h;
            const int dt = r % g.tiles_d; r /= g.tiles_d;
            const int n = r;
            const int w0 = wt * 8, h0 = ht * 4, d0 = dt * 4;
            const int vw = min(8, g.Wo - w0), vh = min(4, g.Ho - h0), vd = min(4, g.Do - d0);

            mbar_wait(tfull0 + 8 * acc, aph);
            tc_fence_after();
            if (et == 0) tma_store_wait_read<0>();                // the previous store has finished reading the staging buffer
            named_bar_sync(2, 128);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 64 + half * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t p0 = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
                    const uint32_t p1 = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                    const uint32_t p2 = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                    const uint32_t p3 = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                    const uint32_t chunk = (uint32_t)(half * 4 + q) ^ (uint32_t)(row & 7);      // SWIZZLE_128B
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(out0 + row * 128 + chunk * 16), "r"(p0), "r"(p1),
                                 "r"(p2), "r"(p3)
                                 : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);       // accumulator drained: the next tile's MMAs may overwrite it
            fence_proxy_async_smem();
            named_bar_sync(2, 128);
            if (et == 0) {
                tma_store_5d(&tmC, out0, 0, w0, h0, d0, n);
                tma_store_commit();
            }
            if (stats_partials) {
                const int c = et & 63, hf = et >> 6;
                float sum = 0.f, sq = 0.f;
                const unsigned char* obp = sm + (out0 - base);
                for (int rr = hf * 64; rr < hf * 64 + 64; ++rr) {
                    const int wi = rr & 7, hi = (rr >> 3) & 3, di = rr >> 5;
                    if (wi < vw && hi < vh && di < vd) {
                        const uint32_t off = rr * 128 + (((uint32_t)(c >> 3) ^ (uint32_t)(rr & 7)) << 4) + (c & 7) * 2;
                        const float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(obp + off));
                        sum += x;
                        sq += x * x;
                    }
                }
                st_sum[hf * 64 + c] += sum;
                st_sq[hf * 64 + c] += sq;
            }
        }
        if (et == 0) tma_store_wait<0>();
        if (stats_partials) {
            named_bar_sync(2, 128);
            if (et < 64) {
                stats_partials[((size_t)blockIdx.x * 64 + et) * 2 + 0] = st_sum[et] + st_sum[64 + et];
                stats_partials[((size_t)blockIdx.x * 64 + et) * 2 + 1] = st_sq[et] + st_sq[64 + et];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: partials[cta][co][K] = sum over this CTA's 128-voxel chunks of dY[v][co] * A[v][K]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStemThreads, 1)
stem_wgrad_s2d_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmDY, const StemGeom g,
                      float* __restrict__ partials) {
    pdl_launch_dependents();
    constexpr int S = 3;
    constexpr uint32_t STAGE = kStemABox + kStemDyBox;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 1, 1);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)S * STAGE);   // full[S], empty[S], tfull
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull = empty0 + 8 * S;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_begin = (int)((long long)g.m_tiles * blockIdx.x / gridDim.x), c_end = (int)((long long)g.m_tiles * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmDY);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_s), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers: chunk i is issued by producer i % 3 (S == 3) ============================
        {
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t s = 0, ph = 0;
            for (int c = c_begin; c < c_end; ++c) {
                if (s == me) {
                    int r = c;
                    const int wt = r % g.tiles_w; r /= g.tiles_w;
                    const int ht = r % g.tiles_h; r /= g.tiles_h;
                    const int dt = r % g.tiles_d; r /= g.tiles_d;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full0 + 8 * s, STAGE);
                        const uint32_t sb = base + s * STAGE;
                        tma_load_5d(sb, &tmA, full0 + 8 * s, 0, wt * 8, ht * 4, r * g.Ds + dt * 4, 0);
                        tma_load_5d(sb + kStemABox, &tmDY, full0 + 8 * s, 0, wt * 8, ht * 4, dt * 4, r);
                    }
                    __syncwarp();
                }
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        {
            uint32_t s = 0, ph = 0, nmma = 0;
            for (int c = c_begin; c < c_end; ++c) {
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t sb = base + s * STAGE;
                if (elect_one()) {
                nmma += 32;
#pragma unroll
                for (int kd = 0; kd < 4; ++kd) {
                    // A, MN-major SWIZZLE_64B: 64-byte rows are voxels (K), 8-row groups 512 B apart, the four 32-wide kh atoms one
                    // atom apart; B = dY, MN-major SWIZZLE_128B: 128-byte rows are voxels, one 64-wide atom
                    const uint64_t adesc = umma_desc_sw64(sb + kd * kStemSlab, kStemAtom, 512);
                    const uint64_t bdesc = umma_desc_sw128(sb + kStemABox, kStemDyBox, 1024);
#pragma unroll
                    for (int j = 0; j < 8; ++j)                   // K16 = 16 voxel rows
                        umma_bf16(tmem_base + kd * 64, adesc + 64 * j, bdesc + 128 * j, IDESC, (c > c_begin || j) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);
                }
                __syncwarp();
                if (++s == S) { s = 0; ph ^= 1; }
            }
            if (elect_one()) { umma_commit(tfull); mma_count_flush(&g_mma_flops_stem, nmma, 2u * 128u * 64u * 16u); }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ============================ epilogue: TMEM -> fp32 partials [cta][co][K] ============================
        const int ew = warp - 4;
        const int m = ew * 32 + lane;                             // accumulator row: K index inside the kd block
        mbar_wait(tfull, 0);
        tc_fence_after();
        float* out = partials + (size_t)blockIdx.x * 64 * kStemK;
        for (int kd = 0; kd < 4; ++kd)
            for (int n0 = 0; n0 < 64; n0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + kd * 64 + n0, v);
                tmem_ld_wait();
                if (c_end == c_begin) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) out[(size_t)(n0 + j) * kStemK + kd * 128 + m] = __uint_as_float(v[j]);
            }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

static int stem_geom(StemGeom& g, int N, int D, int H, int W) {
    g.N = N; g.D = D; g.H = H; g.W = W;
    g.Do = (D - 1) / 2 + 1; g.Ho = (H - 1) / 2 + 1; g.Wo = (W - 1) / 2 + 1;
    g.Ds = g.Do + 3; g.Hs = g.Ho + 3; g.Ws = g.Wo + 3;
    g.tiles_w = (g.Wo + 7) / 8; g.tiles_h = (g.Ho + 3) / 4; g.tiles_d = (g.Do + 3) / 4;
    const long long t = (long long)N * g.tiles_w * g.tiles_h * g.tiles_d;
    if (t > 0x7fffffffLL || (long long)N * g.Ds > 0x7fffffffLL) return -1;
    g.m_tiles = (int)t;
    return 0;
}

// input boxes: overlapping 64-byte rows, see the header comment
static int stem_input_map(CUtensorMap* tm, const void* xs, const StemGeom& g) {
    const uint64_t dims[5] = {32, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.N * g.Ds, 4};
    const uint64_t str[4] = {16, (uint64_t)g.Ws * 16, (uint64_t)g.Hs * g.Ws * 16, (uint64_t)g.Ws * 16};
    const uint32_t box[5] = {32, 8, 4, 7, 4};
    const uint32_t es[5] = {1, 1, 1, 1, 1};
    return make_tmap_bf16_swz(tm, xs, 5, dims, str, box, es, 64);
}

}  // namespace mmad

using namespace mmad;

#define ST ((cudaStream_t)stream)
#define LAUNCH_OK() do { MMAD_CUDA(cudaGetLastError()); count_launch(); return MMAD_OK; } while (0)

extern "C" {

// elements of the space-to-depth tensor for an N x 1 x D x H x W input (bf16, [N][Ds][Hs][Ws][8])
int64_t mmad_stem_s2d_elems(int N, int D, int H, int W) {
    StemGeom g;
    if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || stem_geom(g, N, D, H, W)) return -1;
    return (int64_t)N * g.Ds * g.Hs * g.Ws * 8;
}

int mmad_stem_s2d_pack(const float* x, void* xs, int N, int D, int H, int W, void* stream) {
    MMAD_CHECK_ARG(x && xs && N > 0 && D > 0 && H > 0 && W > 0, "stem_s2d_pack: bad argument");
    StemGeom g;
    MMAD_CHECK_ARG(stem_geom(g, N, D, H, W) == 0, "stem_s2d_pack: volume too large");
    const long long total = (long long)N * g.Ds * g.Hs * g.Ws;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    launch_pdl(stem_s2d_pack_kernel, dim3(grid), dim3(256), 0, ST, x, (uint4*)xs, g);
    LAUNCH_OK();
}

int mmad_stem_s2d_prep_weights(const float* w, void* wk, void* stream) {
    MMAD_CHECK_ARG(w && wk, "stem_s2d_prep_weights: null pointer");
    launch_pdl(stem_s2d_weights_kernel, dim3(64), dim3(512), 0, ST, w, (__nv_bfloat16*)wk);
    LAUNCH_OK();
}

// number of per-CTA statistic partials mmad_stem_s2d_fwd writes (== its grid size)
int mmad_stem_s2d_stats_partials(int N, int D, int H, int W) {
    StemGeom g;
    if (stem_geom(g, N, D, H, W)) return -1;
    return std::min(g.m_tiles, sm_count());
}

int mmad_stem_s2d_fwd(const void* xs, const void* wk, void* y, float* stats_partials, int N, int D, int H, int W, void* stream) {
    MMAD_CHECK_ARG(xs && wk && y && N > 0 && D > 0 && H > 0 && W > 0, "stem_s2d_fwd: bad argument");
    StemGeom g;
    MMAD_CHECK_ARG(stem_geom(g, N, D, H, W) == 0, "stem_s2d_fwd: volume too large");
    CUtensorMap tmA, tmB, tmC;
    int rc = stem_input_map(&tmA, xs, g);
    if (rc) return rc;
    {
        const uint64_t dims[3] = {32, 64, 16};                 // [32 k][co][K atom]
        const uint64_t str[2] = {kStemK * 2, 64};
        const uint32_t box[3] = {32, 64, 16};
        const uint32_t es[3] = {1, 1, 1};
        rc = make_tmap_bf16_swz(&tmB, wk, 3, dims, str, box, es, 64);
        if (rc) return rc;
    }
    {
        const uint64_t dims[5] = {64, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
        const uint64_t str[4] = {128, (uint64_t)g.Wo * 128, (uint64_t)g.Ho * g.Wo * 128, (uint64_t)g.Do * g.Ho * g.Wo * 128};
        const uint32_t box[5] = {64, 8, 4, 4, 1};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        rc = make_tmap_bf16(&tmC, y, 5, dims, str, box, es);
        if (rc) return rc;
    }
    const int smem = 1024 + kStemBBytes + 2 * kStemABox + kStemOutBytes + 9 * 8 + 16 + 256 * 4;
    static DevOnce attr_done;
    if (attr_done.need()) {
        MMAD_CUDA(cudaFuncSetAttribute(stem_conv_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int grid = std::min(g.m_tiles, sm_count());
    launch_pdl(stem_conv_s2d_kernel, dim3(grid), dim3(kStemThreads), smem, ST, tmA, tmB, tmC, g, stats_partials);
    LAUNCH_OK();
}

// fp32 elements of the partials buffer of mmad_stem_s2d_wgrad ([nsplit][64][512]) and its split count
int64_t mmad_stem_s2d_wgrad_workspace(int N, int D, int H, int W, int* nsplit_out) {
    StemGeom g;
    if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || stem_geom(g, N, D, H, W)) return -1;
    const int ns = std::min(g.m_tiles, sm_count());
    if (nsplit_out) *nsplit_out = ns;
    return (int64_t)ns * 64 * kStemK;
}

int mmad_stem_s2d_wgrad(const void* xs, const void* dy, float* partials, int N, int D, int H, int W, void* stream) {
    MMAD_CHECK_ARG(xs && dy && partials && N > 0 && D > 0 && H > 0 && W > 0, "stem_s2d_wgrad: bad argument");
    StemGeom g;
    MMAD_CHECK_ARG(stem_geom(g, N, D, H, W) == 0, "stem_s2d_wgrad: volume too large");
    CUtensorMap tmA, tmDY;
    int rc = stem_input_map(&tmA, xs, g);
    if (rc) return rc;
    {
        const uint64_t dims[5] = {64, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
        const uint64_t str[4] = {128, (uint64_t)g.Wo * 128, (uint64_t)g.Ho * g.Wo * 128, (uint64_t)g.Do * g.Ho * g.Wo * 128};
        const uint32_t box[5] = {64, 8, 4, 4, 1};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        rc = make_tmap_bf16(&tmDY, dy, 5, dims, str, box, es);
        if (rc) return rc;
    }
    const int smem = 1024 + 3 * (kStemABox + kStemDyBox) + 7 * 8 + 32;
    static DevOnce attr_done;
    if (attr_done.need()) {
        MMAD_CUDA(cudaFuncSetAttribute(stem_wgrad_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int grid = std::min(g.m_tiles, sm_count());
    launch_pdl(stem_wgrad_s2d_kernel, dim3(grid), dim3(kStemThreads), smem, ST, tmA, tmDY, g, partials);
    LAUNCH_OK();
}

int mmad_stem_s2d_wgrad_reduce(const float* partials, int nsplit, float* dw, void* stream) {
    MMAD_CHECK_ARG(partials && dw && nsplit > 0, "stem_s2d_wgrad_reduce: bad argument");
    launch_pdl(stem_s2d_wgrad_reduce_kernel, dim3(64), dim3(384), 0, ST, partials, nsplit, dw);
    LAUNCH_OK();
}

}  // extern "C"
