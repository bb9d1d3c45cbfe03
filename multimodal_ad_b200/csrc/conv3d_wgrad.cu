// Conv3d weight gradient as an implicit GEMM on tcgen05 / TMEM (sm_100a).
//   dW[co][tap][ci] = sum over output voxels v of  dY[v][co] * X[v*stride + tap*dil - pad][ci]
// Replaces the weight-gradient half of torch.nn.Conv3d's backward for /root/reference/models/resnet.py convolutions.
//
// GEMM view: the reduction (K) runs over output voxels, so BOTH operands are "MN-major" for the tensor core:
//   A = X  shifted by the tap : 64 voxels x 128 (ci)   -> M = 128 rows of the accumulator (one tap x 128 ci, or, when
//                                                          Cin == 64, two taps x 64 ci)
//   B = dY                     : 64 voxels x NB (co)    -> N = NB <= 256 accumulator columns
// A K-chunk is a box of 64 output voxels; each 64-channel slice of it is ONE TMA box (128-byte rows = 64 channels,
// SWIZZLE_128B), exactly the canonical MN-major SW128 atom, so no transposes are needed.  A CTA keeps NACC accumulator
// blocks (different taps / ci blocks that share the dY tile) in TMEM, walks its slice of the voxels (split-K) and stores
// fp32 partials [split][Cout][taps][Cin]; mmad_wgrad_reduce sums the splits into the torch layout.
#include "tc_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace mmad {

struct WgradGeom {
    int N, D, H, W, Cin;
    int Do, Ho, Wo, Cout;
    int k, taps, stride, pad, dil;
    int tw, th, td;               // output-voxel chunk box, tw*th*td == 64
    int tiles_w, tiles_h, tiles_d, n_chunks;
    int mode2;                    // 1: Cin == 64, an M block is two taps x 64 ci; 0: one tap x 128 ci
    int units;                    // M blocks in total
    int cib;                      // ci blocks per tap (mode 1): Cin / 128
    int nb, n_tiles;              // co tile width, number of co tiles
    int nacc, ugroups;            // accumulator blocks per CTA, unit groups
    int nsplit, stages;
    int can_skip;                 // some (tap, chunk) input box lies entirely in the padding (dilated layers): worth testing
};

// Per accumulator block: input-box origin offsets of its (up to two) taps, computed once per CTA.
struct WgTapOff { int w[2], h[2], d[2]; };
__device__ __forceinline__ void wg_tap_offsets(const WgradGeom& g, int u, WgTapOff& o) {
    for (int h = 0; h < 2; ++h) {
        const int tap = g.mode2 ? min(2 * u + h, g.taps - 1) : u / g.cib;
        o.w[h] = (tap % g.k) * g.dil - g.pad;
        o.h[h] = ((tap / g.k) % g.k) * g.dil - g.pad;
        o.d[h] = (tap / (g.k * g.k)) * g.dil - g.pad;
    }
}
// true when the input boxes of both taps of the block lie entirely in the zero padding for the chunk at (ow0,oh0,od0)
__device__ __forceinline__ bool wg_oob(const WgradGeom& g, const WgTapOff& o, int ow0, int oh0, int od0) {
    bool all = true;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int lw = ow0 * g.stride + o.w[h], lh = oh0 * g.stride + o.h[h], ld = od0 * g.stride + o.d[h];
        all &= lw + (g.tw - 1) * g.stride < 0 || lw >= g.W || lh + (g.th - 1) * g.stride < 0 || lh >= g.H ||
               ld + (g.td - 1) * g.stride < 0 || ld >= g.D;
    }
    return all;
}

constexpr int kWgThreads = 256;
constexpr int kWgProducers = 3;       // warps 0, 2, 3: one elected thread issues ~1 TMA op per 300 cycles, so chunks are dealt round-robin
constexpr int kBoxBytes = 64 * 128;   // 64 voxels x 64 channels bf16

__global__ void __launch_bounds__(kWgThreads, 1)
conv3d_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradGeom g,
                    float* __restrict__ partials) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const int S = g.stages;
    const int nbox_b = g.nb / 64;
    const uint32_t STAGE = (uint32_t)(nbox_b + 2 * g.nacc) * kBoxBytes;     // B boxes, then 2 boxes per accumulator block
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)S * STAGE);   // full[S], empty[S], tfull
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull = empty0 + 8 * S;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item: unit group fastest, K slice slowest, so the CTAs resident at the same time walk the SAME voxel slice
    // (different taps / channel blocks of it) and the slice stays in L2
    int item = blockIdx.x;
    const int ug = item % g.ugroups; item /= g.ugroups;
    const int nt = item % g.n_tiles; item /= g.n_tiles;
    const int ks = item;
    const int u0 = ug * g.nacc;
    const int nu = min(g.nacc, g.units - u0);                               // accumulator blocks this CTA really owns
    const int c_begin = (int)((long long)g.n_chunks * ks / g.nsplit), c_end = (int)((long long)g.n_chunks * (ks + 1) / g.nsplit);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers: executed chunk i is issued by producer i % 3 ============================
        if (lane == 0) {
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t s = 0, ph = 0, turn = 0;
            const uint32_t nprod = (uint32_t)min(kWgProducers, S);     // parity waits: a producer must not lap a slot twice, so stages >= active producers
            WgTapOff toff[4];
            for (int a = 0; a < nu; ++a) wg_tap_offsets(g, u0 + a, toff[a]);
            for (int c = c_begin; c < c_end; ++c) {
                int r = c;
                const int wt = r % g.tiles_w; r /= g.tiles_w;
                const int ht = r % g.tiles_h; r /= g.tiles_h;
                const int dt = r % g.tiles_d; r /= g.tiles_d;
                const int n = r;
                const int ow0 = wt * g.tw, oh0 = ht * g.th, od0 = dt * g.td;
                // a block is skipped for a chunk whose input boxes are all padding - except on the first chunk of the K
                // slice, which always runs so that every accumulator gets initialised
                uint32_t skip = 0;
                if (g.can_skip && c != c_begin)
                    for (int a = 0; a < nu; ++a) skip |= (wg_oob(g, toff[a], ow0, oh0, od0) ? 1u : 0u) << a;
                if (__popc(skip) == nu) continue;                                  // nothing to do for this chunk
                if (turn == me) {
                    const uint32_t bytes = (uint32_t)(nbox_b + 2 * (nu - __popc(skip))) * kBoxBytes;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                    const uint32_t sb = base + s * STAGE;
                    for (int j = 0; j < nbox_b; ++j) tma_load_5d(sb + j * kBoxBytes, &tmDY, full0 + 8 * s, nt * g.nb + 64 * j, ow0, oh0, od0, n);
                    for (int a = 0; a < nu; ++a) {
                        if ((skip >> a) & 1u) continue;
                        const int u = u0 + a;
                        for (int h = 0; h < 2; ++h) {
                            const int ci0 = g.mode2 ? 0 : (u % g.cib) * 128 + 64 * h;
                            tma_load_5d(sb + (uint32_t)(nbox_b + 2 * a + h) * kBoxBytes, &tmX, full0 + 8 * s, ci0,
                                        ow0 * g.stride + toff[a].w[h], oh0 * g.stride + toff[a].h[h], od0 * g.stride + toff[a].d[h], n);
                        }
                    }
                }
                if (++turn == nprod) turn = 0;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, g.nb, 1, 1);
            uint32_t s = 0, ph = 0;
            WgTapOff toff[4];
            for (int a = 0; a < nu; ++a) wg_tap_offsets(g, u0 + a, toff[a]);
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t skip = 0;
                if (g.can_skip && c != c_begin) {
                    int r = c;
                    const int wt = r % g.tiles_w; r /= g.tiles_w;
                    const int ht = r % g.tiles_h; r /= g.tiles_h;
                    const int dt = r % g.tiles_d;
                    const int ow0 = wt * g.tw, oh0 = ht * g.th, od0 = dt * g.td;
                    for (int a = 0; a < nu; ++a) skip |= (wg_oob(g, toff[a], ow0, oh0, od0) ? 1u : 0u) << a;
                }
                if (__popc(skip) == nu) continue;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t sb = base + s * STAGE;
                // MN-major SW128: 128-byte rows are K (voxel) indices, 8-row groups 1024 B apart (SBO), 64-wide M/N atoms one box apart (LBO)
                const uint64_t bdesc = umma_desc_sw128(sb, kBoxBytes, 1024);
                for (int a = 0; a < nu; ++a) {
                    if ((skip >> a) & 1u) continue;
                    const uint64_t adesc = umma_desc_sw128(sb + (uint32_t)(nbox_b + 2 * a) * kBoxBytes, kBoxBytes, 1024);
#pragma unroll
                    for (int j = 0; j < 4; ++j)                   // K16 = 16 voxel rows = 2048 bytes
                        umma_bf16(tmem_base + a * g.nb, adesc + 128 * j, bdesc + 128 * j, idesc, (c > c_begin || j) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
            umma_commit(tfull);
        }
    } else if (warp >= 4) {
        // ============================ epilogue: TMEM -> fp32 partials ============================
        const int ew = warp - 4;
        const int m = ew * 32 + lane;                              // accumulator row
        if (c_end > c_begin) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        const size_t plane = (size_t)g.Cout * g.taps * g.Cin;
        float* out = partials + (size_t)ks * plane;
        for (int a = 0; a < nu; ++a) {
            const int u = u0 + a;
            int tap, ci;
            if (g.mode2) { tap = 2 * u + (m >> 6); ci = m & 63; }
            else { tap = u / g.cib; ci = (u % g.cib) * 128 + m; }
            const bool valid = tap < g.taps;
            for (int n0 = 0; n0 < g.nb; n0 += 32) {
                uint32_t v[32];
                if (c_end > c_begin) {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + a * g.nb + n0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int co = nt * g.nb + n0 + j;
                        out[((size_t)co * g.taps + tap) * g.Cin + ci] = __uint_as_float(v[j]);   // lanes = consecutive ci: coalesced
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

static void pick_chunk(int Wo, int Ho, int Do, int stride, int& tw, int& th, int& td) {
    double best = 1e30;
    for (int a = 1; a <= 64; a *= 2)
        for (int b = 1; a * b <= 64; b *= 2) {
            const int c = 64 / (a * b);
            if (a * stride > 256 || b * stride > 256 || c * stride > 256) continue;
            const double cover = (double)((Wo + a - 1) / a * a) * ((Ho + b - 1) / b * b) * ((Do + c - 1) / c * c);
            const double score = cover - 1e-3 * a;
            if (score < best) { best = score; tw = a; th = b; td = c; }
        }
}

static int fill_geom(WgradGeom& g, int N, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dil, int sms) {
    g = WgradGeom{};
    g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
    g.k = k; g.taps = k * k * k; g.stride = stride; g.pad = pad; g.dil = dil;
    g.Do = (D + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    g.Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    g.Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    if (g.Do <= 0 || g.Ho <= 0 || g.Wo <= 0) return -1;
    pick_chunk(g.Wo, g.Ho, g.Do, stride, g.tw, g.th, g.td);
    g.tiles_w = (g.Wo + g.tw - 1) / g.tw; g.tiles_h = (g.Ho + g.th - 1) / g.th; g.tiles_d = (g.Do + g.td - 1) / g.td;
    const long long chunks = (long long)N * g.tiles_w * g.tiles_h * g.tiles_d;
    if (chunks > 0x7fffffffLL) return -1;
    g.n_chunks = (int)chunks;
    g.mode2 = Cin == 64;
    g.cib = g.mode2 ? 1 : Cin / 128;
    g.units = g.mode2 ? (g.taps + 1) / 2 : g.taps * g.cib;
    g.nb = std::min(256, Cout);
    g.n_tiles = Cout / g.nb;
    g.nacc = g.nb == 256 ? 2 : (g.nb == 128 ? 3 : 4);
    if (const char* e = getenv("MMAD_WG_NACC")) g.nacc = std::max(1, std::min(g.nacc, atoi(e)));   // tuning knob
    g.nacc = std::min(g.nacc, g.units);
    g.ugroups = (g.units + g.nacc - 1) / g.nacc;
    const int items = g.n_tiles * g.ugroups;
    // split count: at least ~2 waves of CTAs, and a grid that fills whole waves (the last wave is the tail)
    int best_s = 1;
    double best_score = -1.0;
    const int sp_max = std::min(g.n_chunks, std::max(96, (4 * sms + items - 1) / items));
    for (int sp = 1; sp <= sp_max; ++sp) {
        const long long ctas = (long long)items * sp;
        if (ctas < 2LL * sms && sp < sp_max) continue;
        const long long waves = (ctas + sms - 1) / sms;
        const double score = (double)ctas / (double)(waves * sms) - 0.004 * sp;
        if (score > best_score) { best_score = score; best_s = sp; }
    }
    g.nsplit = best_s;
    const int stage = (g.nb / 64 + 2 * g.nacc) * kBoxBytes;
    g.stages = std::max(2, std::min(6, (227 * 1024 - 1024 - 256) / stage));
    if (const char* e = getenv("MMAD_WG_STAGES")) g.stages = std::max(1, std::min(g.stages, atoi(e)));   // tuning knob
    // is any (tap, chunk origin) input box entirely padding along some axis?  (only then is the per-chunk test worth running)
    g.can_skip = 0;
    const int ext[3] = {W, H, D}, tl[3] = {g.tw, g.th, g.td}, nt[3] = {g.tiles_w, g.tiles_h, g.tiles_d};
    for (int ax = 0; ax < 3 && !g.can_skip; ++ax)
        for (int t = 0; t < k && !g.can_skip; ++t)
            for (int i = 0; i < nt[ax]; ++i) {
                const int lo = i * tl[ax] * stride + t * dil - pad;
                if (lo + (tl[ax] - 1) * stride < 0 || lo >= ext[ax]) { g.can_skip = 1; break; }
            }
    return 0;
}

}  // namespace mmad

using namespace mmad;

extern "C" {

// fp32 elements of the split-K workspace mmad_conv3d_wgrad_bf16 needs, and the split count it will use.
int64_t mmad_conv3d_wgrad_workspace(int N, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dil, int* nsplit_out) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    WgradGeom g;
    if (fill_geom(g, N, D, H, W, Cin, Cout, k, stride, pad, dil, sms)) return -1;
    if (nsplit_out) *nsplit_out = g.nsplit;
    return (int64_t)g.nsplit * Cout * g.taps * Cin;
}

int mmad_conv3d_wgrad_bf16(const void* x, const void* dy, float* partials, int N, int D, int H, int W, int Cin, int Cout, int k,
                           int stride, int pad, int dil, void* stream) {
    MMAD_CHECK_ARG(x && dy && partials, "conv3d_wgrad: null pointer");
    MMAD_CHECK_ARG(Cin % 64 == 0 && (Cin == 64 || Cin % 128 == 0), "conv3d_wgrad: Cin must be 64 or a multiple of 128");
    MMAD_CHECK_ARG(Cout % 64 == 0 && (Cout == 64 || Cout == 128 || Cout % 256 == 0), "conv3d_wgrad: Cout must be 64, 128 or a multiple of 256");
    MMAD_CHECK_ARG(k >= 1 && k <= 7 && stride >= 1 && stride <= 2 && dil >= 1 && pad >= 0, "conv3d_wgrad: bad kernel geometry");
    int dev = 0, sms = 148;
    MMAD_CUDA(cudaGetDevice(&dev));
    MMAD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WgradGeom g;
    MMAD_CHECK_ARG(fill_geom(g, N, D, H, W, Cin, Cout, k, stride, pad, dil, sms) == 0, "conv3d_wgrad: empty output");
    CUtensorMap tmX, tmDY;
    {
        const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
        const uint64_t str[4] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2, (uint64_t)D * H * W * Cin * 2};
        const uint32_t box[5] = {64, (uint32_t)(g.tw * stride), (uint32_t)(g.th * stride), (uint32_t)(g.td * stride), 1};
        const uint32_t es[5] = {1, (uint32_t)stride, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_tmap_bf16(&tmX, x, 5, dims, str, box, es);
        if (rc) return rc;
    }
    {
        const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
        const uint64_t str[4] = {(uint64_t)Cout * 2, (uint64_t)g.Wo * Cout * 2, (uint64_t)g.Ho * g.Wo * Cout * 2,
                                 (uint64_t)g.Do * g.Ho * g.Wo * Cout * 2};
        const uint32_t box[5] = {64, (uint32_t)g.tw, (uint32_t)g.th, (uint32_t)g.td, 1};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        int rc = make_tmap_bf16(&tmDY, dy, 5, dims, str, box, es);
        if (rc) return rc;
    }
    const int stage = (g.nb / 64 + 2 * g.nacc) * kBoxBytes;
    const int smem = 1024 + g.stages * stage + (2 * g.stages + 1) * 8 + 32;
    static bool attr_done = false;
    if (!attr_done) {
        MMAD_CUDA(cudaFuncSetAttribute(conv3d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    const int grid = g.n_tiles * g.ugroups * g.nsplit;
    conv3d_wgrad_kernel<<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(tmX, tmDY, g, partials);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    return MMAD_OK;
}

}  // extern "C"
