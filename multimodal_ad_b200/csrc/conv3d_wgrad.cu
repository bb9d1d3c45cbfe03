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
#include <array>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>

namespace mmad {

__device__ unsigned long long g_mma_flops_wgrad;           // executed tensor-core flops of this file's kernels (tc_common.cuh)
long long mma_flops_wgrad() {
    unsigned long long v = 0;
    return cudaMemcpyFromSymbol(&v, g_mma_flops_wgrad, sizeof(v)) == cudaSuccess ? (long long)v : -1;
}

struct WgradGeom {
    int N, D, H, W, Cin;
    int Do, Ho, Wo, Cout;
    int k, taps, stride, pad, dil;
    int tw, th, td, tn;           // output-voxel chunk box (tn batch samples deep), tw*th*td*tn == 64 (128: pair kernel)
    int tiles_w, tiles_h, tiles_d, n_chunks;
    int pairk;                    // 1: CTA-pair kernel (cta_group::2)
    int cv;                       // voxels per K chunk (64 or 128)
    int halo;                     // 1: halo kernel (64 -> 64 channels, 3x3x3, unit stride, undilated)
    unsigned char ug_order[64];   // pair kernel: unit groups in decreasing order of work (padding skips), heaviest launched first
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
constexpr int kWgProducers = 3;       // warps 0, 2, 3: chunks are dealt round-robin to three issuing warps
constexpr int kBoxBytes = 64 * 128;   // 64 voxels x 64 channels bf16

__global__ void __launch_bounds__(kWgThreads, 1)
conv3d_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradGeom g,
                    float* __restrict__ partials) {
    pdl_launch_dependents();
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
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers: executed chunk i is issued by producer i % 3 ============================
        {
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
                const int n = r * g.tn;
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
                    if (elect_one()) {
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
                    __syncwarp();
                }
                if (++turn == nprod) turn = 0;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        {
            const uint32_t idesc = umma_idesc_bf16(128, g.nb, 1, 1);
            uint32_t s = 0, ph = 0, nmma = 0;
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
                if (elect_one()) {
                    for (int a = 0; a < nu; ++a) {
                        if ((skip >> a) & 1u) continue;
                        const uint64_t adesc = umma_desc_sw128(sb + (uint32_t)(nbox_b + 2 * a) * kBoxBytes, kBoxBytes, 1024);
#pragma unroll
                        for (int j = 0; j < 4; ++j)               // K16 = 16 voxel rows = 2048 bytes
                            umma_bf16(tmem_base + a * g.nb, adesc + 128 * j, bdesc + 128 * j, idesc, (c > c_begin || j) ? 1u : 0u);
                        nmma += 4;
                    }
                    umma_commit(empty0 + 8 * s);
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
            if (elect_one()) { umma_commit(tfull); mma_count_flush(&g_mma_flops_wgrad, nmma, 2u * 128u * (uint32_t)g.nb * 16u); }
            __syncwarp();
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


// ===============================================================================================================
// CTA-pair variant (tcgen05 cta_group::2) for Cin >= 128 and Cout >= 128.  The single-CTA kernel above is bound by the
// bytes it pulls through the L2->SM fabric per MMA cycle (64 KB of operands per 1024 MMA cycles = 62 B/cycle/SM against a
// measured cap of ~43) and by its M = 128 accumulator blocks.  Here a K chunk is 128 voxels (16 KB boxes), an accumulator block is 256 rows x NB columns spread over the
// two CTAs (each CTA stages the X boxes of ITS unit and only HALF of the dY columns), so a CTA pulls 96 KB per 2048 MMA
// cycles (47 B/cycle).  Block a of a pair holds units u0 + 2a (leader) and u0 + 2a + 1 (peer).  Barrier protocol as in
// conv3d_igemm_pair_kernel: the leader's full barrier collects both CTAs' bytes, tcgen05.commit is multicast to both.
// ===============================================================================================================

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgThreads, 1)
conv3d_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradGeom g,
                         float* __restrict__ partials) {
    pdl_launch_dependents();
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const int S = g.stages;
    const int nbox_b = g.nb / 128;                                          // dY boxes of this CTA's column half
    const uint32_t boxb = (uint32_t)g.cv * 128u;                            // one box: cv voxels x 64 channels bf16
    const uint32_t STAGE = (uint32_t)(nbox_b + 2 * g.nacc) * boxb;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)S * STAGE);   // full[S], empty[S], tfull
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull = empty0 + 8 * S;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    int item = blockIdx.x >> 1;
    const int ug = g.ugroups <= 64 ? (int)g.ug_order[item % g.ugroups] : item % g.ugroups; item /= g.ugroups;
    const int nt = item % g.n_tiles; item /= g.n_tiles;
    const int ks = item;
    const int u0 = ug * 2 * g.nacc;
    const int nblk = min(g.nacc, (g.units - u0 + 1) / 2);                   // accumulator blocks this pair really owns
    const int c_begin = (int)((long long)g.n_chunks * ks / g.nsplit), c_end = (int)((long long)g.n_chunks * (ks + 1) / g.nsplit);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY);
    }
    if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_ptr_s), 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    // Per accumulator block (registers, static indexing): the input-box origin offset of THIS CTA's unit, and for both
    // units of the block one bit mask per axis of the chunk tiles whose input range is entirely padding (bit i = tile i).
    // A block is skipped for a chunk when the boxes of BOTH its units are padding - the same decision in both CTAs and in
    // the MMA issuer; the first chunk of the K slice always runs so that every accumulator gets initialised.
    int offw[2], offh[2], offd[2], cib0[2];
    uint32_t oobw[2][2], oobh[2][2], oobd[2][2];
    auto axis_oob = [&](int off, int t, int tiles, int ext) -> uint32_t {
        uint32_t m = 0;
        if (g.can_skip)
            for (int i = 0; i < tiles; ++i) {
                const int lo = i * t * g.stride + off;
                if (lo + (t - 1) * g.stride < 0 || lo >= ext) m |= 1u << i;
            }
        return m;
    };
    if (warp < 4) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int u = min(u0 + 2 * a + h, g.units - 1);                 // phantom unit past the end: the last one again
                const int tap = u / g.cib;
                const int ow = (tap % g.k) * g.dil - g.pad, oh = ((tap / g.k) % g.k) * g.dil - g.pad, od = (tap / (g.k * g.k)) * g.dil - g.pad;
                oobw[a][h] = axis_oob(ow, g.tw, g.tiles_w, g.W);
                oobh[a][h] = axis_oob(oh, g.th, g.tiles_h, g.H);
                oobd[a][h] = axis_oob(od, g.td, g.tiles_d, g.D);
                if (h == (int)rank) { offw[a] = ow; offh[a] = oh; offd[a] = od; cib0[a] = (u % g.cib) * 128; }
            }
    }
    // chunk iterator: tile coordinates advanced incrementally (the loops below are single-thread critical paths)
    int wt, ht, dt, nn;
    {
        int r = c_begin;
        wt = r % g.tiles_w; r /= g.tiles_w;
        ht = r % g.tiles_h; r /= g.tiles_h;
        dt = r % g.tiles_d; r /= g.tiles_d;
        nn = r;
    }
    auto next_chunk = [&]() {
        if (++wt == g.tiles_w) { wt = 0; if (++ht == g.tiles_h) { ht = 0; if (++dt == g.tiles_d) { dt = 0; ++nn; } } }
    };
    auto chunk_skip = [&](int c) -> uint32_t {
        if (c == c_begin) return 0u;
        uint32_t skip = 0;
#pragma unroll
        for (int a = 0; a < 2; ++a)
            if (a < nblk)
                skip |= ((((oobw[a][0] >> wt) | (oobh[a][0] >> ht) | (oobd[a][0] >> dt)) & ((oobw[a][1] >> wt) | (oobh[a][1] >> ht) | (oobd[a][1] >> dt))) & 1u) << a;
        return skip;
    };

    if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers (in both CTAs) ============================
        {
            // every producer warp takes part in EVERY stage and issues every third box of it, so that the issue time of a
            // (large, two-deep) stage stays off the critical path
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t s = 0, ph = 0;
            for (int c = c_begin; c < c_end; ++c, next_chunk()) {
                const uint32_t skip = chunk_skip(c);
                if (__popc(skip) == nblk) continue;
                const int ow0 = wt * g.tw, oh0 = ht * g.th, od0 = dt * g.td, n0 = nn * g.tn;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                if (elect_one()) {
                    if (leader && me == 0)
                        mbar_arrive_expect_tx(full0 + 8 * s, 2u * (uint32_t)(nbox_b + 2 * (nblk - __popc(skip))) * boxb);
                    const uint32_t sb = base + s * STAGE;
                    uint32_t q = 0;
                    for (int j = 0; j < nbox_b; ++j, ++q)
                        if (q % kWgProducers == me)
                            tma_load_5d_2sm(sb + j * boxb, &tmDY, full0 + 8 * s, nt * g.nb + (int)rank * (g.nb / 2) + 64 * j, ow0, oh0, od0, n0);
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        if (a >= nblk || ((skip >> a) & 1u)) continue;
#pragma unroll
                        for (int h = 0; h < 2; ++h, ++q)
                            if (q % kWgProducers == me)
                                tma_load_5d_2sm(sb + (uint32_t)(nbox_b + 2 * a + h) * boxb, &tmX, full0 + 8 * s, cib0[a] + 64 * h,
                                                ow0 * g.stride + offw[a], oh0 * g.stride + offh[a], od0 * g.stride + offd[a], n0);
                    }
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        }
    }
    if (warp == 1) {
        // ============================ MMA issuer: leader CTA only ============================
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(256, g.nb, 1, 1);
            uint32_t s = 0, ph = 0, nmma = 0;
            for (int c = c_begin; c < c_end; ++c, next_chunk()) {
                const uint32_t skip = chunk_skip(c);
                if (__popc(skip) == nblk) continue;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t sb = base + s * STAGE;
                const uint64_t bdesc = umma_desc_sw128(sb, boxb, 1024);
                if (elect_one()) {
                    for (int a = 0; a < nblk; ++a) {
                        if ((skip >> a) & 1u) continue;
                        const uint64_t adesc = umma_desc_sw128(sb + (uint32_t)(nbox_b + 2 * a) * boxb, boxb, 1024);
                        if (g.cv == 128) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)           // K16 = 16 voxel rows = 2048 bytes
                                umma_bf16_2sm(tmem_base + a * g.nb, adesc + 128 * j, bdesc + 128 * j, idesc, (c > c_begin || j) ? 1u : 0u);
                            nmma += 8;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                umma_bf16_2sm(tmem_base + a * g.nb, adesc + 128 * j, bdesc + 128 * j, idesc, (c > c_begin || j) ? 1u : 0u);
                            nmma += 4;
                        }
                    }
                    umma_commit_2sm(empty0 + 8 * s, 3);
                }
                __syncwarp();
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
            if (elect_one()) { umma_commit_2sm(tfull, 3); mma_count_flush(&g_mma_flops_wgrad, nmma, 2u * 256u * (uint32_t)g.nb * 16u); }
            __syncwarp();
        }
    }
    if (warp >= 4) {
        // ============================ epilogue (both CTAs): own 128 accumulator rows -> fp32 partials ============================
        const int ew = warp - 4;
        const int m = ew * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const size_t plane = (size_t)g.Cout * g.taps * g.Cin;
        float* out = partials + (size_t)ks * plane;
        for (int a = 0; a < nblk; ++a) {
            const int u = u0 + 2 * a + (int)rank;
            const bool valid = u < g.units;
            const int tap = u / g.cib, ci = (u % g.cib) * 128 + m;
            for (int n0 = 0; n0 < g.nb; n0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + a * g.nb + n0, v);
                tmem_ld_wait();
                if (c_end == c_begin) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int co = nt * g.nb + n0 + j;
                        out[((size_t)co * g.taps + tap) * g.Cin + ci] = __uint_as_float(v[j]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}


// ===============================================================================================================
// Halo variant for 3x3x3, unit-stride, undilated convolutions of 64 -> 64 channels (layer1).  The kernels above load one
// shifted copy of the input chunk per tap (9 boxes of 8 KB per 64 voxels) and are bound by that L2 traffic
// at ~0.3 PFLOP/s.  UMMA swizzles on absolute shared-memory address bits, so an operand may start at ANY 128-byte row: one
// input box with a one-voxel halo (10 x 6 x 6 voxels, 45 KB) holds the shifted chunk of every tap.  Per 128-voxel chunk
// (8 x 4 x 4) a CTA loads that box and one dY box and issues, per accumulator block (two taps x 64 ci) and K16 step (two W
// lines of 8 voxels, 10 rows apart in the halo box), one MMA whose A start is the tap's row offset and whose second
// 64-row atom (LBO) is the row distance to the next tap.  14 taps (7 blocks x 64 columns of TMEM) per CTA: grid =
// 2 tap groups x nsplit voxel slices; group 1 recomputes tap 13 (not stored) so that both groups hold whole pairs.
// ===============================================================================================================
constexpr int kHaloRows = 10 * 6 * 6;                 // halo box rows (voxels)
constexpr int kHaloBoxBytes = kHaloRows * 128;        // 46080 = 45 KB
constexpr int kHaloDyBytes = 128 * 128;               // dY box: 128 voxels x 64 channels

__device__ __forceinline__ int halo_row(int tap) {    // row of tap (a, b, c)'s first voxel in the halo box
    return ((tap / 9) * 6 + (tap / 3) % 3) * 10 + tap % 3;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv3d_wgrad_halo64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradGeom g,
                           float* __restrict__ partials) {
    pdl_launch_dependents();
    constexpr int S = 3;
    constexpr uint32_t STAGE = kHaloBoxBytes + kHaloDyBytes;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 1, 1);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)S * STAGE);   // full[S], empty[S], tfull
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull = empty0 + 8 * S;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.x & 1, ks = blockIdx.x >> 1;                   // tap group, voxel slice
    const int tap0 = grp * 13;                                              // group 0: taps 0..13, group 1: taps 13..26
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
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers: chunk i is issued by producer i % 3 (S == 3) ============================
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
                    tma_load_5d(sb, &tmX, full0 + 8 * s, 0, wt * 8 - 1, ht * 4 - 1, dt * 4 - 1, r);
                    tma_load_5d(sb + kHaloBoxBytes, &tmDY, full0 + 8 * s, 0, wt * 8, ht * 4, dt * 4, r);
                }
                __syncwarp();
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        uint32_t s = 0, ph = 0, nmma = 0;
        // descriptor low words = stage base + a per-block constant (tap row offset, LBO = row distance to the next tap) + a
        // compile-time K16 offset; the high words are constants - one 32-bit add per operand in the issue loop (see
        // umma_bf16_lohi, tc_common.cuh: the issuing thread is the narrowest pipe of an N = 64 kernel)
        uint32_t aoff[7];
#pragma unroll
        for (int blk = 0; blk < 7; ++blk) {
            const int t = tap0 + 2 * blk;
            aoff[blk] = (uint32_t)halo_row(t) * 8u + ((((uint32_t)(halo_row(t + 1) - halo_row(t)) * 128u) >> 4) << 16);
        }
        constexpr uint32_t AH = umma_desc_hi_sw128(1280), BH = umma_desc_hi_sw128(1024);
        constexpr uint32_t BOFF = (kHaloBoxBytes >> 4) + ((kHaloDyBytes >> 4) << 16);
        for (int c = c_begin; c < c_end; ++c) {
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const uint32_t sb = base + s * STAGE;
            if (elect_one()) {
                const uint32_t s_lo = (sb & 0x3ffffu) >> 4;
#pragma unroll
                for (int blk = 0; blk < 7; ++blk) {
                    const uint32_t a_blk = s_lo + aoff[blk];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                 // K16 = W lines (h, h+1) of slice d: 2j -> h = 2 * (j & 1), d = j >> 1
                        const uint32_t line8 = (uint32_t)((j >> 1) * 6 + (j & 1) * 2) * 10u * 8u;
                        umma_bf16_lohi(tmem_base + blk * 64, a_blk + line8, AH, s_lo + (BOFF + j * 128), BH, IDESC, (c > c_begin || j) ? 1u : 0u);
                    }
                }
                nmma += 56;
                umma_commit(empty0 + 8 * s);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1; }
        }
        if (elect_one()) { umma_commit(tfull); mma_count_flush(&g_mma_flops_wgrad, nmma, 2u * 128u * 64u * 16u); }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================ epilogue: TMEM -> fp32 partials [slice][co][tap][ci] ============================
        const int ew = warp - 4;
        const int m = ew * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        float* out = partials + (size_t)ks * 64 * 27 * 64;
        for (int blk = 0; blk < 7; ++blk) {
            const int tap = tap0 + 2 * blk + (m >> 6), ci = m & 63;
            const bool store = !(grp == 1 && tap == 13);          // tap 13 belongs to group 0
            for (int n0 = 0; n0 < 64; n0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + blk * 64 + n0, v);
                tmem_ld_wait();
                if (c_end == c_begin) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (store) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) out[((size_t)(n0 + j) * 27 + tap) * 64 + ci] = __uint_as_float(v[j]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// (chunk, tap) pairs along one axis whose input range is not entirely padding (closed form: the stem view has millions of tiles)
long long axis_work(int ext_in, int ext_out, int t, int k, int stride, int pad, int dil) {
    auto ceil_div = [](long long a, long long b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); };
    const long long tiles = (ext_out + t - 1) / t, step = (long long)t * stride;
    long long w = 0;
    for (int tap = 0; tap < k; ++tap) {
        const long long off = (long long)tap * dil - pad;
        // tile i covers input [i*step + off, i*step + off + (t-1)*stride]
        const long long below = std::min(tiles, std::max(0LL, ceil_div(-off - (long long)(t - 1) * stride, step)));   // entirely < 0
        const long long first_above = std::min(tiles, std::max(0LL, ceil_div((long long)ext_in - off, step)));           // first tile >= ext
        w += std::max(0LL, first_above - below);
    }
    return w;
}

// K-chunk box tw x th x td x tn of `cv` output voxels (powers of two; tn > 1 only when the batch divides) with the fewest
// (chunk, tap) pairs left after padding skips; ties go to the widest box in W (longer contiguous rows).
void pick_chunk(int cv, int N, int W, int H, int D, int Wo, int Ho, int Do, int k, int stride, int pad, int dil, int& tw,
                       int& th, int& td, int& tn) {   // (callers may override the result: MMAD_WG_SHAPE)
    double best = 1e30;
    for (int n = 1; n <= 2; n *= 2) {
        if (N % n) continue;
        for (int a = 1; a <= cv / n; a *= 2)
            for (int b = 1; a * b <= cv / n; b *= 2) {
                const int c = cv / n / (a * b);
                if (a * stride > 256 || b * stride > 256 || c * stride > 256) continue;
                const double work = (double)(N / n) * axis_work(W, Wo, a, k, stride, pad, dil) * axis_work(H, Ho, b, k, stride, pad, dil) *
                                    axis_work(D, Do, c, k, stride, pad, dil);
                const double score = work * (1.0 + 1e-3 * n) - 1e-3 * a;
                if (score < best) { best = score; tw = a; th = b; td = c; tn = n; }
            }
    }
}

static int fill_geom_uncached(WgradGeom& g, int N, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dil, int sms) {
    g = WgradGeom{};
    g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
    g.k = k; g.taps = k * k * k; g.stride = stride; g.pad = pad; g.dil = dil;
    g.Do = (D + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    g.Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    g.Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    if (g.Do <= 0 || g.Ho <= 0 || g.Wo <= 0) return -1;
    static int halo_mode = -1;                         // halo kernel: on unless MMAD_WG_HALO=0
    if (halo_mode < 0) { const char* e = getenv("MMAD_WG_HALO"); halo_mode = e ? atoi(e) : 1; }
    if (halo_mode != 0 && Cin == 64 && Cout == 64 && k == 3 && stride == 1 && dil == 1 && pad == 1) {
        g.halo = 1;
        g.tw = 8; g.th = 4; g.td = 4; g.tn = 1; g.cv = 128;
        g.tiles_w = (g.Wo + 7) / 8; g.tiles_h = (g.Ho + 3) / 4; g.tiles_d = (g.Do + 3) / 4;
        const long long chunks = (long long)N * g.tiles_w * g.tiles_h * g.tiles_d;
        if (chunks > 0x7fffffffLL) return -1;
        g.n_chunks = (int)chunks;
        g.nb = 64; g.n_tiles = 1; g.units = 14; g.nacc = 7; g.ugroups = 2; g.stages = 3;
        g.nsplit = (int)std::max<long long>(1, std::min<long long>(chunks, sms / 2));   // 2 tap groups x nsplit slices = one wave
        if (const char* e = getenv("MMAD_WG_NSPLIT")) g.nsplit = std::max(1, std::min(g.n_chunks, atoi(e)));
        return 0;
    }
    g.mode2 = Cin == 64;
    g.cib = g.mode2 ? 1 : Cin / 128;
    g.units = g.mode2 ? (g.taps + 1) / 2 : g.taps * g.cib;
    g.nb = std::min(256, Cout);
    g.n_tiles = Cout / g.nb;
    static int pair_mode = -1;                         // CTA-pair kernel: on unless MMAD_WG_PAIR=0
    if (pair_mode < 0) { const char* e = getenv("MMAD_WG_PAIR"); pair_mode = e ? atoi(e) : 1; }
    g.pairk = pair_mode != 0 && !g.mode2 && g.nb >= 128 && g.units >= 2;
    g.cv = 64;
    if (g.pairk) { static int pcv = -1; if (pcv < 0) { const char* e = getenv("MMAD_WG_PAIR_CV"); pcv = e ? atoi(e) : 128; } g.cv = pcv == 128 ? 128 : 64; }
    pick_chunk(g.cv, N, W, H, D, g.Wo, g.Ho, g.Do, k, stride, pad, dil, g.tw, g.th, g.td, g.tn);
    if (const char* e = getenv("MMAD_WG_SHAPE")) {     // tuning knob: "tw,th,td,tn" (product must equal the chunk size)
        int a, b, c, n;
        if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &n) == 4 && a * b * c * n == g.cv && N % n == 0) { g.tw = a; g.th = b; g.td = c; g.tn = n; }
    }
    g.tiles_w = (g.Wo + g.tw - 1) / g.tw; g.tiles_h = (g.Ho + g.th - 1) / g.th; g.tiles_d = (g.Do + g.td - 1) / g.td;
    const long long chunks = (long long)(N / g.tn) * g.tiles_w * g.tiles_h * g.tiles_d;
    if (chunks > 0x7fffffffLL) return -1;
    g.n_chunks = (int)chunks;
    if (g.pairk) {
        g.nacc = 2;                                    // blocks of 256 rows (one unit per CTA): 2 x nb <= 512 TMEM columns
        g.nacc = std::min(g.nacc, (g.units + 1) / 2);
        g.ugroups = (g.units + 2 * g.nacc - 1) / (2 * g.nacc);
    } else {
        g.nacc = g.nb == 256 ? 2 : (g.nb == 128 ? 3 : 4);
        if (const char* e = getenv("MMAD_WG_NACC")) g.nacc = std::max(1, std::min(g.nacc, atoi(e)));   // tuning knob
        g.nacc = std::min(g.nacc, g.units);
        g.ugroups = (g.units + g.nacc - 1) / g.nacc;
    }
    const int items = g.n_tiles * g.ugroups;
    if (g.pairk) sms /= 2;                             // the schedulable unit is a CTA pair
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
    if (g.pairk) {
        // pair kernel: an item costs ~25 us of fill + epilogue on top of ~1.5 us per 128-voxel chunk (measured), and every
        // split adds a plane of fp32 partials to write and reduce: minimise the modelled time.  In a single wave the slowest
        // item (a tap that never falls in the padding) sets the time; over several waves the padding skips average out.
        const double frac = (double)(axis_work(W, g.Wo, g.tw, k, stride, pad, dil) * axis_work(H, g.Ho, g.th, k, stride, pad, dil) *
                                     axis_work(D, g.Do, g.td, k, stride, pad, dil)) /
                            ((double)g.tiles_w * g.tiles_h * g.tiles_d * g.taps);
        const double t_chunk = std::max(0.8, 1.5 * g.nb / 256.0) * g.cv / 128.0;
        const double plane_mb = (double)Cout * g.taps * Cin * 4.0 / 1e6;
        double best_t = 1e30;
        for (int sp = 1; sp <= std::min(g.n_chunks, 64); ++sp) {
            const long long pairs = (long long)items * sp;
            const long long waves = (pairs + sms - 1) / sms;
            const double f = waves == 1 ? 1.0 : 0.5 + 0.5 * frac;
            const double t = waves * (((g.n_chunks + sp - 1) / sp) * t_chunk * f + 25.0) + sp * plane_mb * 0.25;
            if (t < best_t) { best_t = t; best_s = sp; }
        }
    }
    if (g.pairk && g.ugroups <= 64) {
        // Unit groups whose taps fall into the padding skip chunks and finish early.  Blocks are handed to SMs in index order,
        // so within every voxel slice the groups are launched heaviest first: the light ones fill the tail of the last wave.
        double work[64];
        for (int ug = 0; ug < g.ugroups; ++ug) {
            work[ug] = 0.0;
            for (int u = ug * 2 * g.nacc; u < std::min(g.units, (ug + 1) * 2 * g.nacc); ++u) {
                const int tap = u / g.cib;
                const int off[3] = {(tap % k) * dil - pad, ((tap / k) % k) * dil - pad, (tap / (k * k)) * dil - pad};
                const int ext[3] = {W, H, D}, tl[3] = {g.tw, g.th, g.td}, nt[3] = {g.tiles_w, g.tiles_h, g.tiles_d};
                double wtap = 1.0;
                for (int ax = 0; ax < 3; ++ax) {
                    int ok = 0;
                    for (int i = 0; i < nt[ax]; ++i) {
                        const int lo = i * tl[ax] * stride + off[ax];
                        if (!(lo + (tl[ax] - 1) * stride < 0 || lo >= ext[ax])) ++ok;
                    }
                    wtap *= (double)ok / nt[ax];
                }
                work[ug] += wtap;
            }
        }
        int idx[64];
        for (int i = 0; i < g.ugroups; ++i) idx[i] = i;
        std::stable_sort(idx, idx + g.ugroups, [&](int a, int b) { return work[a] > work[b]; });
        for (int i = 0; i < g.ugroups; ++i) g.ug_order[i] = (unsigned char)idx[i];
    }
    g.nsplit = best_s;
    if (const char* e = getenv("MMAD_WG_NSPLIT")) g.nsplit = std::max(1, std::min(g.n_chunks, atoi(e)));   // tuning knob
    const int stage = g.pairk ? (g.nb / 128 + 2 * g.nacc) * g.cv * 128 : (g.nb / 64 + 2 * g.nacc) * kBoxBytes;
    g.stages = std::max(2, std::min(6, (227 * 1024 - 1024 - 256) / stage));
    if (const char* e = getenv("MMAD_WG_STAGES")) g.stages = std::max(1, std::min(g.stages, atoi(e)));   // tuning knob
    // is any (tap, chunk origin) input box entirely padding along some axis?  (only then is the per-chunk test worth running)
    g.can_skip = 0;
    const bool masks_fit = !g.pairk || (g.tiles_w <= 32 && g.tiles_h <= 32 && g.tiles_d <= 32);   // pair kernel: 32-bit tile masks
    const int ext[3] = {W, H, D}, tl[3] = {g.tw, g.th, g.td}, nt[3] = {g.tiles_w, g.tiles_h, g.tiles_d};
    for (int ax = 0; ax < 3 && !g.can_skip && masks_fit; ++ax)
        for (int t = 0; t < k && !g.can_skip; ++t)
            for (int i = 0; i < nt[ax]; ++i) {
                const int lo = i * tl[ax] * stride + t * dil - pad;
                if (lo + (tl[ax] - 1) * stride < 0 || lo >= ext[ax]) { g.can_skip = 1; break; }
            }
    if (getenv("MMAD_WG_DEBUG"))
        fprintf(stderr, "[wgrad] Cin %d Cout %d k %d s %d dil %d | pair %d cv %d chunk %dx%dx%dx%d chunks %d units %d nacc %d ugroups %d nsplit %d stages %d skip %d\n",
                Cin, Cout, k, stride, dil, g.pairk, g.cv, g.tw, g.th, g.td, g.tn, g.n_chunks, g.units, g.nacc, g.ugroups, g.nsplit, g.stages, g.can_skip);
    return 0;
}

// The geometry (chunk shape search, split model) is pure host arithmetic on the argument tuple: computed once per distinct layer.
static int fill_geom(WgradGeom& g, int N, int D, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dil, int sms) {
    static std::mutex mu;
    static std::map<std::array<int, 11>, std::pair<int, WgradGeom>> cache;
    const std::array<int, 11> key = {N, D, H, W, Cin, Cout, k, stride, pad, dil, sms};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        WgradGeom t;
        const int rc = fill_geom_uncached(t, N, D, H, W, Cin, Cout, k, stride, pad, dil, sms);
        it = cache.emplace(key, std::make_pair(rc, t)).first;
    }
    g = it->second.second;
    return it->second.first;
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

static int wgrad_impl(const void* x, long long ldx, const void* dy, float* partials, int N, int D, int H, int W, int Cin, int Cout, int k,
                      int stride, int pad, int dil, void* stream);

int mmad_conv3d_wgrad_bf16(const void* x, const void* dy, float* partials, int N, int D, int H, int W, int Cin, int Cout, int k,
                           int stride, int pad, int dil, void* stream) {
    return wgrad_impl(x, Cin, dy, partials, N, D, H, W, Cin, Cout, k, stride, pad, dil, stream);
}

// x is a channel slice [c0, c0 + Cin) of a wider NDHWC tensor whose rows are ldx elements apart (x points at channel c0): the
// weight gradient of a convolution over a concatenated input (unet3d.py:77-78) is computed per source, no copies.
int mmad_conv3d_wgrad_ex_bf16(const void* x, int64_t ldx, const void* dy, float* partials, int N, int D, int H, int W, int Cin, int Cout,
                              int k, int stride, int pad, int dil, void* stream) {
    MMAD_CHECK_ARG(ldx >= Cin && ldx % 8 == 0, "conv3d_wgrad_ex: ldx must be >= Cin and a multiple of 8");
    return wgrad_impl(x, ldx, dy, partials, N, D, H, W, Cin, Cout, k, stride, pad, dil, stream);
}

}  // extern "C"

static int wgrad_impl(const void* x, long long ldx, const void* dy, float* partials, int N, int D, int H, int W, int Cin, int Cout, int k,
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
    if (g.halo) {
        {
            const uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
            const uint64_t str[4] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2, (uint64_t)D * H * W * ldx * 2};
            const uint32_t box[5] = {64, 10, 6, 6, 1};
            const uint32_t es[5] = {1, 1, 1, 1, 1};
            int rc = make_tmap_bf16(&tmX, x, 5, dims, str, box, es);
            if (rc) return rc;
        }
        {
            const uint64_t dims[5] = {64, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
            const uint64_t str[4] = {128, (uint64_t)g.Wo * 128, (uint64_t)g.Ho * g.Wo * 128, (uint64_t)g.Do * g.Ho * g.Wo * 128};
            const uint32_t box[5] = {64, 8, 4, 4, 1};
            const uint32_t es[5] = {1, 1, 1, 1, 1};
            int rc = make_tmap_bf16(&tmDY, dy, 5, dims, str, box, es);
            if (rc) return rc;
        }
        const int smem = 1024 + 3 * (kHaloBoxBytes + kHaloDyBytes) + 7 * 8 + 32;
        static DevOnce halo_attr;
        if (halo_attr.need()) {
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_wgrad_halo64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
        launch_pdl(conv3d_wgrad_halo64_kernel, dim3(2 * g.nsplit), dim3(kWgThreads), smem, (cudaStream_t)stream, tmX, tmDY, g, partials);
        MMAD_CUDA(cudaGetLastError());
        count_launch();
        return MMAD_OK;
    }
    {
        const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
        const uint64_t str[4] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2, (uint64_t)D * H * W * ldx * 2};
        const uint32_t box[5] = {64, (uint32_t)(g.tw * stride), (uint32_t)(g.th * stride), (uint32_t)(g.td * stride), (uint32_t)g.tn};
        const uint32_t es[5] = {1, (uint32_t)stride, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_tmap_bf16(&tmX, x, 5, dims, str, box, es);
        if (rc) return rc;
    }
    {
        const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
        const uint64_t str[4] = {(uint64_t)Cout * 2, (uint64_t)g.Wo * Cout * 2, (uint64_t)g.Ho * g.Wo * Cout * 2,
                                 (uint64_t)g.Do * g.Ho * g.Wo * Cout * 2};
        const uint32_t box[5] = {64, (uint32_t)g.tw, (uint32_t)g.th, (uint32_t)g.td, (uint32_t)g.tn};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        int rc = make_tmap_bf16(&tmDY, dy, 5, dims, str, box, es);
        if (rc) return rc;
    }
    const int stage = g.pairk ? (g.nb / 128 + 2 * g.nacc) * g.cv * 128 : (g.nb / 64 + 2 * g.nacc) * kBoxBytes;
    const int smem = 1024 + g.stages * stage + (2 * g.stages + 1) * 8 + 32;
    MMAD_CHECK_ARG(smem <= 227 * 1024, "conv3d_wgrad: shared memory budget exceeded");
    static DevOnce attr_done;
    if (attr_done.need()) {
        MMAD_CUDA(cudaFuncSetAttribute(conv3d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MMAD_CUDA(cudaFuncSetAttribute(conv3d_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int grid = g.n_tiles * g.ugroups * g.nsplit;
    if (g.pairk) launch_pdl(conv3d_wgrad_pair_kernel, dim3(2 * grid), dim3(kWgThreads), smem, (cudaStream_t)stream, tmX, tmDY, g, partials);
    else launch_pdl(conv3d_wgrad_kernel, dim3(grid), dim3(kWgThreads), smem, (cudaStream_t)stream, tmX, tmDY, g, partials);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    return MMAD_OK;
}
