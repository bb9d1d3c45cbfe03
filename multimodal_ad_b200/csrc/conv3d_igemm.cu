// Conv3d as an implicit GEMM on tcgen05 / TMEM, operands staged by TMA (sm_100a).
// Replaces torch.nn.Conv3d of /root/reference/models/resnet.py:14-23 (conv3x3x3, dilated), :126-132 (stem, via
// im2col), :188-194 (1x1x1 downsample) behind include/mmad_b200.h part 2.  The same kernel computes dgrad of
// unit-stride convolutions when given the flipped / transposed weights.
//
//   activations  NDHWC bf16           weights  [Cout][taps][Cin] bf16           output  NDHWC bf16
//   GEMM view:   M = output voxels (128 per tile: a tw x th x td box), N = Cout tile (BN = 64/128/256),
//                K = taps x Cin, walked in slices of 64 channels (= one 128-byte SWIZZLE_128B row).
//   K-step:      ONE 5-D TMA box of the input (tap offset, dilation and stride folded into the box origin /
//                element strides; zero padding = TMA out-of-bounds fill) + ONE 3-D box of the weights,
//                then 4 x tcgen05.mma (M128 x BN x K16), fp32 accumulators in TMEM.
//   Warp roles:  warps 0, 2, 3 TMA producers (K-steps dealt round-robin; every role loop is warp-uniform and the TMA / MMA
//                instructions are issued under elect_one(), see common.cuh), warp 1 MMA issuer, warp 2 also
//                allocates TMEM, warps 4-7 epilogue
//                (tcgen05.ld -> bf16 -> swizzled smem -> TMA store; per-channel sum / sum-of-squares of the stored
//                bf16 values for the BatchNorm that follows).  Two TMEM accumulator buffers, persistent CTAs.
#include "tc_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace mmad {

__device__ unsigned long long g_mma_flops_igemm;           // executed tensor-core flops of this file's kernels (tc_common.cuh)
long long mma_flops_igemm() {
    unsigned long long v = 0;
    return cudaMemcpyFromSymbol(&v, g_mma_flops_igemm, sizeof(v)) == cudaSuccess ? (long long)v : -1;
}

struct ConvGeom {
    int N, D, H, W, Cin;          // input
    int Do, Ho, Wo, Cout;         // output
    int kd, kh, kw, stride, pad, dil;
    int tw, th, td, tn;           // output tile box (tn samples deep), tw*th*td*tn == 128, powers of two
    int lw, lh, ltd;              // log2(tw), log2(th), log2(td)
    int epi;                      // 1: the epilogue applies ConvEpi (per-channel affine / ReLU / fp32 side output)
    int dyn;                      // 1: tiles are drawn from the global counter (dynamic scheduler), 0: static stride
    int wres;                     // pair W-halo kernel, Cin == 64: this CTA's 27 x 32 weight rows stay resident in shared memory
    int sched_depth;              // tile-ring slots in use: 2 when dynamic (a CTA holds at most one tile it has not started), 4 when static
    int tiles_w, tiles_h, tiles_d, tiles_n, m_tiles, n_tiles;   // tile index: sample tile fastest, then w, h, d
    int kc;                       // Cin / 64
    int kj;                       // K16 MMAs per 64-channel slice: 4, fewer when the tensor has < 64 channels (the rest is zero fill)
    int stages;
    int nout;                     // epilogue staging buffers (1 or 2)
};

// Bit t set: along this axis, the input box of a tile whose first output coordinate is o0 (tile extent tl) intersects
// [0, extent) for tap t.  A tap whose box lies entirely in the zero padding contributes nothing: its K-steps are skipped
// (with dilation 4 on a 16^3 grid that is ~1/6 of the K-steps).
__device__ __forceinline__ uint32_t tap_axis_mask(int o0, int tl, int extent, int k, int stride, int pad, int dil) {
    uint32_t m = 0;
    for (int t = 0; t < k; ++t) {
        const int lo = o0 * stride - pad + t * dil, hi = lo + (tl - 1) * stride;
        if (hi >= 0 && lo < extent) m |= 1u << t;
    }
    return m;
}

// Optional epilogue work on the fp32 accumulators (all pointers may be NULL):
//   out_f32[voxel * f32_ld + c] = acc + f32_bias[c]      dense NDHWC fp32 side output (the tensor image_features.py:58-60 hooks:
//                                                        s_block1.conv2's raw output, bias included)
//   stored bf16 value           = act(acc * scale[c] + shift[c])   (scale NULL = 1, shift NULL = 0; relu != 0 applies ReLU):
//                                 a convolution bias (unet3d.py:37-40 / ConvTranspose3d :68), or - in eval mode - the whole
//                                 BatchNorm3d + ReLU that follows the convolution, folded into the producing kernel.
// The BatchNorm statistics (stats_partials) are always those of the STORED values.
//   head_out[n][k][d][h][w]     = head_b[k] + sum_c bf16(stored value)[c] * head_w[k][c]  for voxels inside the crop (hD,hH,hW):
//                                 a 1x1x1 convolution of the 64 output channels (unet3d.py:72 conv3) and the crop back of
//                                 unet3d.py:126-135 folded into the epilogue (Cout == 64, single-CTA kernel); with y == NULL the
//                                 activation itself is never written.
struct ConvEpi {
    const float* scale;
    const float* shift;
    const float* f32_bias;
    float* out_f32;
    long long f32_ld;
    int relu;
    const float* head_w;
    const float* head_b;
    float* head_out;
    int head_k, hD, hH, hW;
    int nostore;
    int cin_tensor;               // host only: channels the input tensor really has (< Cin: the rest of every K slice is TMA zero fill)
};
constexpr int kHeadMaxK = 8;

constexpr int kConvThreads = 288;          // 9 warps: 0, 2, 3 TMA producers, 1 MMA issuer, 4-7 epilogue, 8 tile scheduler
constexpr int kConvProducers = 3;          // warps 0, 2, 3
// Dynamic tile scheduler.  The kernels are persistent (one CTA or CTA pair per SM), but WHICH tiles a CTA computes is decided
// at run time: warp 8 draws tile indices from a global counter (atomicAdd) and hands them to the other roles through a small
// shared-memory ring (sched_tile[] + full / empty mbarriers), a few tiles ahead of the consumers.  A CTA that starts late -
// because another kernel (a weight-gradient GEMM on the side stream, an NCCL all-reduce) still holds its SM - simply finds fewer
// tiles left, and tiles that skip padding taps no longer unbalance a static stride.  g.dyn == 0 (short uniform tiles: 1x1x1
// convolutions) keeps the static assignment tile = blockIdx.x + k * gridDim.x, computed locally by every role without touching
// the ring.  The counter pair {next, done} resets itself: the last CTA to leave zeroes it, so a launch (or a CUDA-graph replay
// of it) always finds it at zero.
constexpr int kSchedSlots = 4;           // ring capacity; g.sched_depth (<= kSchedSlots) slots are used
constexpr int kSchedConsumers = 8;         // 3 producer warps + MMA warp + 4 epilogue warps arrive on a slot's empty barrier
// A producer re-enters the ring every nprod stages and waits on a PARITY, so it must never be two phases ahead of a
// slot: that needs stages >= active producers (nprod = min(kConvProducers, stages)).
constexpr int kATileBytes = 128 * 128;     // 128 voxels x 64 bf16
constexpr int kStageOutBytes = 128 * 128;  // epilogue staging: 128 voxels x 64 bf16

// KS = K-steps per pipeline stage.  For narrow N tiles (BN = 64 / 128, MMA time 132 / 264 cycles per K-step) the per-stage
// handshakes and copy issues weigh as much as the MMAs, so a stage carries KS input boxes
// and ONE weight box covering the KS consecutive K-slices (weights viewed as [64 ci][co][K-slice], see the host code).
//
// WH ("w halo", 3x3x3 unit-stride undilated convolutions of 64 channels, KS == 3): the three kw taps of a (kd, kh) pair
// read the SAME input box, loaded once with a halo of two voxels in W (10 x 4 x 4 voxels, tile 8 x 4 x 4).  UMMA applies
// the 128-byte swizzle to absolute shared-memory address bits, so tap kw's operand is simply the box start plus kw rows
// (128 bytes each), with the 8-row groups (one W line each) 10 rows = 1280 bytes apart.  27 input boxes per tile -> 9, and
// 2.4x less L2->SM traffic, which is what bounds the 64-channel layers.
constexpr int kHaloTileBytes = 160 * 128;  // 10 x 4 x 4 voxels x 64 bf16

// ConvEpi on 32 accumulator columns (channels cb .. cb+31) of this thread's output voxel; coefficients come from shared memory
// (every lane reads the same address: broadcast).
__device__ __forceinline__ void conv_epilogue_affine(uint32_t (&v)[32], int cb, const ConvEpi& ep, const float* ep_sc, const float* ep_sh,
                                                     const float* ep_fb, bool row_ok, long long row_vox) {
    if (ep.out_f32 && row_ok) {
        float* dst = ep.out_f32 + row_vox * ep.f32_ld + cb;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(ep_fb + cb + j);
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y,
                                                              __uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 sc = *reinterpret_cast<const float4*>(ep_sc + cb + j), sh = *reinterpret_cast<const float4*>(ep_sh + cb + j);
        float f0 = fmaf(__uint_as_float(v[j]), sc.x, sh.x), f1 = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
        float f2 = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z), f3 = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
        if (ep.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
        v[j] = __float_as_uint(f0); v[j + 1] = __float_as_uint(f1); v[j + 2] = __float_as_uint(f2); v[j + 3] = __float_as_uint(f3);
    }
}

// EPI: compile the ConvEpi epilogue in (affine / ReLU / fp32 side output / fused head).  The plain instantiation keeps the
// epilogue of the training path free of it: on ResNet3D-50 (57 convolutions, most of them short 1x1x1 tiles whose time IS the
// epilogue) the run-time-switched version cost 2.7 % of the step.
template <int BN, int KS, bool WH = false, bool EPI = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const ConvGeom g, float* __restrict__ stats_partials, const ConvEpi ep,
                    unsigned int* __restrict__ sched_counter) {
    pdl_launch_dependents();
    static_assert(!WH || KS == 3, "the W-halo variant stages the three kw taps of one (kd, kh) pair");
    constexpr int B_TILE = BN * 128;
    constexpr int A_REGION = WH ? kHaloTileBytes : KS * kATileBytes;
    constexpr int STAGE = A_REGION + KS * B_TILE;           // input tiles (or one halo box), then KS weight tiles
    constexpr uint32_t IDESC = umma_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t TMEM_COLS = 2 * BN;      // two accumulator buffers (128, 256 or 512 columns)

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* sm = smem_raw + (base - raw);
    const int S = g.stages;
    const uint32_t stage0 = base;
    const uint32_t out0 = base + (uint32_t)S * STAGE;      // nout x 16 KB epilogue staging
    unsigned char* tail = sm + (size_t)S * STAGE + (size_t)g.nout * kStageOutBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);    // full[S], empty[S], tfull[2], tempty[2], sfull[4], sempty[4]
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull0 = empty0 + 8 * S, tempty0 = tfull0 + 16;
    const uint32_t sfull0 = tempty0 + 16, sempty0 = sfull0 + 8 * kSchedSlots;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 4 + 2 * kSchedSlots);
    volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_ptr_s + 4);   // [kSchedSlots]
    float* st_sum = reinterpret_cast<float*>(tmem_ptr_s + 8);     // [2][Cout] (two row halves)
    float* st_sq = st_sum + 2 * g.Cout;                           // [2][Cout]
    float* ep_sc = st_sq + 2 * g.Cout;                            // epilogue scale / shift / fp32 bias, [Cout] each (only when g.epi)
    float* ep_sh = ep_sc + g.Cout;
    float* ep_fb = ep_sh + g.Cout;
    float* ep_hw = ep_fb + g.Cout;                                // fused head weights [head_k][64] (only with ep.head_out)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        for (int a = 0; a < kSchedSlots; ++a) { mbar_init(sfull0 + 8 * a, 1); mbar_init(sempty0 + 8 * a, kSchedConsumers); }
        mbar_fence_init();
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr_s), TMEM_COLS);
    for (int i = threadIdx.x; i < 4 * g.Cout; i += kConvThreads) st_sum[i] = 0.f;   // st_sum and st_sq are contiguous
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    const int total_tiles = g.m_tiles * g.n_tiles;
    const int taps = g.kd * g.kh * g.kw;
    const int ksteps = taps * g.kc;
    // consumer side of the tile ring: every role warp walks the same sequence of tile indices (>= total_tiles: no more work)
    uint32_t sc_slot = 0, sc_ph = 0;
    int sc_static = (int)blockIdx.x;                        // static mode: no ring traffic at all (the MMA issuer's loop is a critical path)
    auto next_tile = [&]() -> int {
        if (!g.dyn) { const int t = sc_static; sc_static += (int)gridDim.x; return t; }
        mbar_wait(sfull0 + 8 * sc_slot, sc_ph);
        const int t = sched_tile[sc_slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty0 + 8 * sc_slot);
        if (++sc_slot == (uint32_t)g.sched_depth) { sc_slot = 0; sc_ph ^= 1; }
        return t;
    };

    if (warp == 8) {
        // ============================ tile scheduler (dynamic mode only) ============================
        uint32_t slot = 0, ph = 0;
        for (; g.dyn;) {
            mbar_wait(sempty0 + 8 * slot, ph ^ 1);
            int t = 0;
            if (lane == 0) {
                t = (int)atomicAdd(sched_counter, 1u);
                sched_tile[slot] = t;
                mbar_arrive(sfull0 + 8 * slot);             // release: the tile index is visible to the waiters
            }
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= total_tiles) break;
            if (++slot == (uint32_t)g.sched_depth) { slot = 0; ph ^= 1; }
        }
    } else if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers: stage i is issued by producer i % 3 ============================
        if (WH) {
            // W-halo variant: 9 x (Cin / 64) stages per tile, one per (kd, kh, channel slab); 3 producers, 4 ring slots (host guarantees it).  Producer
            // `me` owns kh == me of every kd, i.e. every third stage of the global stage sequence: no per-tap loop overhead in
            // these single-thread loops, which otherwise cost more than the 384 MMA cycles of a stage.
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t G = me;
            for (;;) {
                const int tile = next_tile();
                if (tile >= total_tiles) break;
                int r = tile;                               // n_tiles == 1 (Cout == 64), tn == 1
                const int n = r % g.tiles_n; r /= g.tiles_n;
                const int wt = r % g.tiles_w; r /= g.tiles_w;
                const int ht = r % g.tiles_h; r /= g.tiles_h;
                const int dt = r;
                const int w0 = wt * 8 - 1, h0 = ht * 4 - 1 + (int)me, d0 = dt * 4 - 1;
                // stage order (kd, 64-channel slab, kh): kh stays the fastest index, so producer `me` keeps kh == me; the weight
                // box {64 ci, 64 co, 1 slab, 3 kw taps} of the 4-D weight view lands as three consecutive K-major B tiles
                for (int a = 0; a < 3; ++a)
                    for (int cc = 0; cc < g.kc; ++cc, G += 3) {
                        const uint32_t s = G & 3u, ph = (G >> 2) & 1u;
                        const uint32_t sa = stage0 + s * STAGE;
                        mbar_wait(empty0 + 8 * s, ph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)STAGE);
                            tma_load_4d(sa + A_REGION, &tmB, full0 + 8 * s, 0, 0, cc, (a * 3 + (int)me) * 3);
                            tma_load_5d(sa, &tmA, full0 + 8 * s, cc * 64, w0, h0, d0 + a, n);
                        }
                        __syncwarp();
                    }
            }
        } else {
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t s = 0, ph = 0, turn = 0;
            const uint32_t nprod = (uint32_t)min(kConvProducers, S);     // see the ring invariant above: stages >= active producers
            for (;;) {
                const int tile = next_tile();
                if (tile >= total_tiles) break;
                const int nt = tile % g.n_tiles, mt = tile / g.n_tiles;
                int r = mt;
                const int n = (r % g.tiles_n) * g.tn; r /= g.tiles_n;
                const int wt = r % g.tiles_w; r /= g.tiles_w;
                const int ht = r % g.tiles_h; r /= g.tiles_h;
                const int dt = r;
                const int w0 = wt * g.tw * g.stride - g.pad, h0 = ht * g.th * g.stride - g.pad,
                          d0 = dt * g.td * g.stride - g.pad;
                uint32_t mw = 0xffffffffu, mh = 0xffffffffu, md = 0xffffffffu;
                if (KS == 1) {                              // tap skipping needs per-K-step weight boxes
                    mw = tap_axis_mask(wt * g.tw, g.tw, g.W, g.kw, g.stride, g.pad, g.dil);
                    mh = tap_axis_mask(ht * g.th, g.th, g.H, g.kh, g.stride, g.pad, g.dil);
                    md = tap_axis_mask(dt * g.td, g.td, g.D, g.kd, g.stride, g.pad, g.dil);
                    if (!mw || !mh || !md) mw = mh = md = 0xffffffffu;      // degenerate tile: run everything (all zeros)
                }
                int tap = 0, kslice = 0, gpos = 0;          // gpos: position inside the current stage (0..KS-1)
                for (int a = 0; a < g.kd; ++a)
                    for (int b = 0; b < g.kh; ++b)
                        for (int c = 0; c < g.kw; ++c, ++tap) {
                            if (!((md >> a) & (mh >> b) & (mw >> c) & 1u)) { kslice += g.kc; continue; }
                            for (int cc = 0; cc < g.kc; ++cc, ++kslice) {
                                const uint32_t sa = stage0 + s * STAGE;
                                if (turn == me) {
                                    if (gpos == 0) mbar_wait(empty0 + 8 * s, ph ^ 1);
                                    if (elect_one()) {
                                        if (gpos == 0) {
                                            const int nv = min(KS, ksteps - kslice);    // K-steps in this stage (KS > 1: no skipping)
                                            mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(nv * kATileBytes + KS * B_TILE));
                                            tma_load_3d(sa + A_REGION, &tmB, full0 + 8 * s, 0, nt * BN, kslice);
                                        }
                                        tma_load_5d(sa + gpos * kATileBytes, &tmA, full0 + 8 * s, cc * 64, w0 + c * g.dil, h0 + b * g.dil,
                                                    d0 + a * g.dil, n);
                                    }
                                    __syncwarp();
                                }
                                if (++gpos == KS || kslice + 1 == ksteps) {
                                    gpos = 0;
                                    if (++turn == nprod) turn = 0;
                                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                                }
                            }
                        }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer (whole warp walks the loop, one elected lane issues) ============================
        {
            uint32_t s = 0, ph = 0, it = 0, nmma = 0;
            const int kj = g.kj;
            for (;; ++it) {
                const int tile = next_tile();
                if (tile >= total_tiles) break;
                const uint32_t acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(tempty0 + 8 * acc, aph ^ 1);            // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                // same tap selection as the producer
                int ksteps_t = ksteps;
                if (KS == 1) {
                    int r = tile / g.n_tiles / g.tiles_n;
                    const int wt = r % g.tiles_w; r /= g.tiles_w;
                    const int ht = r % g.tiles_h; r /= g.tiles_h;
                    const int dt = r;
                    const uint32_t mw = tap_axis_mask(wt * g.tw, g.tw, g.W, g.kw, g.stride, g.pad, g.dil);
                    const uint32_t mh = tap_axis_mask(ht * g.th, g.th, g.H, g.kh, g.stride, g.pad, g.dil);
                    const uint32_t md = tap_axis_mask(dt * g.td, g.td, g.D, g.kd, g.stride, g.pad, g.dil);
                    ksteps_t = (!mw || !mh || !md) ? ksteps : __popc(mw) * __popc(mh) * __popc(md) * g.kc;
                }
                for (int k = 0; k < ksteps_t; k += KS) {
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = stage0 + s * STAGE;
                    const int nv = min(KS, ksteps_t - k);
                    if (elect_one()) {
                        // ONE descriptor pair per stage; every other operand is that pair plus a compile-time offset (a 64-bit
                        // immediate add in the uniform datapath).  The issuing thread is the kernel's narrowest pipe: at ~3.6
                        // cycles per instruction the 16 instructions per MMA of the first version (descriptor re-encoded per
                        // tile, constants re-materialised) cost more than a 128 x 128 x 16 MMA takes (32 cycles).
                        const uint32_t a0 = umma_desc_lo(sa, 16), b0 = umma_desc_lo(sa + A_REGION, 16);
                        constexpr uint32_t AH = umma_desc_hi_sw128(WH ? 1280 : 1024), BH = umma_desc_hi_sw128(1024);
                        constexpr uint32_t A_STEP = (WH ? 128 : kATileBytes) >> 4, B_STEP = B_TILE >> 4;
                        if (kj == 4) {
#pragma unroll
                            for (int q = 0; q < KS; ++q) {
                                if (q >= nv) break;
#pragma unroll
                                for (int j = 0; j < 4; ++j)       // 4 x K16 inside the 64-wide (128-byte) swizzled row
                                    umma_bf16_lohi(d_tmem, a0 + (q * A_STEP + 2 * j), AH, b0 + (q * B_STEP + 2 * j), BH, IDESC, (k | q | j) ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < KS; ++q) {
                                if (q >= nv) break;
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (j < kj) umma_bf16_lohi(d_tmem, a0 + (q * A_STEP + 2 * j), AH, b0 + (q * B_STEP + 2 * j), BH, IDESC, (k | q | j) ? 1u : 0u);
                            }
                        }
                        nmma += nv * kj;
                        umma_commit(empty0 + 8 * s);              // frees the smem slot when these MMAs retire
                    }
                    __syncwarp();
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit(tfull0 + 8 * acc);   // accumulator complete -> epilogue
                __syncwarp();
            }
            if (elect_one()) mma_count_flush(&g_mma_flops_igemm, nmma, 2u * 128u * BN * 16u);
        }
    } else if (warp >= 4 && warp < 8) {
        // ============================ epilogue: TMEM -> bf16 -> smem -> TMA store (+ BN statistics) ============================
        const int ew = warp - 4;                                  // == warp % 4: the TMEM lane quarter this warp may read
        const int et = threadIdx.x - 128;                         // 0..127
        const int row = ew * 32 + lane;                           // output voxel inside the tile == TMEM lane
        uint32_t it = 0, nstore = 0;
        if (EPI) {
            for (int c = et; c < g.Cout; c += 128) {
                ep_sc[c] = ep.scale ? ep.scale[c] : 1.f;
                ep_sh[c] = ep.shift ? ep.shift[c] : 0.f;
                ep_fb[c] = ep.f32_bias ? ep.f32_bias[c] : 0.f;
            }
            if (EPI && ep.head_out)
                for (int c = et; c < ep.head_k * 64; c += 128) ep_hw[c] = ep.head_w[c];
            named_bar_sync(2, 128);
        }
        for (;; ++it) {
            const int tile = next_tile();
            if (tile >= total_tiles) break;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            const int nt = tile % g.n_tiles, mt = tile / g.n_tiles;
            int r = mt;
            const int n = (r % g.tiles_n) * g.tn; r /= g.tiles_n;
            const int wt = r % g.tiles_w; r /= g.tiles_w;
            const int ht = r % g.tiles_h; r /= g.tiles_h;
            const int dt = r;
            const int w0 = wt * g.tw, h0 = ht * g.th, d0 = dt * g.td;
            const int vw = min(g.tw, g.Wo - w0), vh = min(g.th, g.Ho - h0), vd = min(g.td, g.Do - d0);   // valid extent
            const int rwi = row & (g.tw - 1), rhi = (row >> g.lw) & (g.th - 1), rdi = (row >> (g.lw + g.lh)) & (g.td - 1),
                      rni = row >> (g.lw + g.lh + g.ltd);
            const bool row_ok = rwi < vw && rhi < vh && rdi < vd;
            const long long row_vox = (((long long)(n + rni) * g.Do + d0 + rdi) * g.Ho + h0 + rhi) * g.Wo + w0 + rwi;

            mbar_wait(tfull0 + 8 * acc, aph);
            tc_fence_after();
            float hacc[kHeadMaxK];
#pragma unroll
            for (int k = 0; k < kHeadMaxK; ++k) hacc[k] = 0.f;
            for (int sub = 0; sub < BN / 64; ++sub, ++nstore) {
                const uint32_t ob = out0 + (g.nout == 2 ? (nstore & 1) : 0u) * kStageOutBytes;
                if (!(EPI && ep.nostore)) {
                    if (et == 0) {                                // the store that last used this buffer has finished reading it
                        if (g.nout == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
                    }
                    named_bar_sync(2, 128);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + sub * 64 + half * 32, v);
                    tmem_ld_wait();
                    if (EPI) conv_epilogue_affine(v, nt * BN + sub * 64 + half * 32, ep, ep_sc, ep_sh, ep_fb, row_ok, row_vox);
                    if (EPI && BN == 64 && ep.head_out) {
                        // 1x1x1 head on the values as they would be stored (rounded to bf16), fp32 accumulation
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j]))));
#pragma unroll
                        for (int k = 0; k < kHeadMaxK; ++k) {
                            if (k < ep.head_k) {
                                const float4* hw4 = reinterpret_cast<const float4*>(ep_hw + k * 64 + half * 32);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 w4 = hw4[j];
                                    hacc[k] = fmaf(__uint_as_float(v[4 * j]), w4.x, hacc[k]); hacc[k] = fmaf(__uint_as_float(v[4 * j + 1]), w4.y, hacc[k]);
                                    hacc[k] = fmaf(__uint_as_float(v[4 * j + 2]), w4.z, hacc[k]); hacc[k] = fmaf(__uint_as_float(v[4 * j + 3]), w4.w, hacc[k]);
                                }
                            }
                        }
                    }
                    if (EPI && ep.nostore) continue;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t p0 = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
                        const uint32_t p1 = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                        const uint32_t p2 = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                        const uint32_t p3 = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                        const uint32_t chunk = (uint32_t)(half * 4 + q) ^ (uint32_t)(row & 7);      // SWIZZLE_128B
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ob + row * 128 + chunk * 16), "r"(p0),
                                     "r"(p1), "r"(p2), "r"(p3)
                                     : "memory");
                    }
                }
                if (EPI && BN == 64 && ep.head_out && row_ok) {
                    const int od = d0 + rdi, oh = h0 + rhi, ow = w0 + rwi;
                    if (od < ep.hD && oh < ep.hH && ow < ep.hW) {
#pragma unroll
                        for (int k = 0; k < kHeadMaxK; ++k)
                            if (k < ep.head_k)
                                ep.head_out[(((long long)(n + rni) * ep.head_k + k) * ep.hD + od) * ((long long)ep.hH * ep.hW) + (long long)oh * ep.hW + ow] =
                                    hacc[k] + __ldg(ep.head_b + k);
                    }
                }
                if (EPI && ep.nostore) continue;
                fence_proxy_async_smem();
                named_bar_sync(2, 128);
                if (et == 0) {
                    tma_store_5d(&tmC, ob, nt * BN + sub * 64, w0, h0, d0, n);
                    tma_store_commit();
                }
                if (stats_partials) {
                    // thread -> (channel c, row half): sum and sum of squares of the stored bf16 values over valid voxels
                    const int c = et & 63, hf = et >> 6;
                    float sum = 0.f, sq = 0.f;
                    const unsigned char* obp = sm + (ob - base);
                    for (int rr = hf * 64; rr < hf * 64 + 64; ++rr) {
                        const int wi = rr & (g.tw - 1), hi = (rr >> g.lw) & (g.th - 1), di = (rr >> (g.lw + g.lh)) & (g.td - 1);
                        if (wi < vw && hi < vh && di < vd) {
                            const uint32_t off = rr * 128 + (((uint32_t)(c >> 3) ^ (uint32_t)(rr & 7)) << 4) + (c & 7) * 2;
                            const float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(obp + off));
                            sum += x;
                            sq += x * x;
                        }
                    }
                    const int ch = nt * BN + sub * 64 + c;
                    st_sum[hf * g.Cout + ch] += sum;              // single owner thread per entry
                    st_sq[hf * g.Cout + ch] += sq;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        }
        if (et == 0) tma_store_wait<0>();
        if (stats_partials) {
            named_bar_sync(2, 128);
            for (int ch = et; ch < g.Cout; ch += 128) {
                stats_partials[((size_t)blockIdx.x * g.Cout + ch) * 2 + 0] = st_sum[ch] + st_sum[g.Cout + ch];
                stats_partials[((size_t)blockIdx.x * g.Cout + ch) * 2 + 1] = st_sq[ch] + st_sq[g.Cout + ch];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
    if (g.dyn && threadIdx.x == 0) {
        // every CTA has drawn its last (out-of-range) index by now; the last one to leave re-arms the counter pair
        if (atomicAdd(sched_counter + 1, 1u) == gridDim.x - 1) { sched_counter[0] = 0u; sched_counter[1] = 0u; __threadfence(); }
    }
}

// ===============================================================================================================
// CTA-pair variant (tcgen05 cta_group::2) for 256-channel N tiles.  A cluster of two CTAs computes a 256-voxel x 256-channel
// super tile with ONE MMA stream issued by the leader: every CTA stages its own 128 voxels of A and only HALF of the
// weight rows, so the shared-memory traffic per MAC (TMA writes + MMA operand reads), which bounds the single-CTA kernel
// at BN = 256, drops by a third.  Protocol (after CUTLASS' sm100 2-SM kernels): per-CTA `empty` barriers released by a
// multicast tcgen05.commit; the leader's `full` barrier collects the TMA bytes of both CTAs; accumulator-full is multicast
// to both epilogues, accumulator-empty is collected on the leader from both.
// ===============================================================================================================
//
// WH = true: the W-halo formulation of the single-CTA kernel (3x3x3, unit stride, Cout = 64, Cin = 64 k) on a CTA pair: a 256-voxel x
// 64-channel super tile, every CTA stages its own 10 x 4 x 4 halo box per (kd, kh, channel slab) and only 32 of the 64 weight rows
// of the three kw taps.  The single-CTA N = 64 kernel is bound by shared-memory bandwidth (44 KB of TMA writes + 72 KB of operand
// reads per 384 MMA cycles); here a CTA moves 32 + 60 KB per stage.
//
// BNP = 128 (WH = false): 128-channel N tiles on the pair.  The single-CTA 128 x 128 tile moves 32 KB per 256 MMA cycles and runs
// AT the L2 -> SM limit (17.5 TB/s, tensor pipe 55 %: profiles/r02_conv_gen128_ncu_full.csv); the pair stages 16 KB of A and
// only 8 KB of B per CTA for the same MMA time.
template <bool WH, bool EPI = false, int BNP = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv3d_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                         const __grid_constant__ CUtensorMap tmC, const ConvGeom g, float* __restrict__ stats_partials, const ConvEpi ep,
                         unsigned int* __restrict__ sched_counter) {
    pdl_launch_dependents();
    constexpr int BN = BNP ? BNP : (WH ? 64 : 256);
    static_assert(!WH || BN == 64, "the W-halo variant is for 64-channel tiles");
    constexpr int B_TAP = (BN / 2) * 128;                   // this CTA's half of one tap's weight rows
    constexpr int B_HALF = (WH ? 3 : 1) * B_TAP;
    constexpr int A_REGION = WH ? kHaloTileBytes : kATileBytes;
    constexpr int STAGE = A_REGION + B_HALF;                // 32 KB per CTA per stage (either variant)
    constexpr uint32_t IDESC = umma_idesc_bf16(256, BN, 0, 0);
    constexpr uint32_t TMEM_COLS = 2 * BN;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    const int S = g.stages;
    // Resident weights (WH, Cin == 64): the kernel is bound by the L2 -> SM fabric (~39 B per cycle per SM: 288 KB per tile at
    // 7.5 k cycles against 3.5 k MMA cycles), and 108 of those 288 KB are the SAME weight rows fetched again for every tile.  With a
    // single channel slab the CTA's half of all 27 taps (27 x 32 rows x 128 B = 108 KB) fits beside four 20 KB input stages: it is
    // loaded once per CTA and a stage carries only the halo box.
    const bool wres = WH && g.wres;
    const uint32_t stage_sz = wres ? (uint32_t)A_REGION : (uint32_t)STAGE;
    const uint32_t wres_bytes = wres ? 27u * B_TAP : 0u;
    const uint32_t stage0 = base;
    const uint32_t wres0 = base + (uint32_t)S * stage_sz;
    const uint32_t out0 = wres0 + wres_bytes;
    unsigned char* tail = sm + (size_t)S * stage_sz + wres_bytes + (size_t)g.nout * kStageOutBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tfull0 = empty0 + 8 * S, tempty0 = tfull0 + 16;
    const uint32_t sfull0 = tempty0 + 16, sempty0 = sfull0 + 8 * kSchedSlots;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * S + 4 + 2 * kSchedSlots);
    volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_ptr_s + 4);   // [kSchedSlots]
    const uint32_t wfull = smem_u32(tmem_ptr_s + 2);             // mbarrier (8 bytes): the resident weights of BOTH CTAs have landed
    float* st_sum = reinterpret_cast<float*>(tmem_ptr_s + 8);
    float* st_sq = st_sum + 2 * g.Cout;
    float* ep_sc = st_sq + 2 * g.Cout;
    float* ep_sh = ep_sc + g.Cout;
    float* ep_fb = ep_sh + g.Cout;
    float* ep_hw = ep_fb + g.Cout;                                // fused head weights (WH variant, ep.head_out)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }   // 4 epilogue warps x 2 CTAs
        // tile ring: the leader's scheduler fills the slot in BOTH CTAs; the consumers of both CTAs release the LEADER's slot
        for (int a = 0; a < kSchedSlots; ++a) { mbar_init(sfull0 + 8 * a, 1); mbar_init(sempty0 + 8 * a, 2 * kSchedConsumers); }
        mbar_init(wfull, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmC);
    }
    if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_ptr_s), TMEM_COLS);
    for (int i = threadIdx.x; i < 4 * g.Cout; i += kConvThreads) st_sum[i] = 0.f;
    tc_fence_before();
    cluster_sync_all();                                     // both CTAs' barriers and TMEM exist before anyone signals the peer
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    pdl_wait();                                             // prologue done; global memory from here on (launch_pdl, common.cuh)

    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_super = (g.m_tiles + 1) >> 1;
    const int total_super = m_super * g.n_tiles;
    const int taps = g.kd * g.kh * g.kw;
    const int ksteps = taps * g.kc;

    // consumer side of the tile ring (see the single-CTA kernel): super-tile indices, the same sequence in both CTAs
    uint32_t sc_slot = 0, sc_ph = 0;
    int sc_static = pair;
    auto next_tile = [&]() -> int {
        if (!g.dyn) { const int t = sc_static; sc_static += n_pairs; return t; }
        // the index lives in shared memory (never in L1), so the CTA-scope wait is enough - the cluster-scope acquire would add
        // a CCTL.IVALL per tile and evict the epilogue's scale / shift vectors from L1
        mbar_wait(sfull0 + 8 * sc_slot, sc_ph);
        const int t = sched_tile[sc_slot];
        __syncwarp();
        if (lane == 0 && t >= 0) mbar_signal_remote(sempty0 + 8 * sc_slot, 0);   // t >= 0 always: the test makes the arrive wait for the load
        if (++sc_slot == (uint32_t)g.sched_depth) { sc_slot = 0; sc_ph ^= 1; }
        return t;
    };

    // tap validity of a super tile = union over its two voxel tiles (both CTAs and the MMA issuer must agree)
    auto super_masks = [&](int mp, uint32_t& mw, uint32_t& mh, uint32_t& md) {
        mw = mh = md = 0;
        for (int h = 0; h < 2; ++h) {
            int r = 2 * mp + h;
            if (r >= g.m_tiles) break;
            r /= g.tiles_n;
            const int wt = r % g.tiles_w; r /= g.tiles_w;
            const int ht = r % g.tiles_h; r /= g.tiles_h;
            const int dt = r;
            uint32_t a = tap_axis_mask(wt * g.tw, g.tw, g.W, g.kw, g.stride, g.pad, g.dil);
            uint32_t b = tap_axis_mask(ht * g.th, g.th, g.H, g.kh, g.stride, g.pad, g.dil);
            uint32_t c = tap_axis_mask(dt * g.td, g.td, g.D, g.kd, g.stride, g.pad, g.dil);
            if (!a || !b || !c) a = b = c = 0xffffffffu;
            mw |= a; mh |= b; md |= c;
        }
    };

    if (warp == 8) {
        // ============================ tile scheduler: leader CTA only, publishes to both CTAs ============================
        if (leader) {
            uint32_t slot = 0, ph = 0;
            for (; g.dyn;) {
                mbar_wait(sempty0 + 8 * slot, ph ^ 1);
                int t = 0;
                if (lane == 0) {
                    t = (int)atomicAdd(sched_counter, 1u);
                    sched_tile[slot] = t;
                    st_shared_remote_u32(smem_u32(const_cast<int*>(sched_tile + slot)), 1, (uint32_t)t);
                    mbar_arrive(sfull0 + 8 * slot);
                    mbar_arrive_remote(sfull0 + 8 * slot, 1);      // release.cluster: the peer sees the index written above
                }
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t >= total_super) break;
                if (++slot == (uint32_t)g.sched_depth) { slot = 0; ph ^= 1; }
            }
        }
    } else if (warp == 0 || warp == 2 || warp == 3) {
        // ============================ TMA producers (in both CTAs) ============================
        {
            const uint32_t me = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t s = 0, ph = 0, turn = 0;
            const uint32_t nprod = (uint32_t)min(kConvProducers, S);     // see the ring invariant above: stages >= active producers
            if (wres && me == 0) {
                // once per CTA: this CTA's 32 rows of all 27 taps (nine boxes of three kw taps); the leader's barrier counts both CTAs
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(wfull, 2u * 27u * B_TAP);
                    for (int t = 0; t < 9; ++t)
                        tma_load_4d_2sm(wres0 + (uint32_t)t * 3u * B_TAP, &tmBh, wfull, 0, (int)rank * (BN / 2), 0, t * 3);
                }
                __syncwarp();
            }
            for (; WH;) {
                // W-halo stages (kd, channel slab, kh): one halo box + this CTA's 32 weight rows of the three kw taps
                const int st = next_tile();
                if (st >= total_super) break;
                int r = 2 * st + (int)rank;                 // n_tiles == 1, tn == 1; the phantom tile past the end loads zeros
                const int n = r % g.tiles_n; r /= g.tiles_n;
                const int wt = r % g.tiles_w; r /= g.tiles_w;
                const int ht = r % g.tiles_h; r /= g.tiles_h;
                const int dt = r;
                const int w0 = wt * 8 - 1, h0 = ht * 4 - 1, d0 = dt * 4 - 1;
                for (int a = 0; a < 3; ++a)
                    for (int cc = 0; cc < g.kc; ++cc)
                        for (int b = 0; b < 3; ++b) {
                            if (turn == me) {
                                mbar_wait(empty0 + 8 * s, ph ^ 1);
                                if (elect_one()) {
                                    if (leader) mbar_arrive_expect_tx(full0 + 8 * s, 2 * stage_sz);
                                    const uint32_t sa = stage0 + s * stage_sz;
                                    tma_load_5d_2sm(sa, &tmA, full0 + 8 * s, cc * 64, w0, h0 + b, d0 + a, n);
                                    if (!wres) tma_load_4d_2sm(sa + A_REGION, &tmBh, full0 + 8 * s, 0, (int)rank * (BN / 2), cc, (a * 3 + b) * 3);
                                }
                                __syncwarp();
                            }
                            if (++turn == nprod) turn = 0;
                            if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                        }
            }
            for (; !WH;) {
                const int st = next_tile();
                if (st >= total_super) break;
                const int nt = st % g.n_tiles, mp = st / g.n_tiles;
                int r = 2 * mp + (int)rank;                 // this CTA's voxel tile (may be the phantom tile past the end:
                const int n = (r % g.tiles_n) * g.tn; r /= g.tiles_n;   // its d index is >= tiles_d, every box is out of bounds -> zeros)
                const int wt = r % g.tiles_w; r /= g.tiles_w;
                const int ht = r % g.tiles_h; r /= g.tiles_h;
                const int dt = r;
                const int w0 = wt * g.tw * g.stride - g.pad, h0 = ht * g.th * g.stride - g.pad, d0 = dt * g.td * g.stride - g.pad;
                uint32_t mw, mh, md;
                super_masks(mp, mw, mh, md);
                int tap = 0;
                for (int a = 0; a < g.kd; ++a)
                    for (int b = 0; b < g.kh; ++b)
                        for (int c = 0; c < g.kw; ++c, ++tap) {
                            if (!((md >> a) & (mh >> b) & (mw >> c) & 1u)) continue;
                            for (int cc = 0; cc < g.kc; ++cc) {
                                if (turn == me) {
                                    mbar_wait(empty0 + 8 * s, ph ^ 1);                  // my own smem slot is free
                                    if (elect_one()) {
                                        if (leader) mbar_arrive_expect_tx(full0 + 8 * s, 2 * STAGE);   // bytes of BOTH CTAs
                                        const uint32_t sa = stage0 + s * STAGE;
                                        tma_load_5d_2sm(sa, &tmA, full0 + 8 * s, cc * 64, w0 + c * g.dil, h0 + b * g.dil, d0 + a * g.dil, n);
                                        tma_load_3d_2sm(sa + A_REGION, &tmBh, full0 + 8 * s, 0, nt * BN + (int)rank * (BN / 2), tap * g.kc + cc);
                                    }
                                    __syncwarp();
                                }
                                if (++turn == nprod) turn = 0;
                                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                            }
                        }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer: leader CTA only (the peer's warp 1 still walks the tile ring) ============================
        if (!leader) {
            while (next_tile() < total_super) {}
        } else {
            uint32_t s = 0, ph = 0, it = 0, nmma = 0;
            const int kj = g.kj;
            if (wres) { mbar_wait(wfull, 0); tc_fence_after(); }
            for (;; ++it) {
                const int st = next_tile();
                if (st >= total_super) break;
                const uint32_t acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(tempty0 + 8 * acc, aph ^ 1);            // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t mw = 0xffffffffu, mh = 0xffffffffu, md = 0xffffffffu;
                if (!WH) super_masks(st / g.n_tiles, mw, mh, md);
                const int ksteps_t = WH ? 9 * g.kc : ((mw == 0xffffffffu) ? ksteps : __popc(mw) * __popc(mh) * __popc(md) * g.kc);
                for (int k = 0; k < ksteps_t; ++k) {
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = stage0 + s * (WH ? stage_sz : (uint32_t)STAGE);
                    if (elect_one()) {
                        // one descriptor pair per stage + compile-time offsets (see the single-CTA kernel's issuer)
                        if (WH) {
                            // stage k of a tile = (kd, kh) = (k / 3, k % 3) when the weights are resident (one channel slab)
                            const uint32_t sb = wres ? wres0 + (uint32_t)k * 3u * B_TAP : sa + A_REGION;
                            const uint32_t a0 = umma_desc_lo(sa, 16), b0 = umma_desc_lo(sb, 16);
                            constexpr uint32_t AH = umma_desc_hi_sw128(1280), BH = umma_desc_hi_sw128(1024);
                            constexpr uint32_t B_STEP = B_TAP >> 4;
                            // kw taps: operand = the halo box shifted by q rows (8 descriptor units), W lines 10 rows apart
                            if (kj == 4) {
#pragma unroll
                                for (int q = 0; q < 3; ++q)
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        umma_bf16_2sm_lohi(d_tmem, a0 + (q * 8 + 2 * j), AH, b0 + (q * B_STEP + 2 * j), BH, IDESC, (k | q | j) ? 1u : 0u);
                            } else if (kj == 2) {
#pragma unroll
                                for (int q = 0; q < 3; ++q)
#pragma unroll
                                    for (int j = 0; j < 2; ++j)
                                        umma_bf16_2sm_lohi(d_tmem, a0 + (q * 8 + 2 * j), AH, b0 + (q * B_STEP + 2 * j), BH, IDESC, (k | q | j) ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int q = 0; q < 3; ++q)
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (j < kj) umma_bf16_2sm_lohi(d_tmem, a0 + (q * 8 + 2 * j), AH, b0 + (q * B_STEP + 2 * j), BH, IDESC, (k | q | j) ? 1u : 0u);
                            }
                            nmma += 3 * kj;
                        } else {
                            const uint32_t a0 = umma_desc_lo(sa, 16), b0 = umma_desc_lo(sa + A_REGION, 16);
                            constexpr uint32_t DH = umma_desc_hi_sw128(1024);
                            if (kj == 4) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) umma_bf16_2sm_lohi(d_tmem, a0 + 2 * j, DH, b0 + 2 * j, DH, IDESC, (k | j) ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (j < kj) umma_bf16_2sm_lohi(d_tmem, a0 + 2 * j, DH, b0 + 2 * j, DH, IDESC, (k | j) ? 1u : 0u);
                            }
                            nmma += kj;
                        }
                        umma_commit_2sm(empty0 + 8 * s, 3);       // frees the slot in both CTAs
                    }
                    __syncwarp();
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit_2sm(tfull0 + 8 * acc, 3);   // accumulator complete -> both epilogues
                __syncwarp();
            }
            if (elect_one()) mma_count_flush(&g_mma_flops_igemm, nmma, 2u * 256u * (uint32_t)BN * 16u);
        }
    } else if (warp >= 4 && warp < 8) {
        // ============================ epilogue (in both CTAs): own 128 accumulator rows ============================
        const int ew = warp - 4;
        const int et = threadIdx.x - 128;
        const int row = ew * 32 + lane;
        uint32_t it = 0, nstore = 0;
        if (EPI) {
            for (int c = et; c < g.Cout; c += 128) {
                ep_sc[c] = ep.scale ? ep.scale[c] : 1.f;
                ep_sh[c] = ep.shift ? ep.shift[c] : 0.f;
                ep_fb[c] = ep.f32_bias ? ep.f32_bias[c] : 0.f;
            }
            if (EPI && BN == 64 && ep.head_out)
                for (int c = et; c < ep.head_k * 64; c += 128) ep_hw[c] = ep.head_w[c];
            named_bar_sync(2, 128);
        }
        for (;; ++it) {
            const int st = next_tile();
            if (st >= total_super) break;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            const int nt = st % g.n_tiles, mp = st / g.n_tiles;
            const int mt = 2 * mp + (int)rank;
            int r = mt;
            const int n = (r % g.tiles_n) * g.tn; r /= g.tiles_n;
            const int wt = r % g.tiles_w; r /= g.tiles_w;
            const int ht = r % g.tiles_h; r /= g.tiles_h;
            const int dt = r;
            const int w0 = wt * g.tw, h0 = ht * g.th, d0 = dt * g.td;
            const bool real = mt < g.m_tiles;
            const int vw = real ? min(g.tw, g.Wo - w0) : 0, vh = min(g.th, g.Ho - h0), vd = min(g.td, g.Do - d0);
            const int rwi = row & (g.tw - 1), rhi = (row >> g.lw) & (g.th - 1), rdi = (row >> (g.lw + g.lh)) & (g.td - 1),
                      rni = row >> (g.lw + g.lh + g.ltd);
            const bool row_ok = rwi < vw && rhi < vh && rdi < vd;
            const long long row_vox = (((long long)(n + rni) * g.Do + d0 + rdi) * g.Ho + h0 + rhi) * g.Wo + w0 + rwi;

            mbar_wait(tfull0 + 8 * acc, aph);
            tc_fence_after();
            float hacc[kHeadMaxK];
#pragma unroll
            for (int k = 0; k < kHeadMaxK; ++k) hacc[k] = 0.f;
            for (int sub = 0; sub < BN / 64; ++sub, ++nstore) {
                const uint32_t ob = out0 + (g.nout == 2 ? (nstore & 1) : 0u) * kStageOutBytes;
                if (!(EPI && ep.nostore)) {
                    if (et == 0) {
                        if (g.nout == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
                    }
                    named_bar_sync(2, 128);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + sub * 64 + half * 32, v);
                    tmem_ld_wait();
                    if (EPI) conv_epilogue_affine(v, nt * BN + sub * 64 + half * 32, ep, ep_sc, ep_sh, ep_fb, row_ok, row_vox);
                    if (EPI && BN == 64 && ep.head_out) {
                        // 1x1x1 head on the values as they would be stored (rounded to bf16), fp32 accumulation (see the single-CTA kernel)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j]))));
#pragma unroll
                        for (int k = 0; k < kHeadMaxK; ++k) {
                            if (k < ep.head_k) {
                                const float4* hw4 = reinterpret_cast<const float4*>(ep_hw + k * 64 + half * 32);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 w4 = hw4[j];
                                    hacc[k] = fmaf(__uint_as_float(v[4 * j]), w4.x, hacc[k]); hacc[k] = fmaf(__uint_as_float(v[4 * j + 1]), w4.y, hacc[k]);
                                    hacc[k] = fmaf(__uint_as_float(v[4 * j + 2]), w4.z, hacc[k]); hacc[k] = fmaf(__uint_as_float(v[4 * j + 3]), w4.w, hacc[k]);
                                }
                            }
                        }
                    }
                    if (EPI && ep.nostore) continue;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t p0 = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
                        const uint32_t p1 = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                        const uint32_t p2 = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                        const uint32_t p3 = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                        const uint32_t chunk = (uint32_t)(half * 4 + q) ^ (uint32_t)(row & 7);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ob + row * 128 + chunk * 16), "r"(p0),
                                     "r"(p1), "r"(p2), "r"(p3)
                                     : "memory");
                    }
                }
                if (EPI && BN == 64 && ep.head_out && row_ok) {
                    const int od = d0 + rdi, oh = h0 + rhi, ow = w0 + rwi;
                    if (od < ep.hD && oh < ep.hH && ow < ep.hW) {
#pragma unroll
                        for (int k = 0; k < kHeadMaxK; ++k)
                            if (k < ep.head_k)
                                ep.head_out[(((long long)(n + rni) * ep.head_k + k) * ep.hD + od) * ((long long)ep.hH * ep.hW) + (long long)oh * ep.hW + ow] =
                                    hacc[k] + __ldg(ep.head_b + k);
                    }
                }
                if (EPI && ep.nostore) continue;
                fence_proxy_async_smem();
                named_bar_sync(2, 128);
                if (et == 0 && real) {
                    tma_store_5d(&tmC, ob, nt * BN + sub * 64, w0, h0, d0, n);
                    tma_store_commit();
                }
                if (stats_partials) {
                    const int c = et & 63, hf = et >> 6;
                    float sum = 0.f, sq = 0.f;
                    const unsigned char* obp = sm + (ob - base);
                    for (int rr = hf * 64; rr < hf * 64 + 64; ++rr) {
                        const int wi = rr & (g.tw - 1), hi = (rr >> g.lw) & (g.th - 1), di = (rr >> (g.lw + g.lh)) & (g.td - 1);
                        if (wi < vw && hi < vh && di < vd) {
                            const uint32_t off = rr * 128 + (((uint32_t)(c >> 3) ^ (uint32_t)(rr & 7)) << 4) + (c & 7) * 2;
                            const float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(obp + off));
                            sum += x;
                            sq += x * x;
                        }
                    }
                    const int ch = nt * BN + sub * 64 + c;
                    st_sum[hf * g.Cout + ch] += sum;
                    st_sq[hf * g.Cout + ch] += sq;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_signal_remote(tempty0 + 8 * acc, 0);   // the leader's MMA issuer waits for both CTAs
        }
        if (et == 0) tma_store_wait<0>();
        if (stats_partials) {
            named_bar_sync(2, 128);
            for (int ch = et; ch < g.Cout; ch += 128) {
                stats_partials[((size_t)blockIdx.x * g.Cout + ch) * 2 + 0] = st_sum[ch] + st_sum[g.Cout + ch];
                stats_partials[((size_t)blockIdx.x * g.Cout + ch) * 2 + 1] = st_sq[ch] + st_sq[g.Cout + ch];
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();                                     // nobody frees TMEM / exits while the peer may still signal it
    if (warp == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    if (g.dyn && threadIdx.x == 0) {
        if (atomicAdd(sched_counter + 1, 1u) == gridDim.x - 1) { sched_counter[0] = 0u; sched_counter[1] = 0u; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* basep, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides) {
    return make_tmap_bf16_swz(out, basep, rank, dims, strides_bytes, box, elem_strides, 128);
}

int make_tmap_bf16_swz(CUtensorMap* out, const void* basep, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(MMAD_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(basep), dims, strides_bytes,
                     box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MMAD_ECUDA, "cuTensorMapEncodeTiled failed (CUresult " + std::to_string((int)r) + ")");
    return MMAD_OK;
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Output tile box tw x th x td x tn = 128 voxels (powers of two, tn samples deep): least padding waste, and for dilated
// layers the shape that leaves the fewest (tile, tap) pairs once taps entirely in the padding are skipped (4 x 4 x 4 x 2
// samples for dilation 4 on 16^3 instead of 16 x 1 x 8: 58 % instead of 83 % of the K-steps).
static void pick_tile(int N, int D, int H, int W, int Wo, int Ho, int Do, int k, int stride, int pad, int dil, int& tw, int& th, int& td,
                      int& tn) {
    pick_chunk(128, N, W, H, D, Wo, Ho, Do, k, stride, pad, dil, tw, th, td, tn);
}

// wres: resident weights of the pair W-halo kernel (27 taps x bn rows x 128 B kept once, stages carry only the halo box)
static int conv_smem_bytes(int bn, int ks, int stages, int nout, int cout, bool halo = false, bool epi = false, bool wres = false) {
    const int a_region = halo ? kHaloTileBytes : ks * kATileBytes;
    const int stage = wres ? a_region : a_region + ks * bn * 128;
    return 1024 + stages * stage + (wres ? 27 * bn * 128 : 0) + nout * kStageOutBytes + (2 * stages + 4 + 2 * kSchedSlots) * 8 + 32 +
           (epi ? 7 : 4) * cout * 4 + (epi ? kHeadMaxK * 64 * 4 : 0);
}

// {next, done} counter pairs of the dynamic tile scheduler: a pool per device, one pair per launch (round robin - two kernels
// that may be in flight together, e.g. on the main and the weight-gradient stream, never share a pair); zeroed once, then
// every kernel leaves its pair at zero.  Allocated on first use (before any CUDA-graph capture: capture follows warm-up).
constexpr int kSchedPool = 4096;
static unsigned int* sched_counter_slot() {
    static unsigned int* pool[64] = {};
    static std::atomic<unsigned> next{0};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    dev &= 63;
    if (!pool[dev]) {
        std::lock_guard<std::mutex> lock(mu);
        if (!pool[dev]) {
            unsigned int* p = nullptr;
            if (cudaMalloc((void**)&p, kSchedPool * 2 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
            if (cudaMemset(p, 0, kSchedPool * 2 * sizeof(unsigned int)) != cudaSuccess) { cudaFree(p); return nullptr; }
            pool[dev] = p;
        }
    }
    return pool[dev] + 2 * (next.fetch_add(1, std::memory_order_relaxed) % kSchedPool);
}
static bool use_dynamic_scheduler() {                    // on unless MMAD_CONV_DYN=0
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("MMAD_CONV_DYN"); mode = e ? atoi(e) : 1; }
    return mode != 0;
}

// W-halo variant (see the kernel): 3x3x3, unit stride, undilated, 64 -> 64 channels; on unless MMAD_CONV_HALO=0
static bool use_halo_kernel(int Cin, int Cout, int k, int stride, int dil) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("MMAD_CONV_HALO"); mode = e ? atoi(e) : 1; }
    return mode != 0 && Cin % 64 == 0 && Cout == 64 && k == 3 && stride == 1 && dil == 1;
}


// CTA-pair kernel for 256-channel N tiles: on unless MMAD_CONV_PAIR=0
static bool use_pair_kernel(int bn, long long m_tiles) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("MMAD_CONV_PAIR"); mode = e ? atoi(e) : 1; }
    // 128-channel tiles on the pair kernel too: on unless MMAD_CONV_PAIR128=0
    static int mode128 = -1;
    if (mode128 < 0) { const char* e = getenv("MMAD_CONV_PAIR128"); mode128 = e ? atoi(e) : 1; }
    return mode != 0 && (bn == 256 || (bn == 128 && mode128 != 0)) && m_tiles >= 2;
}

}  // namespace mmad

using namespace mmad;

extern "C" {

// Number of per-CTA statistic partials mmad_conv3d_fwd_bf16 writes for a given problem (== grid size).
int mmad_conv3d_stats_partials(int N, int D, int H, int W, int Cout, int k, int stride, int pad, int dil) {
    // NOTE: must mirror the grid computed in mmad_conv3d_fwd_bf16
    const int Do = (D + 2 * pad - dil * (k - 1) - 1) / stride + 1, Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1,
              Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    int tw = 0, th = 0, td = 0, tn = 1;
    pick_tile(N, D, H, W, Wo, Ho, Do, k, stride, pad, dil, tw, th, td, tn);
    const int bn = std::min(256, Cout);
    const long long tiles = (long long)(N / tn) * ((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((Do + td - 1) / td) * (Cout / bn);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long m_tiles = tiles / (Cout / bn);
    if (use_pair_kernel(bn, m_tiles)) return 2 * (int)std::min<long long>(((m_tiles + 1) / 2) * (Cout / bn), sms / 2);
    return (int)std::min<long long>(tiles, sms);
}

// Shared launcher.  (kd, kh, kw) taps, output extents and the output's voxel strides (in ELEMENTS; 0 = dense NDHWC) are explicit
// so that the phase convolutions of the stride-2 data gradient (rectangular kernels, outputs interleaved into dx) use the same
// kernels as the public forward entry point.
struct ConvOut { int Do, Ho, Wo; long long sw, sh, sd, sn; int standard; };   // standard: the extents are the convolution's own (only strides differ)
static int conv_fwd_impl(const void* x, const void* w, void* y, float* stats_partials, int N, int D, int H, int W, int Cin, int Cout,
                         int kd, int kh, int kw, int stride, int pad, int dil, const ConvOut* ov, void* stream, const ConvEpi* epi = nullptr) {
    const int k = std::max(kd, std::max(kh, kw));
    MMAD_CHECK_ARG(x && w && (y || (epi && epi->head_out)), "conv3d_fwd: null pointer");
    MMAD_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0, "conv3d_fwd: empty input");
    MMAD_CHECK_ARG(Cin % 64 == 0 && Cin >= 64, "conv3d_fwd: Cin must be a multiple of 64");
    MMAD_CHECK_ARG(Cout % 64 == 0 && Cout >= 64 && (Cout <= 256 || Cout % 256 == 0) && (Cout == 64 || Cout % 128 == 0) &&
                       Cout <= 2048,
                   "conv3d_fwd: Cout must be 64, 128, 256 or a multiple of 256 (<= 2048)");
    MMAD_CHECK_ARG(k >= 1 && k <= 7 && stride >= 1 && stride <= 2 && dil >= 1 && pad >= 0, "conv3d_fwd: bad kernel geometry");
    MMAD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0,
                   "conv3d_fwd: pointers must be 16-byte aligned");
    ConvGeom g = {};
    g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
    g.kd = kd; g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad; g.dil = dil;
    g.Do = ov ? ov->Do : (D + 2 * pad - dil * (kd - 1) - 1) / stride + 1;
    g.Ho = ov ? ov->Ho : (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1;
    g.Wo = ov ? ov->Wo : (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    MMAD_CHECK_ARG(g.Do > 0 && g.Ho > 0 && g.Wo > 0, "conv3d_fwd: empty output");
    int dev = 0, sms = 148;
    MMAD_CUDA(cudaGetDevice(&dev));
    MMAD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pick_tile(N, D, H, W, g.Wo, g.Ho, g.Do, k, stride, pad, dil, g.tw, g.th, g.td, g.tn);
    bool halo = false;
    ConvEpi ep = {};
    if (epi) ep = *epi;
    g.epi = (ep.scale || ep.shift || ep.out_f32 || ep.head_out) ? 1 : 0;
    ep.nostore = y ? 0 : 1;
    MMAD_CHECK_ARG(!ep.head_out || (Cout == 64 && ep.head_w && ep.head_b && ep.head_k >= 1 && ep.head_k <= kHeadMaxK),
                   "conv3d_fwd: the fused 1x1x1 head needs Cout == 64 and 1..8 classes");
    MMAD_CHECK_ARG(!ep.nostore || !stats_partials, "conv3d_fwd: statistics need the stored output");
    MMAD_CHECK_ARG(ep.cin_tensor == 0 || (ep.cin_tensor % 8 == 0 && ep.cin_tensor <= Cin && Cin == 64),
                   "conv3d_fwd: a narrower input tensor needs Cin == 64 and a multiple of 8 channels (16-byte rows)");
    MMAD_CHECK_ARG(!ep.out_f32 || (((uintptr_t)ep.out_f32 & 15) == 0 && ep.f32_ld % 4 == 0), "conv3d_fwd: fp32 side output must be 16-byte aligned");
    if ((!ov || ov->standard) && kd == kh && kh == kw && pad == 1 && use_halo_kernel(Cin, Cout, k, stride, dil)) {
        // the halo variant needs 8 x 4 x 4 tiles; mmad_conv3d_stats_partials (which does not know Cin) sizes the statistics
        // buffer from the regular tile, so only switch when both tilings fill every SM (grid == SM count either way)
        const long long reg = (long long)(N / g.tn) * ((g.Wo + g.tw - 1) / g.tw) * ((g.Ho + g.th - 1) / g.th) * ((g.Do + g.td - 1) / g.td);
        const long long hal = (long long)N * ((g.Wo + 7) / 8) * ((g.Ho + 3) / 4) * ((g.Do + 3) / 4);
        if (reg >= sms && hal >= sms) { halo = true; g.tw = 8; g.th = 4; g.td = 4; g.tn = 1; }
    }
    g.lw = ilog2(g.tw); g.lh = ilog2(g.th); g.ltd = ilog2(g.td);
    g.tiles_w = (g.Wo + g.tw - 1) / g.tw; g.tiles_h = (g.Ho + g.th - 1) / g.th; g.tiles_d = (g.Do + g.td - 1) / g.td;
    g.tiles_n = N / g.tn;
    g.m_tiles = g.tiles_n * g.tiles_w * g.tiles_h * g.tiles_d;
    const int bn = std::min(256, Cout);
    g.n_tiles = Cout / bn;
    g.kc = Cin / 64;
    g.kj = ep.cin_tensor > 0 ? (ep.cin_tensor + 15) / 16 : 4;          // K16 steps that can be non-zero
    static int halo_pair_mode = -1;                        // CTA-pair W-halo kernel: on unless MMAD_CONV_HALO_PAIR=0
    if (halo_pair_mode < 0) { const char* e = getenv("MMAD_CONV_HALO_PAIR"); halo_pair_mode = e ? atoi(e) : 1; }
    const bool halo_pair = halo && halo_pair_mode != 0 && g.m_tiles >= 2;
    const bool pairk = halo_pair || use_pair_kernel(bn, g.m_tiles);
    // resident weights (pair W-halo kernel, Cin == 64), with ONE epilogue staging buffer so that five (training) / four (ConvEpi)
    // input stages still fit: UNet3D eval forward 16.7 - 16.8 ms with, 17.3 - 17.8 ms without; ResNet3D-18 step neutral.
    // On unless MMAD_CONV_WRES=0.
    static int wres_mode = -1;
    if (wres_mode < 0) { const char* e = getenv("MMAD_CONV_WRES"); wres_mode = e ? atoi(e) : 1; }
    g.wres = (halo_pair && Cin == 64 && wres_mode != 0) ? 1 : 0;
    const int bn_stage = pairk ? bn / 2 : bn;              // weight rows a CTA stages per K-step
    static int ks_mode = -1;                               // MMAD_CONV_KS=1 forces one K-step per stage (tuning knob)
    if (ks_mode < 0) { const char* e = getenv("MMAD_CONV_KS"); ks_mode = e ? atoi(e) : 0; }
    // K-steps per stage (one weight box per stage); the pair kernel stages one K-step
    const int ks = halo ? 3 : ((ks_mode == 1 || pairk) ? 1 : (bn == 64 ? 4 : (bn == 128 ? 2 : 1)));
    g.nout = ((bn == 256 && !pairk) || g.wres) ? 1 : 2;   // resident weights: one staging buffer buys the fifth input stage
    int stages = 8;
    while (stages > 2 && conv_smem_bytes(bn_stage, ks, stages, g.nout, Cout, halo, g.epi != 0, g.wres != 0) > 227 * 1024) --stages;
    if (conv_smem_bytes(bn_stage, ks, stages, g.nout, Cout, halo, g.epi != 0, g.wres != 0) > 227 * 1024 || (stages < 3 && g.nout == 2 && ks > 1)) {
        g.nout = 1;                                        // trade the second epilogue buffer for pipeline depth
        stages = 8;
        while (stages > 2 && conv_smem_bytes(bn_stage, ks, stages, g.nout, Cout, halo, g.epi != 0, g.wres != 0) > 227 * 1024) --stages;
    }
    if (halo && !pairk) stages = 4;                        // the single-CTA halo producers assume a 4-slot ring (fits: 4 x 44 KB + 2 x 16 KB)
    g.stages = stages;
    const int smem = conv_smem_bytes(bn_stage, ks, stages, g.nout, Cout, halo, g.epi != 0, g.wres != 0);
    MMAD_CHECK_ARG(smem <= 227 * 1024, "conv3d_fwd: shared memory budget exceeded");

    CUtensorMap tmA, tmB, tmC;
    {
        // ct < Cin: the tensor has only ct channels per voxel (a 32-channel activation: 64-byte rows); the 64-wide box then
        // reaches past dimension 0 and TMA zero-fills the rest of the 128-byte row - the K slice is padded in flight, neither
        // stored nor fetched (half of the L2 -> SM bytes of that layer)
        const uint64_t ct = (uint64_t)(ep.cin_tensor > 0 ? ep.cin_tensor : Cin);
        const uint64_t dims[5] = {ct, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)N};
        const uint64_t str[4] = {ct * 2, (uint64_t)W * ct * 2, (uint64_t)H * W * ct * 2, (uint64_t)D * H * W * ct * 2};
        const uint32_t box[5] = {64, (uint32_t)(halo ? g.tw + 2 : g.tw * stride), (uint32_t)(g.th * stride), (uint32_t)(g.td * stride), (uint32_t)g.tn};
        const uint32_t es[5] = {1, (uint32_t)stride, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_tmap_bf16(&tmA, x, 5, dims, str, box, es);
        if (rc) return rc;
    }
    if (halo) {
        // weights [Cout][taps][Cin] viewed as [64 ci][co][ci slab][tap]: a box of one slab x three consecutive taps (the kw taps of a
        // (kd, kh) pair) lands as three consecutive canonical K-major B tiles
        const int taps = kd * kh * kw;
        const uint64_t dims[4] = {64, (uint64_t)Cout, (uint64_t)(Cin / 64), (uint64_t)taps};
        const uint64_t str[3] = {(uint64_t)taps * Cin * 2, 128, (uint64_t)Cin * 2};
        const uint32_t box[4] = {64, (uint32_t)bn_stage, 1, 3};
        const uint32_t es[4] = {1, 1, 1, 1};
        int rc = make_tmap_bf16(&tmB, w, 4, dims, str, box, es);
        if (rc) return rc;
    } else {
        // weights [Cout][taps][Cin] viewed as [64 ci][co][K-slice]: K-slice = tap * (Cin/64) + ci block, 128 bytes apart, so a
        // box of KS consecutive K-slices lands as KS consecutive canonical K-major B tiles
        const int taps = kd * kh * kw;
        const uint64_t dims[3] = {64, (uint64_t)Cout, (uint64_t)taps * (Cin / 64)};
        const uint64_t str[2] = {(uint64_t)taps * Cin * 2, 128};
        const uint32_t box[3] = {64, (uint32_t)bn_stage, (uint32_t)ks};
        const uint32_t es[3] = {1, 1, 1};
        int rc = make_tmap_bf16(&tmB, w, 3, dims, str, box, es);
        if (rc) return rc;
    }
    {
        const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
        const uint64_t dense[4] = {(uint64_t)Cout * 2, (uint64_t)g.Wo * Cout * 2, (uint64_t)g.Ho * g.Wo * Cout * 2,
                                   (uint64_t)g.Do * g.Ho * g.Wo * Cout * 2};
        const uint64_t strided[4] = {(uint64_t)(ov ? ov->sw : 0) * 2, (uint64_t)(ov ? ov->sh : 0) * 2, (uint64_t)(ov ? ov->sd : 0) * 2,
                                     (uint64_t)(ov ? ov->sn : 0) * 2};
        const uint64_t* str = (ov && ov->sw) ? strided : dense;
        const uint32_t box[5] = {64, (uint32_t)g.tw, (uint32_t)g.th, (uint32_t)g.td, (uint32_t)g.tn};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        int rc = make_tmap_bf16(&tmC, y ? y : const_cast<void*>(x), 5, dims, str, box, es);   // y == NULL: never stored (fused head)
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* sched = nullptr;
    // dynamic draws pay off where tiles are long and uneven (3x3x3 taps, padding skips, co-running kernels); the short uniform
    // tiles of a 1x1x1 convolution keep the static stride (measured on ResNet3D-50: the per-tile atomic costs 1 % there)
    // ... and layers with only a few tiles per CTA (ResNet3D-50's 3x3x3 convolutions at batch 8: 1.7 - 3.5) lose more to the draw's
    // start / end latency than they can gain: +0.6 ms on that network's 16.7 ms step.  Dynamic from 4 tiles per CTA (pair) on.
    const long long work_items = pairk ? (long long)((g.m_tiles + 1) / 2) * g.n_tiles : (long long)g.m_tiles * g.n_tiles;
    const long long ctas = pairk ? sms / 2 : sms;
    // The per-tile hand-off is a few plain shared-memory / mbarrier operations (mbar_signal_remote: no fences); with the fenced
    // arrives of the first version a UNet3D eval forward was 3.5 % SLOWER with dynamic draws than without (16.70 vs 16.14 ms),
    // now it is 2.4 % faster (15.63 vs 16.02 ms, batch 8).
    g.dyn = (use_dynamic_scheduler() && kd * kh * kw * g.kc >= 16 && work_items >= 4 * ctas) ? 1 : 0;
    // ring depth 2: a CTA holds at most one tile it has not started.  Deeper look-ahead was measured twice and lost both times
    // (ResNet3D-18 step 12.31 ms at depth 2, 13.62 ms at depth 4 for layers with >= 16 tiles per CTA; UNet3D eval 16.77 / 16.66 /
    // 16.97 ms at depth 2 / 3 / 4).  MMAD_CONV_SCHED_DEPTH overrides (2..4) for experiments.
    static int depth_knob = -1;
    if (depth_knob < 0) { const char* e = getenv("MMAD_CONV_SCHED_DEPTH"); depth_knob = e ? std::max(2, std::min(kSchedSlots, atoi(e))) : 2; }
    g.sched_depth = g.dyn ? depth_knob : kSchedSlots;
    if (g.dyn) {
        sched = sched_counter_slot();
        if (!sched) return fail(MMAD_ECUDA, "conv3d_fwd: cannot allocate the tile-scheduler counters");
    }
    if (pairk) {
        static DevOnce attr_done;
        if (attr_done.need()) {
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_pair_kernel<false, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_pair_kernel<false, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
        const int pairs = (int)std::min<long long>((long long)((g.m_tiles + 1) / 2) * g.n_tiles, sms / 2);
        const dim3 pgrid(2 * pairs);
        if (halo_pair && g.epi) launch_pdl(conv3d_igemm_pair_kernel<true, true>, pgrid, dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);
        else if (halo_pair) launch_pdl(conv3d_igemm_pair_kernel<true, false>, pgrid, dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);
        else if (bn == 128 && g.epi) launch_pdl(conv3d_igemm_pair_kernel<false, true, 128>, pgrid, dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);
        else if (bn == 128) launch_pdl(conv3d_igemm_pair_kernel<false, false, 128>, pgrid, dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);
        else if (g.epi) launch_pdl(conv3d_igemm_pair_kernel<false, true>, pgrid, dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);
        else launch_pdl(conv3d_igemm_pair_kernel<false, false>, pgrid, dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);
        MMAD_CUDA(cudaGetLastError());
        count_launch();
        return MMAD_OK;
    }
    const int grid = (int)std::min<long long>((long long)g.m_tiles * g.n_tiles, sms);
#define MMAD_CONV_LAUNCH1(...)                                                                                              \
    do {                                                                                                                    \
        static DevOnce attr_done;                                                                                           \
        if (attr_done.need()) {                                                                                             \
            MMAD_CUDA(cudaFuncSetAttribute(conv3d_igemm_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        }                                                                                                                   \
        launch_pdl(conv3d_igemm_kernel<__VA_ARGS__>, dim3(grid), dim3(kConvThreads), smem, st, tmA, tmB, tmC, g, stats_partials, ep, sched);    \
    } while (0)
#define MMAD_CONV_LAUNCH(BN_, KS_, WH_)                                                                                     \
    do {                                                                                                                    \
        if (g.epi) MMAD_CONV_LAUNCH1(BN_, KS_, WH_, true); else MMAD_CONV_LAUNCH1(BN_, KS_, WH_, false);                    \
    } while (0)
    if (halo) MMAD_CONV_LAUNCH(64, 3, true);
    else if (bn == 64 && ks == 4) MMAD_CONV_LAUNCH(64, 4, false);
    else if (bn == 64) MMAD_CONV_LAUNCH(64, 1, false);
    else if (bn == 128 && ks == 2) MMAD_CONV_LAUNCH(128, 2, false);
    else if (bn == 128) MMAD_CONV_LAUNCH(128, 1, false);
    else MMAD_CONV_LAUNCH(256, 1, false);
#undef MMAD_CONV_LAUNCH1
#undef MMAD_CONV_LAUNCH
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    return MMAD_OK;
}

int mmad_conv3d_fwd_bf16(const void* x, const void* w, void* y, float* stats_partials, int N, int D, int H, int W,
                         int Cin, int Cout, int k, int stride, int pad, int dil, void* stream) {
    return conv_fwd_impl(x, w, y, stats_partials, N, D, H, W, Cin, Cout, k, k, k, stride, pad, dil, nullptr, stream);
}

// Extended forward: output rows `ldy` elements apart (y may be a channel slice of a wider NDHWC tensor - the concatenation
// buffer of unet3d.py:77 is written in place, no torch.cat copy), per-channel epilogue (ConvEpi) and fp32 side output.
int mmad_conv3d_fwd_ex_bf16(const void* x, const void* w, void* y, int64_t ldy, float* stats_partials, const float* ep_scale,
                            const float* ep_shift, int ep_relu, float* out_f32, const float* f32_bias, int N, int D, int H, int W,
                            int Cin, int Cout, int k, int stride, int pad, int dil, int cin_tensor, void* stream) {
    MMAD_CHECK_ARG(ldy == 0 || (ldy >= Cout && ldy % 8 == 0), "conv3d_fwd_ex: ldy must be 0 (dense) or >= Cout and a multiple of 8");
    ConvEpi ep = {ep_scale, ep_shift, f32_bias, out_f32, (long long)Cout, ep_relu};
    ep.cin_tensor = cin_tensor;
    ConvOut ov = {};
    ov.Do = (D + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    ov.Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    ov.Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    const long long ld = ldy ? ldy : Cout;
    ov.sw = ld; ov.sh = ld * ov.Wo; ov.sd = ld * ov.Wo * ov.Ho; ov.sn = ld * ov.Wo * ov.Ho * ov.Do;
    ov.standard = 1;
    return conv_fwd_impl(x, w, y, stats_partials, N, D, H, W, Cin, Cout, k, k, k, stride, pad, dil, &ov, stream, &ep);
}

// 3x3x3 convolution (padding 1) to 64 channels + BatchNorm/ReLU epilogue + 1x1x1 head + crop, the activation never written:
// the tail of unet3d.py's forward in eval mode (s_block1.conv2 -> bn -> relu -> conv3 -> _crop_back) as ONE kernel.
// head_out fp32 (N, K, Dc, Hc, Wc); out_f32 (optional) = the raw convolution output + f32_bias, fp32 NDHWC (N, D, H, W, 64).
int mmad_conv3d_fwd_head_bf16(const void* x, const void* w, const float* ep_scale, const float* ep_shift, float* out_f32,
                              const float* f32_bias, const float* head_w, const float* head_b, float* head_out, int K, int N, int D, int H,
                              int W, int Dc, int Hc, int Wc, int Cin, void* stream) {
    MMAD_CHECK_ARG(head_w && head_b && head_out && Dc <= D && Hc <= H && Wc <= W && Dc > 0 && Hc > 0 && Wc > 0, "conv3d_fwd_head: bad argument");
    ConvEpi ep = {ep_scale, ep_shift, f32_bias, out_f32, 64, 1, head_w, head_b, head_out, K, Dc, Hc, Wc, 1};
    return conv_fwd_impl(x, w, nullptr, nullptr, N, D, H, W, Cin, 64, 3, 3, 3, 1, 1, 1, nullptr, stream, &ep);
}

// ConvTranspose3d(kernel 2, stride 2) forward (unet3d.py:68, :75): y[n][2v+p][co] = bias[co] + sum_ci x[n][v][ci] * w[ci][co][p]
// - eight 1x1x1 implicit GEMMs (one per output phase p) that write their outputs interleaved through strided tensor maps,
// rows `ldy` elements apart (so the result lands directly in the first channels of the concatenation buffer).
// w_phases: [8][Cout][Cin] bf16 from mmad_convtranspose3d_prep_weights.
int mmad_convtranspose3d_k2s2_fwd_bf16(const void* x, const void* w_phases, const float* bias, void* y, int64_t ldy, int N, int D, int H,
                                       int W, int Cin, int Cout, void* stream) {
    MMAD_CHECK_ARG(x && w_phases && y && N > 0 && D > 0 && H > 0 && W > 0, "convtranspose3d_k2s2: bad argument");
    MMAD_CHECK_ARG(ldy == 0 || (ldy >= Cout && ldy % 8 == 0), "convtranspose3d_k2s2: ldy must be 0 (dense) or >= Cout and a multiple of 8");
    const long long ld = ldy ? ldy : Cout;
    ConvEpi ep = {nullptr, bias, nullptr, nullptr, 0, 0};
    for (int p = 0; p < 8; ++p) {
        const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        ConvOut ov = {};
        ov.Do = D; ov.Ho = H; ov.Wo = W;
        ov.sw = 2 * ld; ov.sh = 2 * ld * (2 * W); ov.sd = 2 * ld * (2 * W) * (2 * H); ov.sn = ld * (2LL * W) * (2 * H) * (2 * D);
        char* out = static_cast<char*>(y) + (((size_t)pd * 2 * H + ph) * 2 * W + pw) * ld * 2;
        const int rc = conv_fwd_impl(x, static_cast<const char*>(w_phases) + (size_t)p * Cout * Cin * 2, out, nullptr, N, D, H, W, Cin, Cout,
                                     1, 1, 1, 1, 0, 1, &ov, stream, bias ? &ep : nullptr);
        if (rc) return rc;
    }
    return MMAD_OK;
}

// Data gradient of a 3x3x3, stride-2, padding-1 convolution WITHOUT zero insertion: dx positions of parity (pd, ph, pw) only
// see the taps of matching parity (1 tap on an even axis, 2 on an odd one), so dx is 8 interleaved stride-1 convolutions of
// dy with (1+pd) x (1+ph) x (1+pw) kernels - 27 taps over 1/8 of the voxels each instead of 27 taps over all of them.
// w_phases: mmad_conv3d_prep_weights_s2.  dx (N,D,H,W,Cdx) bf16, dy (N,Do,Ho,Wo,Cdy) bf16, Do = (D-1)/2+1 etc.
int mmad_conv3d_dgrad_s2_bf16(const void* dy, const void* w_phases, void* dx, int N, int D, int H, int W, int Cdx, int Cdy,
                              void* stream) {
    MMAD_CHECK_ARG(dy && w_phases && dx && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_dgrad_s2: bad argument");
    const int Do = (D - 1) / 2 + 1, Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    size_t woff = 0;
    for (int p = 0; p < 8; ++p) {
        const int pd = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        ConvOut ov = {};
        ov.Do = (D - pd + 1) / 2; ov.Ho = (H - ph + 1) / 2; ov.Wo = (W - pw + 1) / 2;
        ov.sw = 2ll * Cdx; ov.sh = 2ll * W * Cdx; ov.sd = 2ll * H * W * Cdx; ov.sn = (long long)D * H * W * Cdx;
        const int taps = (1 + pd) * (1 + ph) * (1 + pw);
        if (ov.Do > 0 && ov.Ho > 0 && ov.Wo > 0) {
            char* out = static_cast<char*>(dx) + (((size_t)pd * H + ph) * W + pw) * Cdx * 2;
            const int rc = conv_fwd_impl(dy, static_cast<const char*>(w_phases) + woff, out, nullptr, N, Do, Ho, Wo, Cdy, Cdx, 1 + pd,
                                         1 + ph, 1 + pw, 1, 0, 1, &ov, stream);
            if (rc) return rc;
        }
        woff += (size_t)taps * Cdx * Cdy * 2;
    }
    return MMAD_OK;
}


}  // extern "C"
