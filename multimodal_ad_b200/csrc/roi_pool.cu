// Atlas ROI pooling for sm_100a: segmented mean / max / argmax of MRI volumes
// over an atlas label map.  Replaces /root/reference/image_features.py:80-82,
// 111-114 (one-hot mask product-sum) behind include/mmad_b200.h, part 1.
//
// Design (DESIGN.md "ROI pooling" has the long form):
//  * The atlas is the same for every volume of a batch, so the label map is
//    run-length encoded ONCE on the host into a per-tile "run programme"; the
//    kernel never touches per-voxel labels.
//  * A CTA owns 32 volumes x a contiguous range of voxel tiles.  Four producer
//    warps stage each tile (32 rows of TILE voxels, one row per volume) into
//    shared memory with 1-D bulk async copies (TMA engine) through an mbarrier
//    ring; the tile's run programme rides along on the same barrier.
//  * Consumer warps run with lane = volume.  Every lane sees the same labels,
//    so control flow is warp-uniform and lanes never collide.  A tile's records
//    (<= 8 voxels of one ROI each, sorted by ROI) are split evenly over the
//    consumer warps; a warp keeps the ROI it is on in registers and, when the ROI
//    changes, folds its partial into the CTA's shared accumulators with
//    shared-memory CAS atomics (rare: once per ROI change).
//  * Sums are carried in double per (ROI, volume); max and argmax travel as one
//    64-bit key (ordered value bits, inverted voxel index), so the first maximal
//    voxel wins whatever the order of the merges.  Partials leave the CTA once
//    per work item; a small second kernel reduces them in a fixed order.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

namespace mmad {

constexpr int kProducerWarps = 4;                 // NP: bulk-copy issue is serialised per warp (UBLKCP takes uniform regs)
constexpr int kRowsPerProducer = 32 / kProducerWarps;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kRecLen = 8;                        // a run record covers at most 8 consecutive voxels

// Per-tile programme: 4 header words (record count, 0, 0, 0), then the records sorted by (ROI, start).
constexpr int kHdrWords = 4;
__host__ __device__ constexpr int row_pitch(int tile) { return tile + 12; }                // floats; == 12 (mod 32): 4 | pitch, slack for 8-wide over-read
__host__ __device__ constexpr int prog_words(int tile) { return (kHdrWords + tile + 3) / 4 * 4; }
__host__ __device__ constexpr int stage_bytes(int tile) { return 32 * row_pitch(tile) * 4 + prog_words(tile) * 4; }

// max/argmax key: high word = float bits mapped to an order-preserving unsigned (with -0 folded into +0),
// low word = ~voxel index, so that of two equal values the LOWER index gives the larger key.  0 = "no voxel yet".
__host__ __device__ __forceinline__ unsigned long long roi_pack_key(float v, int idx) {
    v = v + 0.0f;
    unsigned int u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(v);
#else
    std::memcpy(&u, &v, 4);
#endif
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned int)idx);
}
__device__ __forceinline__ void roi_unpack_key(unsigned long long key, float& v, int& idx) {
    unsigned int u = (unsigned int)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    v = __uint_as_float(u);
    idx = (int)(0xffffffffu - (unsigned int)(key & 0xffffffffu));
}

struct RoiParams {
    const float* vols;
    long long n_vols;
    long long V;
    int R;
    int ns;        // pipeline stages
    int n_items;
    const uint32_t* prog;        // run programme, 16-byte units addressed by prog_off
    const int32_t* prog_off;     // [n_tiles + 1], in 16-byte units
    const int32_t* item_group;   // [n_items]
    const int32_t* item_t0;      // [n_items]
    const int32_t* item_t1;      // [n_items]
    const int32_t* item_slot_ptr;  // [n_items + 1]
    const uint8_t* slot_label;     // [n_slots] in item order: label 1..R of every partial the item writes
    const int32_t* slot_dst;       // [n_slots] in item order: destination slot (slots of one (group, ROI) are contiguous)
    double* slot_sum;              // [n_slots][32]
    unsigned long long* slot_key;  // [n_slots][32]
};

// Row of volume `lane` inside a stage.  Rows of the four volumes of a quad land
// 8 rows apart, so that with pitch == 4 (mod 32) and the per-volume alignment
// shift (distinct inside a quad when V is odd) the 32 lanes of a consumer warp
// read 32 different banks.
__device__ __forceinline__ int stage_row(int lane) { return (lane >> 2) + 8 * (lane & 3); }

// One record of n <= 8 consecutive voxels of one ROI, for the 32 volumes of the warp (lane = volume).
// `a` holds the 8 values at the record's start (values past n are never used).  Straight-line tree per n:
// pairwise fp32 sum -> one double add; first-occurrence argmax tree (ties keep the lower index).
template <int N>
__device__ __forceinline__ void roi_record(const float (&a)[kRecLen], int gidx, double& ds, float& mx, int& arg) {
    float t;
    if constexpr (N == 1) t = a[0];
    else if constexpr (N == 2) t = a[0] + a[1];
    else if constexpr (N == 3) t = (a[0] + a[1]) + a[2];
    else if constexpr (N == 4) t = (a[0] + a[1]) + (a[2] + a[3]);
    else if constexpr (N == 5) t = ((a[0] + a[1]) + (a[2] + a[3])) + a[4];
    else if constexpr (N == 6) t = ((a[0] + a[1]) + (a[2] + a[3])) + (a[4] + a[5]);
    else if constexpr (N == 7) t = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + a[6]);
    else t = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    ds += (double)t;
    // level 1
    float v[4];
    int ix[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (2 * k + 1 < N) {
            const bool g = a[2 * k + 1] > a[2 * k];
            v[k] = g ? a[2 * k + 1] : a[2 * k];
            ix[k] = g ? 2 * k + 1 : 2 * k;
        } else if (2 * k < N) {
            v[k] = a[2 * k];
            ix[k] = 2 * k;
        }
    }
    // level 2
    float w0 = v[0], w1 = 0.f;
    int j0 = ix[0], j1 = 0;
    if (N > 2) { const bool g = v[1] > v[0]; w0 = g ? v[1] : v[0]; j0 = g ? ix[1] : ix[0]; }
    if (N > 6) { const bool g = v[3] > v[2]; w1 = g ? v[3] : v[2]; j1 = g ? ix[3] : ix[2]; }
    else if (N > 4) { w1 = v[2]; j1 = ix[2]; }
    // level 3
    if (N > 4) { const bool g = w1 > w0; w0 = g ? w1 : w0; j0 = g ? j1 : j0; }
    if (w0 > mx || arg < 0) { mx = w0; arg = gidx + j0; }
}

template <int TILE, int NW>
__global__ void __launch_bounds__((NW + kProducerWarps) * 32, 1) roi_stream_kernel(const RoiParams p) {
    constexpr int P = row_pitch(TILE);
    constexpr int STAGE = stage_bytes(TILE);
    extern __shared__ __align__(16) unsigned char smem[];

    const int R = p.R;
    double* bins_sum = reinterpret_cast<double*>(smem);
    unsigned long long* bins_key = reinterpret_cast<unsigned long long*>(smem + (size_t)R * 256);
    unsigned char* stages = smem + (size_t)R * 512;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.ns * STAGE);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + p.ns);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ns = p.ns;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(full0 + 8 * s, kProducerWarps);
            mbar_init(empty0 + 8 * s, NW);
        }
        mbar_fence_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();               // everything above overlapped the previous kernel's tail; global memory is read below

    const int rho = stage_row(lane);
    uint32_t s = 0, ph = 0;   // ring position: stage and phase parity of the next tile

    if (warp >= NW) {
        // ===================== producer warps: one bulk copy per volume row, 8 rows per warp =====================
        const int pw = warp - NW;
        const int prow = pw * kRowsPerProducer + lane;          // volume (row) this lane copies; lanes >= 8 idle
        const int prho = stage_row(prow & 31);
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int g = p.item_group[item];
            const int t0 = p.item_t0[item], t1 = p.item_t1[item];
            const long long vol = (long long)g * 32 + prow;
            const bool act = lane < kRowsPerProducer && vol < p.n_vols;
            const float* vbase = p.vols + (act ? vol : 0) * p.V;
            const uint32_t sb = (uint32_t)((reinterpret_cast<uintptr_t>(vbase) >> 2) & 3);
            for (int tb = t0; tb < t1; tb += 32) {
                const int tt = tb + lane;
                const int o0 = (pw == 0 && tt < t1) ? p.prog_off[tt] : 0;
                const int o1 = (pw == 0 && tt < t1) ? p.prog_off[tt + 1] : 0;
                const int nt = min(32, t1 - tb);
                for (int k = 0; k < nt; ++k) {
                    const int t = tb + k;
                    const int po0 = __shfl_sync(0xffffffffu, o0, k);
                    const int po1 = __shfl_sync(0xffffffffu, o1, k);
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    const long long v0 = (long long)t * TILE;
                    const int L = (int)min((long long)TILE, p.V - v0);
                    const uint32_t bytes = act ? ((sb + (uint32_t)L + 3u) & ~3u) * 4u : 0u;
                    const uint32_t pbytes = (uint32_t)(po1 - po0) * 16u;      // 0 unless pw == 0
                    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes) + pbytes;
                    if (lane == 0) mbar_arrive_expect_tx(full0 + 8 * s, total);
                    __syncwarp();
                    const uint32_t sbase = smem_u32(stages + (size_t)s * STAGE);
                    if (act)
                        bulk_g2s(sbase + (uint32_t)prho * P * 4u,
                                 reinterpret_cast<const char*>(vbase + v0) - sb * 4u, bytes, full0 + 8 * s);
                    if (pw == 0 && lane == 0)
                        bulk_g2s(sbase + 32u * P * 4u, p.prog + (size_t)po0 * 4, pbytes, full0 + 8 * s);
                    if (++s == (uint32_t)ns) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ===================== consumer warps: lane = volume; a tile's records are split evenly over the warps =====================
    constexpr int NCT = NW * 32;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int g = p.item_group[item];
        const int t0 = p.item_t0[item], t1 = p.item_t1[item];
        const long long vol = (long long)g * 32 + lane;
        const bool act = vol < p.n_vols;
        const float* vbase = p.vols + (act ? vol : 0) * p.V;
        const uint32_t sb = (uint32_t)((reinterpret_cast<uintptr_t>(vbase) >> 2) & 3);

        for (int i = threadIdx.x; i < R * 32; i += NCT) {
            bins_sum[i] = 0.0;
            bins_key[i] = 0ull;
        }
        named_bar_sync(1, NCT);

        // partial of the ROI this warp is currently on (registers; survives tile boundaries)
        int cur = 0;
        double ds = 0.0;
        float mx = -INFINITY;
        int arg = -1;
        auto flush = [&]() {
            if (cur && arg >= 0) {
                const int b = (cur - 1) * 32 + lane;
                atomicAdd(&bins_sum[b], ds);
                atomicMax(&bins_key[b], roi_pack_key(mx, arg));
            }
        };

        for (int t = t0; t < t1; ++t) {
            mbar_wait(full0 + 8 * s, ph);
            const unsigned char* sbase = stages + (size_t)s * STAGE;
            const float* rowp = reinterpret_cast<const float*>(sbase) + rho * P + sb;
            const uint32_t* prog = reinterpret_cast<const uint32_t*>(sbase + 32 * P * 4);
            const uint32_t cnt = prog[0];
            const uint32_t r0 = cnt * (uint32_t)warp / NW, r1 = cnt * (uint32_t)(warp + 1) / NW;
            const int gbase = t * TILE;

            if (r0 < r1 && act) {
                uint32_t rec = prog[kHdrWords + r0];
                float a[kRecLen];
                {
                    const float* src = rowp + ((rec >> 12) & 0xfffu);
#pragma unroll
                    for (int k = 0; k < kRecLen; ++k) a[k] = src[k];
                }
                for (uint32_t i = r0; i < r1; ++i) {
                    const uint32_t crec = rec;
                    float c[kRecLen];
#pragma unroll
                    for (int k = 0; k < kRecLen; ++k) c[k] = a[k];
                    if (i + 1 < r1) {                               // prefetch the next record while this one is reduced
                        rec = prog[kHdrWords + i + 1];
                        const float* src = rowp + ((rec >> 12) & 0xfffu);
#pragma unroll
                        for (int k = 0; k < kRecLen; ++k) a[k] = src[k];
                    }
                    const int label = (int)(crec >> 24);
                    const int gidx = gbase + (int)((crec >> 12) & 0xfffu);
                    if (label != cur) {
                        flush();
                        cur = label; ds = 0.0; mx = -INFINITY; arg = -1;
                    }
                    switch (crec & 7u) {
                        case 0: roi_record<1>(c, gidx, ds, mx, arg); break;
                        case 1: roi_record<2>(c, gidx, ds, mx, arg); break;
                        case 2: roi_record<3>(c, gidx, ds, mx, arg); break;
                        case 3: roi_record<4>(c, gidx, ds, mx, arg); break;
                        case 4: roi_record<5>(c, gidx, ds, mx, arg); break;
                        case 5: roi_record<6>(c, gidx, ds, mx, arg); break;
                        case 6: roi_record<7>(c, gidx, ds, mx, arg); break;
                        default: roi_record<8>(c, gidx, ds, mx, arg); break;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
            if (++s == (uint32_t)ns) { s = 0; ph ^= 1; }
        }
        flush();

        named_bar_sync(1, NCT);
        const int sp0 = p.item_slot_ptr[item], sp1 = p.item_slot_ptr[item + 1];
        for (int j = sp0 + warp; j < sp1; j += NW) {
            const int b = ((int)p.slot_label[j] - 1) * 32 + lane;
            const size_t o = (size_t)p.slot_dst[j] * 32 + lane;
            p.slot_sum[o] = bins_sum[b];
            p.slot_key[o] = bins_key[b];
        }
        named_bar_sync(1, NCT);
    }
}

// Second pass: one block of four warps per (volume group, ROI), lane = volume.  The ROI's partial slots are contiguous and
// in ascending tile order; warp q adds slots q, q+4, ... in that order and the four warp sums are combined in warp order (a
// fixed association: results are reproducible run to run), keys are max-reduced (order independent).  Four warps instead of
// one shorten the dependent load chain of this latency-bound pass.
__global__ void __launch_bounds__(128) roi_finalize_kernel(const double* __restrict__ slot_sum,
                                                           const unsigned long long* __restrict__ slot_key,
                                                           const int32_t* __restrict__ fin_ptr,
                                                           const int32_t* __restrict__ counts, int n_groups, int R,
                                                           long long n_vols, float* __restrict__ mean,
                                                           float* __restrict__ mx_out, int32_t* __restrict__ arg_out) {
    __shared__ double sh_s[4][32];
    __shared__ unsigned long long sh_k[4][32];
    const int w = blockIdx.x;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int g = w / R, r = w - g * R;
    double s = 0.0;
    unsigned long long key = 0ull;
    const int k0 = fin_ptr[w], k1 = fin_ptr[w + 1];
    int k = k0 + q;
    for (; k + 12 < k1; k += 16) {
        double sv[4];
        unsigned long long kv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            sv[u] = slot_sum[(size_t)(k + 4 * u) * 32 + lane];
            kv[u] = slot_key[(size_t)(k + 4 * u) * 32 + lane];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s += sv[u];
            key = kv[u] > key ? kv[u] : key;
        }
    }
    for (; k < k1; k += 4) {
        s += slot_sum[(size_t)k * 32 + lane];
        const unsigned long long kk = slot_key[(size_t)k * 32 + lane];
        key = kk > key ? kk : key;
    }
    sh_s[q][lane] = s;
    sh_k[q][lane] = key;
    __syncthreads();
    if (q != 0) return;
    s = ((sh_s[0][lane] + sh_s[1][lane]) + sh_s[2][lane]) + sh_s[3][lane];
#pragma unroll
    for (int u = 1; u < 4; ++u) key = sh_k[u][lane] > key ? sh_k[u][lane] : key;
    const long long vol = (long long)g * 32 + lane;
    if (vol >= n_vols) return;
    const int cnt = counts[r];
    const float den = fmaxf((float)cnt, 1e-6f);
    float mx = 0.0f;
    int arg = -1;
    if (cnt && key) roi_unpack_key(key, mx, arg);
    const size_t o = (size_t)vol * R + r;
    if (mean) mean[o] = (float)s / den;
    if (mx_out) mx_out[o] = mx;
    if (arg_out) arg_out[o] = arg;
}

// Per-ROI voxel counts of the uint8 label map (plan creation, once per atlas).
__global__ void __launch_bounds__(256) roi_count_kernel(const uint8_t* __restrict__ labels, long long V,
                                                        int32_t* __restrict__ counts, int R) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(&hist[labels[i]], 1);
    __syncthreads();
    const int l = threadIdx.x;
    if (l >= 1 && l <= R && hist[l]) atomicAdd(&counts[l - 1], hist[l]);
}

// d mean / d vols  (image_features.py:111-114 under autograd).
__global__ void __launch_bounds__(256) roi_mean_bwd_kernel(const float* __restrict__ gmean,
                                                           const uint8_t* __restrict__ labels,
                                                           const int32_t* __restrict__ counts, long long V, int R,
                                                           long long n_vols, float* __restrict__ gvol) {
    const long long total = n_vols * V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / V;
        const long long v = i - n * V;
        const int l = labels[v];
        float o = 0.0f;
        if (l) o = __ldg(&gmean[n * R + (l - 1)]) / fmaxf((float)__ldg(&counts[l - 1]), 1e-6f);
        gvol[i] = o;
    }
}

// ------------------------------------------------------------------------------
// Channels-last variant: pools a (N, Dp, Hp, Wp, 64) fp32 NDHWC feature map - the layout the convolution epilogue writes
// the tensor image_features.py:58-60 hooks - over the atlas (D, H, W) <= (Dp, Hp, Wp) (the crop of :104 folded in), mean only
// (what image_features.py:111-114 computes).  A warp walks one (d, h) line: lane = channel pair, the line's labels sit in
// registers (one ballot per 32 voxels), and ONLY the rows of labelled voxels are read (256 contiguous bytes per voxel);
// consecutive voxels of the same ROI are summed in registers and folded into the block's shared double accumulators
// [R][64] when the label changes.  Blocks add their accumulators to acc[n][R][64] (double atomics), roi_cl_finalize_kernel
// divides by the clamped count in fp32 exactly as image_features.py:113-114 does.
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) roi_cl_pool_kernel(const float2* __restrict__ feats, const uint8_t* __restrict__ labels,
                                                          double* __restrict__ acc, int Dp, int Hp, int Wp, int D, int H, int W, int R) {
    extern __shared__ double sacc[];                     // [R][64]
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < R * 64; i += 256) sacc[i] = 0.0;
    __syncthreads();
    int cur = 0;
    float2 a = make_float2(0.f, 0.f);
    auto flush = [&]() {
        if (cur) {
            atomicAdd(&sacc[(cur - 1) * 64 + 2 * lane], (double)a.x);
            atomicAdd(&sacc[(cur - 1) * 64 + 2 * lane + 1], (double)a.y);
        }
        a = make_float2(0.f, 0.f);
    };
    for (int row = blockIdx.x * 8 + warp; row < D * H; row += gridDim.x * 8) {
        const int d = row / H, h = row - d * H;
        const uint8_t* lrow = labels + (long long)row * W;
        const float2* frow = feats + ((((long long)n * Dp + d) * Hp + h) * Wp) * 32;
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int lab = (w0 + lane < W) ? (int)lrow[w0 + lane] : 0;
            unsigned m = __ballot_sync(0xffffffffu, lab != 0);
            while (m) {
                int ix[4], cnt = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ix[q] = 0;
                    if (m) { ix[q] = __ffs(m) - 1; m &= m - 1; ++cnt; }
                }
                float2 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[q] = q < cnt ? __ldg(frow + (long long)(w0 + ix[q]) * 32 + lane) : make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int l = __shfl_sync(0xffffffffu, lab, ix[q]);
                    if (q < cnt) {
                        if (l != cur) { flush(); cur = l; }
                        a.x += v[q].x; a.y += v[q].y;
                    }
                }
            }
        }
    }
    flush();
    __syncthreads();
    double* dst = acc + (long long)n * R * 64;
    for (int i = threadIdx.x; i < R * 64; i += 256)
        if (sacc[i] != 0.0) atomicAdd(&dst[i], sacc[i]);
}
__global__ void __launch_bounds__(256) roi_cl_finalize_kernel(const double* __restrict__ acc, const int32_t* __restrict__ counts,
                                                              float* __restrict__ mean, long long total, int R) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)((i >> 6) % R);
    mean[i] = (float)acc[i] / fmaxf((float)__ldg(&counts[r]), 1e-6f);
}

// ------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------
struct Binding {
    long long n_vols = 0;
    int n_groups = 0, n_items = 0, grid = 0, n_slots = 0;
    std::vector<int32_t> h_item_group, h_item_t0, h_item_t1, h_item_slot_ptr, h_slot_dst, h_fin_ptr;
    std::vector<uint8_t> h_slot_label;
    int32_t *d_item_group = nullptr, *d_item_t0 = nullptr, *d_item_t1 = nullptr, *d_item_slot_ptr = nullptr;
    int32_t *d_slot_dst = nullptr, *d_fin_ptr = nullptr;
    uint8_t* d_slot_label = nullptr;
    double* d_slot_sum = nullptr;
    unsigned long long* d_slot_key = nullptr;
    void release() {
        cudaFree(d_item_group); cudaFree(d_item_t0); cudaFree(d_item_t1); cudaFree(d_item_slot_ptr);
        cudaFree(d_slot_dst); cudaFree(d_fin_ptr); cudaFree(d_slot_label);
        cudaFree(d_slot_sum); cudaFree(d_slot_key);
    }
};

}  // namespace mmad

struct mmad_roi_plan {
    long long V = 0;
    int R = 0, tile = 256, nw = 8, n_tiles = 0, ns = 0, sms = 148, device = 0;
    size_t smem_bytes = 0;
    std::vector<uint32_t> h_prog;         // words
    std::vector<int32_t> h_prog_off;      // 16-byte units, n_tiles + 1
    std::vector<uint64_t> h_tile_mask;    // n_tiles x 4 (bit l set: label l has voxels in the tile)
    std::vector<int32_t> h_counts;
    uint32_t* d_prog = nullptr;
    int32_t* d_prog_off = nullptr;
    uint8_t* d_labels = nullptr;
    int32_t* d_counts = nullptr;
    std::map<long long, mmad::Binding*> bindings;
    // host-buffer pipeline (mmad_roi_pool_host_f32)
    float* d_stage[2] = {nullptr, nullptr};
    long long stage_vols = 0;
    float* d_omean = nullptr; float* d_omax = nullptr; int32_t* d_oarg = nullptr;
    long long out_cap = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // channels-last pooling (mmad_roi_pool_ndhwc_f32): double accumulators [n_vols][R][64]
    double* d_cl_acc = nullptr;
    long long cl_cap = 0;
};

namespace mmad {

// Host-only: run-length encode the label map into the per-tile programme (records of <= 8 voxels,
// sorted by (ROI, start) inside a tile).
static int build_programme(const int32_t* labels, long long V, int R, int tile, std::vector<uint32_t>& prog,
                           std::vector<int32_t>& prog_off, std::vector<uint64_t>& tile_mask, int& n_tiles) {
    n_tiles = (int)((V + tile - 1) / tile);
    prog.clear();
    prog_off.assign((size_t)n_tiles + 1, 0);
    tile_mask.assign((size_t)n_tiles * 4, 0ull);
    std::vector<uint32_t> recs;
    for (int t = 0; t < n_tiles; ++t) {
        recs.clear();
        const long long v0 = (long long)t * tile;
        const int L = (int)std::min<long long>(tile, V - v0);
        int q = 0;
        while (q < L) {
            const int32_t l = labels[v0 + q];
            if (l < 0 || l > R) return -1;
            int e = q + 1;
            while (e < L && labels[v0 + e] == l) ++e;
            if (l) {
                for (int a = q; a < e; a += kRecLen)
                    recs.push_back(((uint32_t)l << 24) | ((uint32_t)a << 12) | (uint32_t)(std::min(kRecLen, e - a) - 1));
                tile_mask[(size_t)t * 4 + (l >> 6)] |= 1ull << (l & 63);
            }
            q = e;
        }
        std::sort(recs.begin(), recs.end());
        prog_off[t] = (int32_t)(prog.size() / 4);
        prog.push_back((uint32_t)recs.size());
        for (int k = 1; k < kHdrWords; ++k) prog.push_back(0u);
        prog.insert(prog.end(), recs.begin(), recs.end());
        while (prog.size() % 4) prog.push_back(0u);
    }
    prog_off[n_tiles] = (int32_t)(prog.size() / 4);
    return 0;
}

// Host-only: split the (volume group x tile) space into work items and lay out the partial slots.
// Slots of one (group, ROI) are contiguous and ordered by tile range, so the second pass reads them linearly.
static void build_binding_host(const mmad_roi_plan& pl, long long n_vols, Binding& b) {
    b.n_vols = n_vols;
    b.n_groups = (int)((n_vols + 31) / 32);
    int pieces = 1;
    if (b.n_groups <= pl.sms) pieces = std::max(1, std::min(pl.sms / b.n_groups, pl.n_tiles));
    b.n_items = b.n_groups * pieces;
    b.grid = std::min(b.n_items, pl.sms);
    b.h_item_group.resize(b.n_items); b.h_item_t0.resize(b.n_items); b.h_item_t1.resize(b.n_items);
    b.h_item_slot_ptr.assign((size_t)b.n_items + 1, 0);
    b.h_slot_label.clear();
    b.h_slot_dst.clear();
    // pass 1: which ROIs each piece touches; count slots per (group, ROI)
    std::vector<std::vector<uint8_t>> piece_labels((size_t)pieces);
    std::vector<int32_t> per_roi((size_t)pl.R, 0);
    for (int piece = 0; piece < pieces; ++piece) {
        const int t0 = (int)((long long)pl.n_tiles * piece / pieces);
        const int t1 = (int)((long long)pl.n_tiles * (piece + 1) / pieces);
        uint64_t m[4] = {0, 0, 0, 0};
        for (int t = t0; t < t1; ++t)
            for (int k = 0; k < 4; ++k) m[k] |= pl.h_tile_mask[(size_t)t * 4 + k];
        for (int l = 1; l <= pl.R; ++l)
            if (m[l >> 6] >> (l & 63) & 1ull) { piece_labels[piece].push_back((uint8_t)l); per_roi[l - 1]++; }
    }
    b.h_fin_ptr.assign((size_t)b.n_groups * pl.R + 1, 0);
    for (int g = 0; g < b.n_groups; ++g)
        for (int r = 0; r < pl.R; ++r)
            b.h_fin_ptr[(size_t)g * pl.R + r + 1] = b.h_fin_ptr[(size_t)g * pl.R + r] + per_roi[r];
    b.n_slots = b.h_fin_ptr.back();
    // pass 2: items (item = piece * n_groups + g: neighbouring CTAs stream the same tile range of different groups)
    std::vector<int32_t> next(b.h_fin_ptr.begin(), b.h_fin_ptr.end() - 1);
    for (int piece = 0; piece < pieces; ++piece) {
        const int t0 = (int)((long long)pl.n_tiles * piece / pieces);
        const int t1 = (int)((long long)pl.n_tiles * (piece + 1) / pieces);
        for (int g = 0; g < b.n_groups; ++g) {
            const int item = piece * b.n_groups + g;
            b.h_item_group[item] = g; b.h_item_t0[item] = t0; b.h_item_t1[item] = t1;
            b.h_item_slot_ptr[item] = (int32_t)b.h_slot_label.size();
            for (uint8_t l : piece_labels[piece]) {
                b.h_slot_label.push_back(l);
                b.h_slot_dst.push_back(next[(size_t)g * pl.R + (l - 1)]++);      // ascending piece => ascending tile
            }
        }
    }
    b.h_item_slot_ptr[b.n_items] = (int32_t)b.h_slot_label.size();
}

template <typename T>
static cudaError_t upload(T** dptr, const std::vector<T>& h) {
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)dptr, bytes);
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static int get_binding(mmad_roi_plan* pl, long long n_vols, Binding** out) {
    auto itb = pl->bindings.find(n_vols);
    if (itb != pl->bindings.end()) { *out = itb->second; return MMAD_OK; }
    Binding* b = new Binding();
    build_binding_host(*pl, n_vols, *b);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = upload(&b->d_item_group, b->h_item_group);
    if (e == cudaSuccess) e = upload(&b->d_item_t0, b->h_item_t0);
    if (e == cudaSuccess) e = upload(&b->d_item_t1, b->h_item_t1);
    if (e == cudaSuccess) e = upload(&b->d_item_slot_ptr, b->h_item_slot_ptr);
    if (e == cudaSuccess) e = upload(&b->d_slot_label, b->h_slot_label);
    if (e == cudaSuccess) e = upload(&b->d_slot_dst, b->h_slot_dst);
    if (e == cudaSuccess) e = upload(&b->d_fin_ptr, b->h_fin_ptr);
    const size_t ns = std::max(b->n_slots, 1);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_slot_sum, ns * 32 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_slot_key, ns * 32 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        b->release();
        delete b;
        return fail(MMAD_ECUDA, std::string("roi binding upload: ") + cudaGetErrorString(e));
    }
    if (pl->bindings.size() >= 16) {                       // bounded cache: drop the oldest batch size (cudaFree waits for its kernels)
        auto old = pl->bindings.begin();
        old->second->release();
        delete old->second;
        pl->bindings.erase(old);
    }
    pl->bindings[n_vols] = b;
    *out = b;
    return MMAD_OK;
}

static int launch_pool(mmad_roi_plan* pl, const float* vols_dev, long long n_vols, float* mean_dev, float* max_dev,
                       int32_t* argmax_dev, cudaStream_t st) {
    Binding* b = nullptr;
    int rc = get_binding(pl, n_vols, &b);
    if (rc) return rc;
    RoiParams p;
    p.vols = vols_dev; p.n_vols = n_vols; p.V = pl->V; p.R = pl->R; p.ns = pl->ns; p.n_items = b->n_items;
    p.prog = pl->d_prog; p.prog_off = pl->d_prog_off;
    p.item_group = b->d_item_group; p.item_t0 = b->d_item_t0; p.item_t1 = b->d_item_t1;
    p.item_slot_ptr = b->d_item_slot_ptr; p.slot_label = b->d_slot_label;
    p.slot_dst = b->d_slot_dst; p.slot_sum = b->d_slot_sum; p.slot_key = b->d_slot_key;
    // Both kernels carry the programmatic-stream-serialization attribute: their prologues overlap the
    // previous kernel's tail, and each waits (griddepcontrol.wait) before it touches global memory.
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(b->grid);
    cfg.blockDim = dim3((pl->nw + kProducerWarps) * 32);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define MMAD_ROI_LAUNCH(T, W) MMAD_CUDA(cudaLaunchKernelEx(&cfg, roi_stream_kernel<T, W>, p))
    const int key = pl->tile * 100 + pl->nw;
    switch (key) {
        case 12808: MMAD_ROI_LAUNCH(128, 8); break;
        case 25608: MMAD_ROI_LAUNCH(256, 8); break;
        case 51208: MMAD_ROI_LAUNCH(512, 8); break;
        case 12816: MMAD_ROI_LAUNCH(128, 16); break;
        case 25616: MMAD_ROI_LAUNCH(256, 16); break;
        case 51216: MMAD_ROI_LAUNCH(512, 16); break;
        default: return fail(MMAD_EUNSUPPORTED, "roi tile must be 128, 256 or 512 and consumer warps 8 or 16");
    }
#undef MMAD_ROI_LAUNCH
    if (!mean_dev && !max_dev && !argmax_dev) {      // no output requested: partials only (bench.py times this kernel alone)
        count_launch(1);
        return MMAD_OK;
    }
    const int warps = b->n_groups * pl->R;
    cfg.gridDim = dim3(warps);                            // one block (four warps) per (group, ROI)
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    MMAD_CUDA(cudaLaunchKernelEx(&cfg, roi_finalize_kernel, (const double*)b->d_slot_sum,
                                 (const unsigned long long*)b->d_slot_key, (const int32_t*)b->d_fin_ptr,
                                 (const int32_t*)pl->d_counts, b->n_groups, pl->R, (long long)n_vols, mean_dev, max_dev,
                                 argmax_dev));
    count_launch(2);
    return MMAD_OK;
}

}  // namespace mmad

using namespace mmad;

extern "C" {

// Extended constructor used by bench.py / tests for tuning: tile in {128,256,512},
// stages 0 = as many as fit (max 4).
int mmad_roi_plan_create_ex(const int32_t* labels_host, int64_t n_voxels, int32_t n_rois, int32_t tile,
                            int32_t stages, int32_t consumer_warps, int32_t host_only, mmad_roi_plan** plan_out) {
    MMAD_CHECK_ARG(labels_host && plan_out, "roi_plan_create: null pointer");
    MMAD_CHECK_ARG(n_voxels > 0 && n_voxels < (1ll << 31) - 4096, "roi_plan_create: n_voxels must be in (0, 2^31)");
    MMAD_CHECK_ARG(n_rois >= 1 && n_rois <= 255, "roi_plan_create: n_rois must be 1..255");
    MMAD_CHECK_ARG(tile == 128 || tile == 256 || tile == 512, "roi_plan_create: tile must be 128, 256 or 512");
    if (consumer_warps == 0) consumer_warps = 16;
    MMAD_CHECK_ARG(consumer_warps == 8 || consumer_warps == 16, "roi_plan_create: consumer_warps must be 8 or 16");
    mmad_roi_plan* pl = new mmad_roi_plan();
    pl->V = n_voxels; pl->R = n_rois; pl->tile = tile; pl->nw = consumer_warps;
    if (build_programme(labels_host, n_voxels, n_rois, tile, pl->h_prog, pl->h_prog_off, pl->h_tile_mask, pl->n_tiles)) {
        delete pl;
        return fail(MMAD_EINVAL, "roi_plan_create: label outside [0, n_rois]");
    }
    const size_t bins = (size_t)n_rois * 512;
    int ns = (int)((kMaxSmem - bins - 64) / stage_bytes(tile));
    ns = std::min(ns, 4);
    if (stages > 0) ns = std::min(ns, (int)stages);
    if (ns < 2) { delete pl; return fail(MMAD_EUNSUPPORTED, "roi_plan_create: shared memory too small for this tile / n_rois"); }
    pl->ns = ns;
    pl->smem_bytes = bins + (size_t)ns * stage_bytes(tile) + 64;
    if (host_only) { *plan_out = pl; return MMAD_OK; }

    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { delete pl; return fail(MMAD_ECUDA, std::string("roi_plan_create: no CUDA device: ") + cudaGetErrorString(e)); }
    pl->device = dev; pl->sms = sms;
    std::vector<uint8_t> lab8((size_t)n_voxels);
    for (long long i = 0; i < n_voxels; ++i) lab8[i] = (uint8_t)labels_host[i];
    if (e == cudaSuccess) e = upload(&pl->d_prog, pl->h_prog);
    if (e == cudaSuccess) e = upload(&pl->d_prog_off, pl->h_prog_off);
    if (e == cudaSuccess) e = upload(&pl->d_labels, lab8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_counts, sizeof(int32_t) * n_rois);
    if (e == cudaSuccess) e = cudaMemset(pl->d_counts, 0, sizeof(int32_t) * n_rois);
    if (e == cudaSuccess) {
        roi_count_kernel<<<std::max(1, std::min(4 * sms, (int)((n_voxels + 255) / 256))), 256>>>(pl->d_labels, n_voxels, pl->d_counts, n_rois);
        count_launch();
        e = cudaGetLastError();
    }
    pl->h_counts.resize(n_rois);
    if (e == cudaSuccess) e = cudaMemcpy(pl->h_counts.data(), pl->d_counts, sizeof(int32_t) * n_rois, cudaMemcpyDeviceToHost);
#define MMAD_ROI_ATTR(T, W) if (e == cudaSuccess) e = cudaFuncSetAttribute(roi_stream_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)
    MMAD_ROI_ATTR(128, 8); MMAD_ROI_ATTR(256, 8); MMAD_ROI_ATTR(512, 8);
    MMAD_ROI_ATTR(128, 16); MMAD_ROI_ATTR(256, 16); MMAD_ROI_ATTR(512, 16);
#undef MMAD_ROI_ATTR
    if (e != cudaSuccess) {
        std::string msg = std::string("roi_plan_create: ") + cudaGetErrorString(e);
        mmad_roi_plan_destroy(pl);
        return fail(MMAD_ECUDA, msg);
    }
    *plan_out = pl;
    return MMAD_OK;
}

int mmad_roi_plan_create(const int32_t* labels_host, int64_t n_voxels, int32_t n_rois, mmad_roi_plan** plan_out) {
    return mmad_roi_plan_create_ex(labels_host, n_voxels, n_rois, 256, 0, 0, 0, plan_out);
}

int mmad_roi_plan_destroy(mmad_roi_plan* pl) {
    if (!pl) return MMAD_OK;
    for (auto& kv : pl->bindings) { kv.second->release(); delete kv.second; }
    cudaFree(pl->d_prog); cudaFree(pl->d_prog_off); cudaFree(pl->d_labels); cudaFree(pl->d_counts);
    cudaFree(pl->d_stage[0]); cudaFree(pl->d_stage[1]);
    cudaFree(pl->d_omean); cudaFree(pl->d_omax); cudaFree(pl->d_oarg); cudaFree(pl->d_cl_acc);
    for (int i = 0; i < 2; ++i) {
        if (pl->ev_copied[i]) cudaEventDestroy(pl->ev_copied[i]);
        if (pl->ev_done[i]) cudaEventDestroy(pl->ev_done[i]);
    }
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->s_comp) cudaStreamDestroy(pl->s_comp);
    delete pl;
    return MMAD_OK;
}

int mmad_roi_plan_counts(const mmad_roi_plan* pl, int32_t* counts_host) {
    MMAD_CHECK_ARG(pl && counts_host, "roi_plan_counts: null pointer");
    MMAD_CHECK_ARG(!pl->h_counts.empty(), "roi_plan_counts: host-only plan has no GPU counts");
    std::memcpy(counts_host, pl->h_counts.data(), sizeof(int32_t) * pl->R);
    return MMAD_OK;
}

int mmad_roi_plan_counts_dev(const mmad_roi_plan* pl, const int32_t** counts_dev) {
    MMAD_CHECK_ARG(pl && counts_dev && pl->d_counts, "roi_plan_counts_dev: null pointer / host-only plan");
    *counts_dev = pl->d_counts;
    return MMAD_OK;
}

int mmad_roi_pool_f32(mmad_roi_plan* pl, const float* vols_dev, int64_t n_vols, float* mean_dev, float* max_dev,
                      int32_t* argmax_dev, void* stream) {
    MMAD_CHECK_ARG(pl && pl->d_prog, "roi_pool: null or host-only plan");
    MMAD_CHECK_ARG(n_vols >= 0, "roi_pool: n_vols < 0");
    if (n_vols == 0) return MMAD_OK;
    MMAD_CHECK_ARG(vols_dev, "roi_pool: null volumes");
    MMAD_CHECK_ARG((reinterpret_cast<uintptr_t>(vols_dev) & 3) == 0, "roi_pool: volumes must be 4-byte aligned");
    return launch_pool(pl, vols_dev, n_vols, mean_dev, max_dev, argmax_dev, (cudaStream_t)stream);
}

int mmad_roi_pool_mean_backward_f32(mmad_roi_plan* pl, const float* grad_mean_dev, int64_t n_vols,
                                    float* grad_vols_dev, void* stream) {
    MMAD_CHECK_ARG(pl && pl->d_labels, "roi_pool_mean_backward: null or host-only plan");
    MMAD_CHECK_ARG(n_vols >= 0, "roi_pool_mean_backward: n_vols < 0");
    if (n_vols == 0) return MMAD_OK;
    MMAD_CHECK_ARG(grad_mean_dev && grad_vols_dev, "roi_pool_mean_backward: null pointer");
    const long long total = (long long)n_vols * pl->V;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)pl->sms * 16);
    roi_mean_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_mean_dev, pl->d_labels, pl->d_counts, pl->V, pl->R,
                                                                n_vols, grad_vols_dev);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    return MMAD_OK;
}

// Channels-last pooling of a convolution output that never left the GPU (see roi_cl_pool_kernel).
int mmad_roi_pool_ndhwc_f32(mmad_roi_plan* pl, const float* feats_dev, int64_t n_vols, int Dp, int Hp, int Wp, int D, int H, int W,
                            int C, float* mean_dev, void* stream) {
    MMAD_CHECK_ARG(pl && pl->d_labels, "roi_pool_ndhwc: null or host-only plan");
    MMAD_CHECK_ARG(n_vols >= 0, "roi_pool_ndhwc: n_vols < 0");
    if (n_vols == 0) return MMAD_OK;
    MMAD_CHECK_ARG(feats_dev && mean_dev, "roi_pool_ndhwc: null pointer");
    MMAD_CHECK_ARG(C == 64, "roi_pool_ndhwc: the channels-last kernel pools 64-channel feature maps (s_block1.conv2)");
    MMAD_CHECK_ARG((long long)D * H * W == pl->V && D <= Dp && H <= Hp && W <= Wp && D > 0 && H > 0 && W > 0,
                   "roi_pool_ndhwc: the atlas grid must match the plan and lie inside the feature grid");
    MMAD_CHECK_ARG((reinterpret_cast<uintptr_t>(feats_dev) & 7) == 0, "roi_pool_ndhwc: features must be 8-byte aligned");
    const size_t smem = (size_t)pl->R * 64 * sizeof(double);
    MMAD_CHECK_ARG(smem <= (size_t)kMaxSmem, "roi_pool_ndhwc: too many ROIs for the shared accumulators");
    cudaStream_t st = (cudaStream_t)stream;
    const long long need = (long long)n_vols * pl->R * 64;
    if (need > pl->cl_cap) {
        cudaFree(pl->d_cl_acc);
        pl->d_cl_acc = nullptr; pl->cl_cap = 0;
        MMAD_CUDA(cudaMalloc((void**)&pl->d_cl_acc, (size_t)need * sizeof(double)));
        pl->cl_cap = need;
    }
    MMAD_CUDA(cudaMemsetAsync(pl->d_cl_acc, 0, (size_t)need * sizeof(double), st));
    static DevOnce attr_done;
    if (attr_done.need()) MMAD_CUDA(cudaFuncSetAttribute(roi_cl_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    const int per_sm = std::max(1, std::min(2, (int)((size_t)kMaxSmem / (smem + 1024))));
    const long long lines = ((long long)D * H + 7) / 8;
    const int gx = (int)std::max<long long>(1, std::min<long long>(lines, ((long long)pl->sms * per_sm + n_vols - 1) / n_vols));
    roi_cl_pool_kernel<<<dim3(gx, (unsigned)n_vols), 256, smem, st>>>(reinterpret_cast<const float2*>(feats_dev), pl->d_labels, pl->d_cl_acc,
                                                                   Dp, Hp, Wp, D, H, W, pl->R);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    roi_cl_finalize_kernel<<<(unsigned)((need + 255) / 256), 256, 0, st>>>(pl->d_cl_acc, pl->d_counts, mean_dev, need, pl->R);
    MMAD_CUDA(cudaGetLastError());
    count_launch();
    return MMAD_OK;
}

int mmad_roi_pool_host_f32(mmad_roi_plan* pl, const float* vols_host, int64_t n_vols, float* mean_host,
                           float* max_host, int32_t* argmax_host) {
    MMAD_CHECK_ARG(pl && pl->d_prog, "roi_pool_host: null or host-only plan");
    MMAD_CHECK_ARG(n_vols >= 0, "roi_pool_host: n_vols < 0");
    if (n_vols == 0) return MMAD_OK;
    MMAD_CHECK_ARG(vols_host, "roi_pool_host: null volumes");
    const long long chunk = 8;   // volumes per H2D chunk: PCIe time >> kernel time, small chunks hide the kernel
    if (!pl->s_copy) {
        MMAD_CUDA(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
        MMAD_CUDA(cudaStreamCreateWithFlags(&pl->s_comp, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            MMAD_CUDA(cudaEventCreateWithFlags(&pl->ev_copied[i], cudaEventDisableTiming));
            MMAD_CUDA(cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming));
        }
    }
    if (pl->stage_vols < chunk) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(pl->d_stage[i]);
            pl->d_stage[i] = nullptr;
            MMAD_CUDA(cudaMalloc((void**)&pl->d_stage[i], sizeof(float) * chunk * pl->V));
        }
        pl->stage_vols = chunk;
    }
    if (pl->out_cap < n_vols) {
        cudaFree(pl->d_omean); cudaFree(pl->d_omax); cudaFree(pl->d_oarg);
        pl->d_omean = nullptr; pl->d_omax = nullptr; pl->d_oarg = nullptr;
        MMAD_CUDA(cudaMalloc((void**)&pl->d_omean, sizeof(float) * n_vols * pl->R));
        MMAD_CUDA(cudaMalloc((void**)&pl->d_omax, sizeof(float) * n_vols * pl->R));
        MMAD_CUDA(cudaMalloc((void**)&pl->d_oarg, sizeof(int32_t) * n_vols * pl->R));
        pl->out_cap = n_vols;
    }
    int k = 0;
    for (long long v0 = 0; v0 < n_vols; v0 += chunk, ++k) {
        const int buf = k & 1;
        const long long n = std::min(chunk, (long long)n_vols - v0);
        if (k >= 2) MMAD_CUDA(cudaStreamWaitEvent(pl->s_copy, pl->ev_done[buf], 0));   // buffer free again
        MMAD_CUDA(cudaMemcpyAsync(pl->d_stage[buf], vols_host + v0 * pl->V, sizeof(float) * n * pl->V,
                                  cudaMemcpyHostToDevice, pl->s_copy));
        MMAD_CUDA(cudaEventRecord(pl->ev_copied[buf], pl->s_copy));
        MMAD_CUDA(cudaStreamWaitEvent(pl->s_comp, pl->ev_copied[buf], 0));
        int rc = launch_pool(pl, pl->d_stage[buf], n, pl->d_omean + v0 * pl->R, pl->d_omax + v0 * pl->R,
                             pl->d_oarg + v0 * pl->R, pl->s_comp);
        if (rc) return rc;
        MMAD_CUDA(cudaEventRecord(pl->ev_done[buf], pl->s_comp));
    }
    if (mean_host) MMAD_CUDA(cudaMemcpyAsync(mean_host, pl->d_omean, sizeof(float) * n_vols * pl->R, cudaMemcpyDeviceToHost, pl->s_comp));
    if (max_host) MMAD_CUDA(cudaMemcpyAsync(max_host, pl->d_omax, sizeof(float) * n_vols * pl->R, cudaMemcpyDeviceToHost, pl->s_comp));
    if (argmax_host) MMAD_CUDA(cudaMemcpyAsync(argmax_host, pl->d_oarg, sizeof(int32_t) * n_vols * pl->R, cudaMemcpyDeviceToHost, pl->s_comp));
    MMAD_CUDA(cudaStreamSynchronize(pl->s_comp));
    return MMAD_OK;
}

int64_t mmad_roi_pool_algorithmic_bytes(const mmad_roi_plan* pl, int64_t n_vols) {
    if (!pl || n_vols <= 0) return 0;
    // every voxel of every volume once, the run programme once, three outputs per (volume, ROI)
    return (int64_t)n_vols * pl->V * 4 + (int64_t)pl->h_prog.size() * 4 + (int64_t)n_vols * pl->R * 12;
}

// ---- host-only introspection for the CPU test-suite (no CUDA calls) ----------------------
// Copies the run programme of a plan: words (may be NULL to query sizes) and 16-byte-unit offsets.
int mmad_roi_plan_programme(const mmad_roi_plan* pl, uint32_t* words, int64_t* n_words, int32_t* offs,
                            int32_t* n_tiles, int32_t* stages, int64_t* smem_bytes, int32_t* consumer_warps) {
    MMAD_CHECK_ARG(pl, "roi_plan_programme: null plan");
    if (n_words) *n_words = (int64_t)pl->h_prog.size();
    if (n_tiles) *n_tiles = pl->n_tiles;
    if (stages) *stages = pl->ns;
    if (smem_bytes) *smem_bytes = (int64_t)pl->smem_bytes;
    if (consumer_warps) *consumer_warps = pl->nw;
    if (words) std::memcpy(words, pl->h_prog.data(), pl->h_prog.size() * 4);
    if (offs) std::memcpy(offs, pl->h_prog_off.data(), pl->h_prog_off.size() * 4);
    return MMAD_OK;
}

// Work-item / slot layout the plan would use for n_vols volumes on a GPU with `sms` SMs
// (host-only).  Arrays may be NULL to query sizes first.
int mmad_roi_plan_binding(mmad_roi_plan* pl, int64_t n_vols, int32_t sms, int32_t* n_items, int32_t* n_slots,
                          int32_t* grid, int32_t* item_group, int32_t* item_t0, int32_t* item_t1,
                          int32_t* item_slot_ptr, uint8_t* slot_label, int32_t* slot_dst, int32_t* fin_ptr) {
    MMAD_CHECK_ARG(pl && n_vols > 0 && sms > 0, "roi_plan_binding: bad argument");
    const int keep = pl->sms;
    pl->sms = sms;
    Binding b;
    build_binding_host(*pl, n_vols, b);
    pl->sms = keep;
    if (n_items) *n_items = b.n_items;
    if (n_slots) *n_slots = b.n_slots;
    if (grid) *grid = b.grid;
    if (item_group) std::memcpy(item_group, b.h_item_group.data(), b.h_item_group.size() * 4);
    if (item_t0) std::memcpy(item_t0, b.h_item_t0.data(), b.h_item_t0.size() * 4);
    if (item_t1) std::memcpy(item_t1, b.h_item_t1.data(), b.h_item_t1.size() * 4);
    if (item_slot_ptr) std::memcpy(item_slot_ptr, b.h_item_slot_ptr.data(), b.h_item_slot_ptr.size() * 4);
    if (slot_label) std::memcpy(slot_label, b.h_slot_label.data(), b.h_slot_label.size());
    if (slot_dst) std::memcpy(slot_dst, b.h_slot_dst.data(), b.h_slot_dst.size() * 4);
    if (fin_ptr) std::memcpy(fin_ptr, b.h_fin_ptr.data(), b.h_fin_ptr.size() * 4);
    return MMAD_OK;
}

}  // extern "C"
