// Shared host/device helpers for the mmad_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>
#include <stdint.h>
#include <string>
#include <atomic>
#include <cstdio>
#include "../../include/mmad_b200.h"

namespace mmad {

// ---- host-side error plumbing ------------------------------------------------
std::string& last_error_ref();
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const std::string& msg) {
    last_error_ref() = msg;
    return code;
}

#define MMAD_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess)                                                       \
            return ::mmad::fail(MMAD_ECUDA, std::string(#call) + ": " +              \
                                                cudaGetErrorString(_e));             \
    } while (0)

#define MMAD_CHECK_ARG(cond, msg)                                                    \
    do {                                                                             \
        if (!(cond)) return ::mmad::fail(MMAD_EINVAL, std::string(msg));             \
    } while (0)

// "done once per device" flag (function attributes such as MaxDynamicSharedMemorySize are per device, not per process)
struct DevOnce {
    std::atomic<uint64_t> mask{0};
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        const uint64_t bit = 1ull << (d & 63);
        if (mask.load(std::memory_order_relaxed) & bit) return false;
        mask.fetch_or(bit, std::memory_order_relaxed);
        return true;
    }
};
// SM count of the current device (cached per device)
inline int sm_count() {
    static std::atomic<int> cache[64];
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return 148;
    int v = cache[d & 63].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
        cache[d & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- device-side PTX wrappers (mbarrier / bulk async copy) ---------------------

// Kernel launch with the programmatic-stream-serialization attribute (programmatic dependent launch): the kernel's CTAs may be
// scheduled, and run their prologue, while the previous kernel of the stream is still draining; every kernel launched this
// way executes griddepcontrol.wait (pdl_wait) before it touches global memory and griddepcontrol.launch_dependents
// (pdl_launch_dependents) at its top.  Measured on the ResNet3D-18 step: no gain (13.59 ms with, 13.47 ms without - the
// persistent kernels hold their SMs to the end, so the dependents' prologues cannot start earlier anyway), so the attribute
// is OFF unless MMAD_PDL=1; without it the two instructions are no-ops.
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("MMAD_PDL"); mode = e ? atoi(e) : 0; }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = mode ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Make barrier initialisation visible to the async (TMA) proxy.
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP).  src, dst and
// bytes must be multiples of 16.  Completion is signalled on `bar` as tx bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// Programmatic dependent launch: let the next kernel of the stream start launching / wait for the
// previous kernel of the stream to complete and flush (no-ops when launched without the attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// One lane of a fully converged warp (always the same one).  Warp-uniform loops whose asynchronous-unit instructions (TMA,
// tcgen05) are issued under `if (elect_one())` compile to plain predicated UTMALDG / UTCHMMA on uniform registers; the
// same instructions inside an `if (lane == 0)` region are wrapped by the compiler in an ELECT / R2UR / BRA.U.ANY
// serialisation loop (~10 extra instructions per issue), which is what limits a "single thread" to one copy per ~300 cycles.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

}  // namespace mmad
