// Developer probe (not product code): what saturates B200 HBM for a read-only stream?
//   mode 0: LDG.128 grid-stride, U loads in flight per thread
//   mode 1: 1-D bulk async copies (UBLKCP), NS stages x CS bytes per CTA, contiguous per CTA
//   mode 2: like 1 but each stage is 32 copies of CS/32 bytes from 32 streams (volume rows)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok;
    do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(b), "r"(ph) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int U>
__global__ void __launch_bounds__(512) ldg_kernel(const float4* __restrict__ in, size_t n4, float* out) {
    float acc = 0.f;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(in + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// mode 1/2
__global__ void __launch_bounds__(128) bulk_kernel(const char* __restrict__ in, size_t total, int ns, int cs, int streams, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = (uint64_t*)(smem + (size_t)ns * cs);
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) mbar_init(s32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // this CTA's share
    size_t per = total / gridDim.x / cs * cs;
    const char* base = in + per * blockIdx.x;
    size_t nchunks = per / cs;
    float acc = 0.f;
    if (threadIdx.x < 32) {
        int lane = threadIdx.x;
        size_t sub = cs / streams;             // bytes per stream copy
        size_t stream_span = per / streams;    // each stream covers a contiguous 1/streams of the CTA share
        auto issue = [&](size_t c, int s) {
            if (lane == 0) mbar_expect(s32(bars + s), cs);
            __syncwarp();
            if (streams == 1) { if (lane == 0) bulk(s32(smem + (size_t)s * cs), base + c * cs, cs, s32(bars + s)); }
            else if (lane < streams) bulk(s32(smem + (size_t)s * cs + lane * sub), base + lane * stream_span + c * sub, sub, s32(bars + s));
        };
        for (int s = 0; s < ns && s < nchunks; ++s) issue(s, s);
        for (size_t c = 0; c < nchunks; ++c) {
            int s = c % ns; uint32_t ph = (c / ns) & 1;
            mbar_wait(s32(bars + s), ph);
            acc += ((float*)(smem + (size_t)s * cs))[lane];
            __syncwarp();
            if (c + ns < nchunks) issue(c + ns, s);
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
    size_t bytes = (size_t)1 << 30;
    char* d; float* o;
    cudaMalloc(&d, bytes); cudaMalloc(&o, 4);
    cudaMemset(d, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto time = [&](auto f, const char* name) {
        for (int i = 0; i < 2; ++i) f();
        cudaEventRecord(a);
        const int K = 5;
        for (int i = 0; i < K; ++i) f();
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        cudaError_t e = cudaGetLastError();
        printf("%-44s %8.1f GB/s %s\n", name, bytes / (ms / K * 1e-3) / 1e9, e ? cudaGetErrorString(e) : "");
        fflush(stdout);
    };
    char name[128];
    for (int blocks : {148 * 2, 148 * 4}) {
        snprintf(name, 128, "ldg128 U=4 grid=%d x512", blocks); time([&] { ldg_kernel<4><<<blocks, 512>>>((const float4*)d, bytes / 16, o); }, name);
        snprintf(name, 128, "ldg128 U=8 grid=%d x512", blocks); time([&] { ldg_kernel<8><<<blocks, 512>>>((const float4*)d, bytes / 16, o); }, name);
    }
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int ctas_per_sm : {1, 2})
        for (int cs : {4096, 8192, 16384, 32768})
            for (int ns : {2, 4, 6, 12})
                for (int streams : {1, 8, 32}) {
                    size_t smem = (size_t)ns * cs + 128;
                    if (smem * ctas_per_sm > 220 * 1024) continue;
                    snprintf(name, 128, "bulk cs=%d ns=%d streams=%d cta/sm=%d inflight=%zuKB", cs, ns, streams, ctas_per_sm, (size_t)ns * cs * ctas_per_sm / 1024);
                    time([&] { bulk_kernel<<<148 * ctas_per_sm, 128, smem>>>(d, bytes, ns, cs, streams, o); }, name);
                }
    return 0;
}
