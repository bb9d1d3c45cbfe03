// Developer probe 2 (not product code): the ROI kernel's exact fetch pattern without consumers.
// 32 rows (volumes) of stride `vstride` bytes, each stage = 32 copies of `rowbytes` (+ optional 16B-misaligned starts),
// issued by `pw` producer warps (rows split among them); one waiter warp recycles stages.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok;
    do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(b), "r"(ph) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// mode ldg: same pattern but 8 warps fetch the rows with LDG.128 into registers -> STS (register staged)
__global__ void __launch_bounds__(512) pat_kernel(const char* __restrict__ in, size_t vstride, int ntiles_total, int rowbytes, int ns, int pw, int rows_per_copy_split, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int pitch = (rowbytes + 127) / 128 * 128 + 16;
    uint64_t* full = (uint64_t*)(smem + (size_t)ns * 32 * pitch);
    uint64_t* empty = full + ns;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) { mbar_init(s32(full + s), pw); mbar_init(s32(empty + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // like the ROI kernel: CTA -> (group g = cta % 2, piece = cta / 2); group g reads volumes 32g..32g+31
    const int g = blockIdx.x & 1, piece = blockIdx.x >> 1, pieces = gridDim.x >> 1;
    in += (size_t)g * 32 * vstride;
    const int t0 = (int)((long long)ntiles_total * piece / pieces), t1 = (int)((long long)ntiles_total * (piece + 1) / pieces);
    float acc = 0.f;
    if (warp < pw) {
        const int rpw = 32 / pw;
        const bool act = lane < rpw;
        const int row = warp * rpw + lane;
        uint32_t it = 0;
        for (int t = t0; t < t1; ++t, ++it) {
            const int s = it % ns; const uint32_t ph = (it / ns) & 1;
            mbar_wait(s32(empty + s), ph ^ 1);
            if (lane == 0) mbar_expect(s32(full + s), rpw * rowbytes);
            __syncwarp();
            if (act) bulk(s32(smem + ((size_t)s * 32 + row) * pitch), in + (size_t)row * vstride + (size_t)t * (rowbytes - 16), rowbytes, s32(full + s));
        }
    } else if (warp == pw) {
        uint32_t it = 0;
        for (int t = t0; t < t1; ++t, ++it) {
            const int s = it % ns; const uint32_t ph = (it / ns) & 1;
            mbar_wait(s32(full + s), ph);
            acc += ((float*)(smem + ((size_t)s * 32 + lane) * pitch))[lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(s32(empty + s));
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
int main() {
    const size_t V = 902629;            // voxels
    const int nvol = 64;
    size_t bytes = (size_t)nvol * V * 4 + 4096;
    char* dbuf[3]; float* o;
    for (int i = 0; i < 3; ++i) { cudaMalloc(&dbuf[i], bytes + (1 << 20)); cudaMemset(dbuf[i], 1, bytes); }
    cudaMalloc(&o, 4);
    int rot = 0;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(pat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int aligned : {0, 1})
      for (int rowbytes : {1040, 2064, 4112})
        for (int ns : {2, 4, 6})
          for (int pw : {1, 4, 8}) {
            const int pitch = (rowbytes + 127) / 128 * 128 + 16;
            size_t smem = (size_t)ns * 32 * pitch + 256;
            if (smem > 220 * 1024) continue;
            // aligned: stride multiple of 128 and rows of rowbytes-16 (+16 overlap keeps the arithmetic the same)
            size_t vstride = aligned ? (V * 4 / 128 * 128) : (V * 4 / 16 * 16 + 16 * 1);   // misaligned variant: 16B aligned, != 0 mod 128
            int ntiles = (int)(V * 4 / (rowbytes - 16)) - 1;
            size_t moved = (size_t)ntiles * 32 * rowbytes * 2;   // two groups
            auto f = [&] { pat_kernel<<<148, (pw + 1) * 32, smem>>>(dbuf[rot++ % 3], vstride, ntiles, rowbytes, ns, pw, 0, o); };
            // emulate 2 groups by launching on 148 CTAs over ntiles (each CTA: ntiles/148 tiles of 32 rows); group 2 = second launch omitted, scale bytes
            moved = (size_t)ntiles * 64 * rowbytes;
            for (int i = 0; i < 2; ++i) f();
            cudaEventRecord(a); const int K = 10; for (int i = 0; i < K; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            printf("aligned=%d row=%d ns=%d pw=%d inflight=%zuKB: %7.1f GB/s %s\n", aligned, rowbytes, ns, pw, smem / 1024, moved / (ms / K * 1e-3) / 1e9, e ? cudaGetErrorString(e) : "");
            fflush(stdout);
          }
    return 0;
}
