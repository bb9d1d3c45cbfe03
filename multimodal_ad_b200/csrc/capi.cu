// Library-wide C-ABI plumbing (error string, ABI version, launch counter).
#include "common.cuh"

namespace mmad {
std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
std::atomic<int64_t> g_launches{0};
long long mma_flops_igemm();
long long mma_flops_wgrad();
long long mma_flops_stem();
}  // namespace mmad

extern "C" {
const char* mmad_last_error(void) { return mmad::last_error_ref().c_str(); }
int mmad_abi_version(void) { return 2; }   // 2: round 2 (im2col stem entry points removed, UNet3D entry points added)
int64_t mmad_launch_count(void) { return mmad::g_launches.load(); }
int64_t mmad_executed_mma_flops(void) {
    const long long a = mmad::mma_flops_igemm(), b = mmad::mma_flops_wgrad(), c = mmad::mma_flops_stem();
    if (a < 0 || b < 0 || c < 0) return -1;
    return (int64_t)(a + b + c);
}
}
