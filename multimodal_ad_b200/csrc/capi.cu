// Library-wide C-ABI plumbing (error string, ABI version, launch counter).
#include "common.cuh"

namespace mmad {
std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
std::atomic<int64_t> g_launches{0};
}  // namespace mmad

extern "C" {
const char* mmad_last_error(void) { return mmad::last_error_ref().c_str(); }
int mmad_abi_version(void) { return 1; }
int64_t mmad_launch_count(void) { return mmad::g_launches.load(); }
}
