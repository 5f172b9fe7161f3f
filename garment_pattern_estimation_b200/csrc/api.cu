// Library-level entry points of libnt_b200.so (see include/nt_b200.h).
#include "common.cuh"

namespace nt {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace nt

extern "C" const char *nt_last_error(void) { return nt::g_err; }
extern "C" int nt_version(void) { return 1; }
extern "C" int nt_built_arch(void) {
#ifdef NT_BUILT_ARCH
    return NT_BUILT_ARCH;
#else
    return 100;
#endif
}
extern "C" int64_t nt_launch_count(void) { return nt::g_launches.load(std::memory_order_relaxed); }
