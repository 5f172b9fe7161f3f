// Library-level entry points of libnt_b200.so (see include/nt_b200.h).
#include "common.cuh"

namespace nt {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace nt

extern "C" const char *nt_last_error(void) { return nt::g_err; }
extern "C" int nt_version(void) { return 3; }       // 3: composite nt_edgeconv_train_fwd / _bwd, engine 6; 2: per-call engine, LSTM / PointNet++ / train-step
extern "C" int nt_built_arch(void) {
#ifdef NT_BUILT_ARCH
    return NT_BUILT_ARCH;
#else
    return 100;
#endif
}
extern "C" int64_t nt_launch_count(void) { return nt::g_launches.load(std::memory_order_relaxed); }

// sizeof of the argument structs of this ABI, by name (0 = unknown): lets a binding verify its mirror of the layout at load time
#include <string.h>
extern "C" int nt_sizeof(const char *struct_name) {
    if (!struct_name) return 0;
    if (!strcmp(struct_name, "nt_gemm_args")) return (int)sizeof(nt_gemm_args);
    if (!strcmp(struct_name, "nt_pattern_loss_args")) return (int)sizeof(nt_pattern_loss_args);
    if (!strcmp(struct_name, "nt_lstm_sizes_t")) return (int)sizeof(nt_lstm_sizes_t);
    if (!strcmp(struct_name, "nt_edgeconv_args")) return (int)sizeof(nt_edgeconv_args);
    return 0;
}
