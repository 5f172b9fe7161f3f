// Shared helpers for libnt_b200.so (sm_100a).  See include/nt_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/nt_b200.h"

namespace nt {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(const char *fmt, const char *a = "", long b = 0, long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return 1;
}

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__host__ __device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace nt

#define NT_REQUIRE(cond, msg)                                  \
    do {                                                       \
        if (!(cond)) return nt::fail("%s", msg);               \
    } while (0)
