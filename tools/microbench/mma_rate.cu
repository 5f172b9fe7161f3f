// Issue rate of tcgen05.mma kind::f16 (M128 x N x K16, bf16) from no-swizzle shared-memory operands: K-major vs MN-major layouts.
// One CTA per SM, thread 0 issues REPS MMAs back to back, clock64 around issue + commit wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I garment_pattern_estimation_b200/csrc -o tools/_build/mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace nt::tc;

__global__ void kern(int n, int lbo, int sbo, int mn, int reps, int tf32, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3c003c00, 0x3c003c00, 0x3c003c00, 0x3c003c00);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = tf32 ? make_idesc_tf32(128, n, mn, mn) : make_idesc_bf16(128, n, mn, mn);
        const uint64_t da = make_smem_desc(smem_u32(smem), lbo, sbo), db = make_smem_desc(smem_u32(smem + 80 * 1024), lbo, sbo);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (tf32) umma_tf32(tmem, da, db, idesc, r ? 1u : 0u);
            else umma_bf16(tmem, da, db, idesc, r ? 1u : 0u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
    long long *d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    struct V { const char *name; int n, lbo, sbo, mn, tf32; } vs[] = {
        {"bf16 K-major  N=256 (LBO 128, SBO 256)", 256, 128, 256, 0, 0},
        {"bf16 K-major  N=208 (LBO 128, SBO 256)", 208, 128, 256, 0, 0},
        {"bf16 K-major  N=208 (LBO 4160, SBO 128) k-chunk planes", 208, 4160, 128, 0, 0},
        {"bf16 MN-major N=256 (LBO 128, SBO 256) dense", 256, 128, 256, 1, 0},
        {"bf16 MN-major N=208 (LBO 128, SBO 272)", 208, 128, 272, 1, 0},
        {"bf16 MN-major N=208 (LBO 128, SBO 256)", 208, 128, 256, 1, 0},
        {"bf16 MN-major N=256 (LBO 128, SBO 512)", 256, 128, 512, 1, 0},
        {"bf16 MN-major N=256 (LBO 128, SBO 528)", 256, 128, 528, 1, 0},
        {"tf32 K-major  N=208 (LBO 128, SBO 256) K=8", 208, 128, 256, 0, 1},
        {"tf32 K-major  N=208 (LBO 3328, SBO 128) K=8", 208, 3328, 128, 0, 1},
    };
    for (auto &v : vs) {
        for (int grid : {1, 148}) {
            const int reps = 256;
            kern<<<grid, 128, 160 * 1024>>>(v.n, v.lbo, v.sbo, v.mn, reps, v.tf32, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%-58s grid %3d: %s  %7.1f cycles / MMA (ideal %d)\n", v.name, grid, cudaGetErrorString(e), (double)mx / reps, v.n / 2);
        }
    }
    return 0;
}
