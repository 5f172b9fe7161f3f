// Developer microbenchmark: cost of tcgen05.mma (no-swizzle K-major operands), of the producer->MMA mbarrier handshake and
// of cp.async.bulk, measured with clock64() on one CTA per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../garment_pattern_estimation_b200/csrc/tc_common.cuh"
using namespace nt::tc;

__global__ void __launch_bounds__(192, 1) k_mma(int n_tile, int iters, int tf32, int mode, const uint8_t *gsrc, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar, full[2], empty[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3f800000u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 128 + (mode == 2 ? 1 : 0)); mbar_init(&empty[s], 1); } mbar_fence_init(); }
    if (warp == 5) tmem_alloc(&slot, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t a = smem_u32(smem), b = a + 16384;
    const uint32_t lbo_a = 128 * 16, lbo_b = n_tile * 16;
    const uint32_t idesc = tf32 ? make_idesc_tf32(128, n_tile, 0, 0) : make_idesc_bf16(128, n_tile, 0, 0);
    long long t0 = 0, t1 = 0;
    if (mode == 0) {                      // back-to-back MMAs, one commit at the end
        if (warp == 4 && lane == 0) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                uint64_t da = make_smem_desc(a + (i & 1) * 2 * lbo_a, lbo_a, 128), db = make_smem_desc(b + (i & 1) * 2 * lbo_b, lbo_b, 128);
                if (tf32) umma_tf32(tm, da, db, idesc, i ? 1u : 0u); else umma_bf16(tm, da, db, idesc, i ? 1u : 0u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    } else {                              // mode 1: full producer<->MMA handshake, 6 MMAs per stage, producers only store+fence+arrive
        if (warp < 4) {                   // mode 2: plus a cp.async.bulk of n_tile*128 bytes per stage
            for (int kb = 0; kb < iters; ++kb) {
                const int s = kb & 1, use = kb >> 1;
                mbar_wait(&empty[s], (use & 1) ^ 1);
                if (mode == 2 && threadIdx.x == 0) { mbar_arrive_expect_tx(&full[s], n_tile * 128); bulk_g2s(smem + 65536 + s * 32768, gsrc + (size_t)(kb & 7) * 32768, n_tile * 128, &full[s]); }
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4 *>(smem + s * 32768 + j * 2048 + threadIdx.x * 16) = make_uint4(1, 2, 3, 4);
                fence_proxy_async();
                mbar_arrive(&full[s]);
            }
        } else if (warp == 4 && lane == 0) {
            t0 = clock64();
            for (int kb = 0; kb < iters; ++kb) {
                const int s = kb & 1, use = kb >> 1;
                mbar_wait(&full[s], use & 1);
                tc_fence_after();
                for (int i = 0; i < 6; ++i) {
                    uint64_t da = make_smem_desc(a + s * 32768 + (i & 1) * 2 * lbo_a, lbo_a, 128), db = make_smem_desc(b + (i & 1) * 2 * lbo_b, lbo_b, 128);
                    if (tf32) umma_tf32(tm, da, db, idesc, (kb | i) ? 1u : 0u); else umma_bf16(tm, da, db, idesc, (kb | i) ? 1u : 0u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 5) tmem_dealloc(tm, 256);
}

int main() {
    long long *d; cudaMalloc(&d, 148 * 8);
    uint8_t *g; cudaMalloc(&g, 8 * 32768); cudaMemset(g, 0, 8 * 32768);
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[148];
    for (int mode = 0; mode < 3; ++mode)
        for (int tf32 = 0; tf32 < 2; ++tf32)
            for (int n : {64, 208, 256}) {
                const int iters = 512;
                k_mma<<<148, 192, 160 * 1024>>>(n, iters, tf32, mode, g, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
                printf("mode %d %s N=%3d: %8.1f cycles per %s   (%s)\n", mode, tf32 ? "tf32" : "bf16", n, avg / iters,
                       mode == 0 ? "MMA" : "stage of 6 MMAs", cudaGetErrorString(e));
            }
    return 0;
}
