// Developer microbenchmark: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st, thread = row = TMEM
// lane, one 32-bit column per K element) against the same product with A in shared memory -- (1) is the result identical, and
// which column offset addresses the second K = 8 step; (2) cycles per MMA, SS vs TS, at the N the row GEMMs use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_mode_test ts_mode_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../garment_pattern_estimation_b200/csrc/tc_common.cuh"
using namespace nt::tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// A [128][16] fp32, B [n][16] fp32 (global, row-major).  mode 0: SS (A from shared memory), mode 1: TS (A from TMEM columns 256..271)
__global__ void __launch_bounds__(192, 1) k_check(const float *A, const float *B, int n, int mode, float *D) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    uint8_t *sa = smem, *sb = smem + 8192;                   // [4 chunks][rows][16 B]
    if (tid < 128)
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4 *>(sa + j * 2048 + tid * 16) = *reinterpret_cast<const float4 *>(A + tid * 16 + 4 * j);
    for (int r = tid; r < n; r += blockDim.x)
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4 *>(sb + j * n * 16 + r * 16) = *reinterpret_cast<const float4 *>(B + r * 16 + 4 * j);
    fence_proxy_async();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 5) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (mode == 1 && warp < 4) {
        uint32_t r[16];
        for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(A[tid * 16 + c]);
        tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + 256, r);
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 4 && lane == 0) {
        const uint32_t idesc = make_idesc_tf32(128, n, 0, 0);
        const uint32_t lbo_a = 128 * 16, lbo_b = n * 16;
        for (int kk = 0; kk < 2; ++kk) {
            const uint64_t db = make_smem_desc(smem_u32(sb) + kk * 2 * lbo_b, lbo_b, 128);
            if (mode == 0) umma_tf32(tm, make_smem_desc(smem_u32(sa) + kk * 2 * lbo_a, lbo_a, 128), db, idesc, kk);
            else umma_tf32_ts(tm, tm + 256 + kk * 8, db, idesc, kk);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        for (int c0 = 0; c0 < n; c0 += 16) {
            float v[16];
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int c = 0; c < 16; ++c) D[tid * n + c0 + c] = v[c];
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 5) tmem_dealloc(tm, 512);
}

// cycles per MMA, back to back: mode 0 SS, mode 1 TS; pattern 3 = the hi.hi + hi.lo + lo.hi triple of the row GEMMs
__global__ void __launch_bounds__(192, 1) k_rate(int n, int iters, int mode, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3f800000u;
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 5) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp < 4) {
        uint32_t r[16];
        for (int c = 0; c < 16; ++c) r[c] = 0x3f800000u;
        tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + 416, r);
        tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + 432, r);
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 4 && lane == 0) {
        const uint32_t idesc = make_idesc_tf32(128, n, 0, 0);
        const uint32_t a = smem_u32(smem), b = a + 32768, lbo_a = 128 * 16, lbo_b = n * 16;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int kk = i & 1;
            const uint64_t dbh = make_smem_desc(b + kk * 2 * lbo_b, lbo_b, 128), dbl = make_smem_desc(b + 4 * lbo_b + kk * 2 * lbo_b, lbo_b, 128);
            if (mode == 0) {
                const uint64_t dah = make_smem_desc(a + kk * 2 * lbo_a, lbo_a, 128), dal = make_smem_desc(a + 8192 + kk * 2 * lbo_a, lbo_a, 128);
                umma_tf32(tm, dah, dbh, idesc, i ? 1u : 0u);
                umma_tf32(tm, dah, dbl, idesc, 1u);
                umma_tf32(tm, dal, dbh, idesc, 1u);
            } else {
                umma_tf32_ts(tm, tm + 416 + kk * 8, dbh, idesc, i ? 1u : 0u);
                umma_tf32_ts(tm, tm + 416 + kk * 8, dbl, idesc, 1u);
                umma_tf32_ts(tm, tm + 432 + kk * 8, dbh, idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 5) tmem_dealloc(tm, 512);
}

int main() {
    cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int n : {16, 64, 208}) {
        float *hA = new float[128 * 16], *hB = new float[n * 16], *hD = new float[128 * n], *hD2 = new float[128 * n];
        srand(n);
        for (int i = 0; i < 128 * 16; ++i) hA[i] = (float)((rand() % 17) - 8) * 0.25f;
        for (int i = 0; i < n * 16; ++i) hB[i] = (float)((rand() % 13) - 6) * 0.5f;
        float *dA, *dB, *dD;
        cudaMalloc(&dA, 128 * 16 * 4); cudaMalloc(&dB, n * 16 * 4); cudaMalloc(&dD, 128 * n * 4);
        cudaMemcpy(dA, hA, 128 * 16 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, n * 16 * 4, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 2; ++mode) {
            cudaMemset(dD, 0xff, 128 * n * 4);
            k_check<<<1, 192, 64 * 1024>>>(dA, dB, n, mode, dD);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(mode ? hD2 : hD, dD, 128 * n * 4, cudaMemcpyDeviceToHost);
            int bad = 0; double maxerr = 0;
            for (int r = 0; r < 128; ++r)
                for (int c = 0; c < n; ++c) {
                    double want = 0;
                    for (int k = 0; k < 16; ++k) want += (double)hA[r * 16 + k] * hB[c * 16 + k];
                    const double err = fabs(want - (mode ? hD2 : hD)[r * n + c]);
                    if (!(err <= 1e-6)) ++bad;
                    if (err > maxerr) maxerr = err;
                }
            printf("check N=%3d %s: %d wrong of %d (max err %g)  [%s]\n", n, mode ? "TS (A in TMEM)" : "SS (A in smem)", bad, 128 * n, maxerr,
                   cudaGetErrorString(e));
        }
        int diff = 0;
        for (int i = 0; i < 128 * n; ++i) diff += hD[i] != hD2[i];
        printf("      N=%3d: %d elements differ between SS and TS\n", n, diff);
    }
    long long *d; cudaMalloc(&d, 148 * 8);
    long long h[148];
    for (int mode = 0; mode < 2; ++mode)
        for (int n : {160, 208, 256}) {
            const int iters = 512;
            k_rate<<<148, 192, 96 * 1024>>>(n, iters, mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            printf("rate %s N=%3d: %7.1f cycles per hi.hi+hi.lo+lo.hi triple = %6.1f per MMA  (%s)\n", mode ? "TS" : "SS", n, avg / iters,
                   avg / iters / 3, cudaGetErrorString(e));
        }
    return 0;
}
