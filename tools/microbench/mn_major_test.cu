// Does tcgen05.mma kind::f16 accept MN-major (transposed) no-swizzle shared-memory operands, and which of LBO / SBO is the stride
// along M/N and which along K?  One CTA, one MMA (M128 N64 K16, bf16, fp32 accumulate) per variant.
//   operand bytes: [mn-group g][k-row r][8 mn-elements x 2 B]  (core matrix = 8 k-rows x 16 B = 128 B; k-rows dense at 16 B),
//   mn-groups GSTRIDE bytes apart -- the layout of the LSTM kernels' split blocks read "sideways".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I garment_pattern_estimation_b200/csrc -o /tmp/mn_major_test tools/microbench/mn_major_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace nt::tc;

constexpr int M = 128, N = 64, K = 16;
constexpr int GSTRIDE = 2048;          // bytes between 8-element groups along M / N (128 k-rows x 16 B in the real blocks)

__global__ void kern(const uint8_t *a_img, const uint8_t *b_img, int lbo, int sbo, int a_mn, int b_mn, float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa = smem, *sb = smem + (M / 8) * GSTRIDE;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sb + (N / 8) * GSTRIDE);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (M / 8) * GSTRIDE / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sa)[i] = reinterpret_cast<const uint4 *>(a_img)[i];
    for (int i = tid; i < (N / 8) * GSTRIDE / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sb)[i] = reinterpret_cast<const uint4 *>(b_img)[i];
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(slot, 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(M, N, a_mn, b_mn);
        umma_bf16(tmem, make_smem_desc(smem_u32(sa), lbo, sbo), make_smem_desc(smem_u32(sb), lbo, sbo), idesc, 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    float v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) out[tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static uint16_t bf16_bits(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf16_val(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    std::vector<float> A(M * K), B(N * K);
    for (auto &v : A) v = bf16_val(bf16_bits((float)(rand() % 17 - 8) / 4.f));
    for (auto &v : B) v = bf16_val(bf16_bits((float)(rand() % 13 - 6) / 8.f));
    std::vector<uint8_t> ia((M / 8) * GSTRIDE, 0), ib((N / 8) * GSTRIDE, 0);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { uint16_t b = bf16_bits(A[m * K + k]); memcpy(&ia[(m / 8) * GSTRIDE + k * 16 + (m % 8) * 2], &b, 2); }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { uint16_t b = bf16_bits(B[n * K + k]); memcpy(&ib[(n / 8) * GSTRIDE + k * 16 + (n % 8) * 2], &b, 2); }
    std::vector<float> ref(M * N, 0.f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) ref[m * N + n] += A[m * K + k] * B[n * K + k];
    uint8_t *da, *db; float *dout;
    cudaMalloc(&da, ia.size()); cudaMalloc(&db, ib.size()); cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(da, ia.data(), ia.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, ib.data(), ib.size(), cudaMemcpyHostToDevice);
    const size_t smem = (M / 8 + N / 8) * GSTRIDE + 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct { int lbo, sbo; const char *name; } variants[] = {{GSTRIDE, 128, "LBO = mn-group stride, SBO = k 8-row stride"},
                                                              {128, GSTRIDE, "LBO = k 8-row stride, SBO = mn-group stride"}};
    for (auto &v : variants) {
        cudaMemset(dout, 0, M * N * 4);
        kern<<<1, 128, smem>>>(da, db, v.lbo, v.sbo, 1, 1, dout);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> out(M * N);
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        double err = 0, nz = 0;
        for (int i = 0; i < M * N; ++i) { err = fmax(err, fabs(out[i] - ref[i])); nz += out[i] != 0.f; }
        printf("%-48s : %s  max|err| %.4g  nonzeros %d / %d\n", v.name, cudaGetErrorString(e), err, (int)nz, M * N);
    }
    return 0;
}
