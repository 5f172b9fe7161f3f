// Per-point attention head of GarmentSegmentPattern3D (reference nn/nets.py:223-226,254-279):
//   * sparsemax over the P (=23) panel scores of every point      (sparsemax.Sparsemax(dim=1), nets.py:225)
//   * the 23-iteration "weights * features -> global_mean_pool" loop as ONE contraction enc = W^T F / N
// Both are bandwidth-trivial (rows x 23 and rows x 153 fp32) and run one warp per row / one CTA per point chunk.
#include "common.cuh"

namespace nt {

// One warp per row, lane j holds column j (P <= 32).  Sort-free formulation of the published algorithm:
// for each lane, rank = #{i : z_i > z_j or (z_i == z_j and i < j)} and the sum of those entries; the support
// size is k* = max{ rank_j + 1 : 1 + (rank_j + 1) z_j > sum_j + z_j }; tau = (sum of the k* largest - 1) / k*.
__global__ void sparsemax_fwd_kernel(const float *__restrict__ z, int64_t rows, int P, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const bool live = lane < P;
    float v = live ? z[row * P + lane] : -INFINITY;
    float mx = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    v = live ? v - mx : -INFINITY;
    int rank = 0;
    float sum_gt = 0.f;
    for (int i = 0; i < P; ++i) {
        float zi = __shfl_sync(0xffffffffu, v, i);
        bool before = (zi > v) || (zi == v && i < lane);
        if (before) { rank++; sum_gt += zi; }
    }
    const float kf = (float)(rank + 1);
    int kc = (live && (1.f + kf * v > sum_gt + v)) ? rank + 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kc = max(kc, __shfl_xor_sync(0xffffffffu, kc, o));
    float part = (live && rank < kc) ? v : 0.f;
    part = warp_sum(part);
    const float tau = (part - 1.f) / (float)kc;
    if (live) out[row * P + lane] = fmaxf(0.f, v - tau);
}

__global__ void sparsemax_bwd_kernel(const float *__restrict__ out, const float *__restrict__ g, int64_t rows, int P,
                                     float *__restrict__ gz) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const bool live = lane < P;
    const float o = live ? out[row * P + lane] : 0.f;
    const float gv = live ? g[row * P + lane] : 0.f;
    const bool nz = o != 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    const float s = warp_sum(nz ? gv : 0.f);
    const float mean = s / (float)__popc(m);
    if (live) gz[row * P + lane] = nz ? gv - mean : 0.f;
}

// enc[b, p, f] += scale * sum_{n in chunk} w[b, n, p] * feat[b, n, f]
constexpr int AP_CHUNK = 128;   // points per CTA
constexpr int AP_MAXP = 32;

__global__ void __launch_bounds__(256) attn_pool_fwd_kernel(const float *__restrict__ w, const float *__restrict__ feat,
                                                            int ldf, int N, int P, int F, float scale,
                                                            float *__restrict__ enc) {
    __shared__ float ws[AP_CHUNK][AP_MAXP + 1];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * AP_CHUNK;
    const int nn = min(AP_CHUNK, N - n0);
    for (int i = threadIdx.x; i < nn * P; i += blockDim.x) {
        int n = i / P, p = i - n * P;
        ws[n][p] = w[((int64_t)b * N + n0 + n) * P + p];
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float acc[AP_MAXP];
#pragma unroll
        for (int p = 0; p < AP_MAXP; ++p) acc[p] = 0.f;
        const float *fp = feat + ((int64_t)b * N + n0) * ldf + f;
        for (int n = 0; n < nn; ++n) {
            const float fv = __ldg(fp + (int64_t)n * ldf);
#pragma unroll
            for (int p = 0; p < AP_MAXP; ++p)
                if (p < P) acc[p] = fmaf(ws[n][p], fv, acc[p]);
        }
#pragma unroll
        for (int p = 0; p < AP_MAXP; ++p)
            if (p < P) atomicAdd(enc + ((int64_t)b * P + p) * F + f, acc[p] * scale);
    }
}

constexpr int APB_CHUNK = 32;   // points per CTA in the backward

__global__ void __launch_bounds__(256) attn_pool_bwd_kernel(const float *__restrict__ genc, const float *__restrict__ w,
                                                            const float *__restrict__ feat, int ldf, int N, int P,
                                                            int F, float scale, float *__restrict__ gw,
                                                            float *__restrict__ gfeat, int ldgf, int accumulate) {
    extern __shared__ float sm[];            // genc[b]: [P][F]  then  w chunk: [APB_CHUNK][P]
    float *ge = sm;
    float *wc = sm + P * F;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * APB_CHUNK;
    const int nn = min(APB_CHUNK, N - n0);
    for (int i = threadIdx.x; i < P * F; i += blockDim.x) ge[i] = genc[(int64_t)b * P * F + i];
    for (int i = threadIdx.x; i < nn * P; i += blockDim.x) wc[i] = w[((int64_t)b * N + n0) * P + i];
    __syncthreads();
    // gfeat[b, n, f] = scale * sum_p w[n, p] * genc[p, f]
    if (gfeat) {
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            for (int n = 0; n < nn; ++n) {
                float acc = 0.f;
                for (int p = 0; p < P; ++p) acc = fmaf(wc[n * P + p], ge[p * F + f], acc);
                float *dst = gfeat + ((int64_t)b * N + n0 + n) * ldgf + f;
                *dst = accumulate ? (*dst + acc * scale) : (acc * scale);
            }
        }
    }
    // gw[b, n, p] = scale * sum_f genc[p, f] * feat[n, f]   (one warp per (n, p) pair)
    if (gw) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int pair = warp; pair < nn * P; pair += nwarps) {
            const int n = pair / P, p = pair - n * P;
            const float *fp = feat + ((int64_t)b * N + n0 + n) * ldf;
            float acc = 0.f;
            for (int f = lane; f < F; f += 32) acc = fmaf(ge[p * F + f], __ldg(fp + f), acc);
            acc = warp_sum(acc);
            if (lane == 0) gw[((int64_t)b * N + n0 + n) * P + p] = acc * scale;
        }
    }
}

}  // namespace nt

using namespace nt;

extern "C" int nt_sparsemax_fwd(const float *z, int64_t rows, int P, float *out, void *stream) {
    NT_REQUIRE(z && out && rows >= 0 && P >= 1 && P <= 32, "nt_sparsemax_fwd: need 1 <= P <= 32");
    if (rows == 0) return 0;
    const int64_t threads = rows * 32;
    sparsemax_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(z, rows, P, out);
    return check_launch("nt_sparsemax_fwd");
}

extern "C" int nt_sparsemax_bwd(const float *out, const float *g, int64_t rows, int P, float *gz, void *stream) {
    NT_REQUIRE(out && g && gz && rows >= 0 && P >= 1 && P <= 32, "nt_sparsemax_bwd: need 1 <= P <= 32");
    if (rows == 0) return 0;
    const int64_t threads = rows * 32;
    sparsemax_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, g, rows, P, gz);
    return check_launch("nt_sparsemax_bwd");
}

extern "C" int nt_attn_pool_fwd(const float *w, const float *feat, int ldf, int B, int N, int P, int F, float scale,
                                float *enc, void *stream) {
    NT_REQUIRE(w && feat && enc && B >= 0 && N >= 1 && P >= 1 && P <= AP_MAXP && F >= 1 && ldf >= F,
               "nt_attn_pool_fwd: bad arguments (P <= 32)");
    if (B == 0) return 0;
    dim3 grid((N + AP_CHUNK - 1) / AP_CHUNK, B);
    attn_pool_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, feat, ldf, N, P, F, scale, enc);
    return check_launch("nt_attn_pool_fwd");
}

extern "C" int nt_attn_pool_bwd(const float *genc, const float *w, const float *feat, int ldf, int B, int N, int P,
                                int F, float scale, float *gw, float *gfeat, int ldgf, int accumulate_gfeat,
                                void *stream) {
    NT_REQUIRE(genc && w && feat && B >= 0 && N >= 1 && P >= 1 && P <= AP_MAXP && F >= 1 && ldf >= F,
               "nt_attn_pool_bwd: bad arguments (P <= 32)");
    NT_REQUIRE(!gfeat || ldgf >= F, "nt_attn_pool_bwd: bad gfeat stride");
    if (B == 0) return 0;
    size_t smem = ((size_t)P * F + (size_t)APB_CHUNK * P) * sizeof(float);
    if (smem > 48 * 1024) {
        NT_REQUIRE(smem <= 200 * 1024, "nt_attn_pool_bwd: P*F too large");
        cudaError_t e = cudaFuncSetAttribute(attn_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail("nt_attn_pool_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    dim3 grid((N + APB_CHUNK - 1) / APB_CHUNK, B);
    attn_pool_bwd_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(genc, w, feat, ldf, N, P, F, scale,
                                                                                     gw, gfeat, ldgf, accumulate_gfeat);
    return check_launch("nt_attn_pool_bwd");
}
