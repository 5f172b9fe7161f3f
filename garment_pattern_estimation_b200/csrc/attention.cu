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

// ---------------------------------------------------------------------------------------------------------------------
// Attention pooling.  Three small contractions per cloud (N points, P <= 32 panels, F features):
//   forward   enc[p, f]  = scale * sum_n w[n, p] * feat[n, f]
//   backward  gfeat[n,f] = scale * sum_p w[n, p] * genc[p, f]        gw[n, p] = scale * sum_f genc[p, f] * feat[n, f]
// They are bandwidth-trivial (feat is read once), so the kernels stage 64-point chunks of feat / w (and genc of the cloud)
// in shared memory with coalesced loads, rows padded to a multiple of 4 floats, and run register-tiled FMA loops on
// 16-byte shared-memory reads.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AP_MAXP = 32;
constexpr int AP_PTS = 64;           // points per staged chunk
constexpr int AP_THREADS = 256;
constexpr int AP_FWD_CHUNKS = 4;     // chunks per CTA in the forward (fewer atomics on enc)

__device__ __forceinline__ void ap_stage(float *dst, int pitch, const float *src, int ld, int rows, int cols, int rows_pad) {
    // dst[r][c] = src[r][c] for r < rows, c < cols; zero elsewhere (r < rows_pad, c < pitch).  Warp = 4 rows at a time, lane =
    // column (coalesced); the 4 row loads of a column group are issued back to back so that every thread keeps several
    // independent global loads in flight (the staging phase is latency-bound otherwise).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = AP_THREADS / 32;
    for (int r0 = warp * 4; r0 < rows_pad; r0 += nwarps * 4) {
        for (int c = lane; c < pitch; c += 32) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u;
                v[u] = (r < rows && c < cols) ? __ldg(src + (int64_t)r * ld + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < rows_pad) dst[(r0 + u) * pitch + c] = v[u];
        }
    }
}

// grid (ceil(N / (64*4)), B); dynamic smem: feat_s[64][Fp] + w_s[64][Pp]
__global__ void __launch_bounds__(AP_THREADS) attn_pool_fwd_kernel(const float *__restrict__ w, const float *__restrict__ feat,
                                                                   int ldf, int N, int P, int F, float scale,
                                                                   float *__restrict__ enc) {
    extern __shared__ __align__(16) float sm[];
    const int Fp = (F + 3) & ~3, Pp = (P + 3) & ~3;
    float *feat_s = sm, *w_s = sm + AP_PTS * Fp;
    const int b = blockIdx.y;
    const int nf4 = Fp >> 2, np4 = Pp >> 2, items = nf4 * np4;     // work item = 4 features x 4 panels
    float acc[2][16];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[j][e] = 0.f;
    for (int ch = 0; ch < AP_FWD_CHUNKS; ++ch) {
        const int n0 = (blockIdx.x * AP_FWD_CHUNKS + ch) * AP_PTS;
        if (n0 >= N) break;
        const int nn = min(AP_PTS, N - n0);
        __syncthreads();
        ap_stage(feat_s, Fp, feat + ((int64_t)b * N + n0) * ldf, ldf, nn, F, AP_PTS);
        ap_stage(w_s, Pp, w + ((int64_t)b * N + n0) * P, P, nn, P, AP_PTS);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int item = threadIdx.x + j * AP_THREADS;
            if (item >= items) continue;
            const int f4 = item % nf4, p4 = item / nf4;
            const float *fs = feat_s + 4 * f4, *ws = w_s + 4 * p4;
            for (int n = 0; n < AP_PTS; ++n) {
                const float4 fv = *reinterpret_cast<const float4 *>(fs + n * Fp);
                const float4 wv = *reinterpret_cast<const float4 *>(ws + n * Pp);
                const float ff[4] = {fv.x, fv.y, fv.z, fv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int pi = 0; pi < 4; ++pi)
#pragma unroll
                    for (int fi = 0; fi < 4; ++fi) acc[j][pi * 4 + fi] = fmaf(ww[pi], ff[fi], acc[j][pi * 4 + fi]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int item = threadIdx.x + j * AP_THREADS;
        if (item >= items) continue;
        const int f4 = item % nf4, p4 = item / nf4;
#pragma unroll
        for (int pi = 0; pi < 4; ++pi)
#pragma unroll
            for (int fi = 0; fi < 4; ++fi) {
                const int p = 4 * p4 + pi, f = 4 * f4 + fi;
                if (p < P && f < F) atomicAdd(enc + ((int64_t)b * P + p) * F + f, acc[j][pi * 4 + fi] * scale);
            }
    }
}

// grid (ceil(N / 64), B); dynamic smem: feat_s[64][Fp] (re-used for the gfeat tile) + ge_s[Pp][Fp] + w_s[64][Pp]
__global__ void __launch_bounds__(AP_THREADS) attn_pool_bwd_kernel(const float *__restrict__ genc, const float *__restrict__ w,
                                                                   const float *__restrict__ feat, int ldf, int N, int P,
                                                                   int F, float scale, float *__restrict__ gw,
                                                                   float *__restrict__ gfeat, int ldgf, int accumulate) {
    extern __shared__ __align__(16) float sm[];
    const int Fp = (F + 3) & ~3, Pp = (P + 3) & ~3;
    float *feat_s = sm, *ge_s = sm + AP_PTS * Fp, *w_s = ge_s + Pp * Fp;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * AP_PTS;
    const int nn = min(AP_PTS, N - n0);
    const int nf4 = Fp >> 2, np4 = Pp >> 2;
    ap_stage(ge_s, Fp, genc + (int64_t)b * P * F, F, P, F, Pp);
    ap_stage(w_s, Pp, w + ((int64_t)b * N + n0) * P, P, nn, P, AP_PTS);
    if (gw) ap_stage(feat_s, Fp, feat + ((int64_t)b * N + n0) * ldf, ldf, nn, F, AP_PTS);
    __syncthreads();
    // gw[n, p8 .. p8+7] = scale * sum_f genc[p, f] * feat[n, f]: work item = (point, 8 panels) with the POINT index fastest, so the
    // 32 lanes of a warp read 32 different feat rows (pitch Fp = 4 mod 8 float4s: conflict-free quarter-warps) and ONE genc row
    // address (broadcast).  (ncu of the previous (point, 4 panels)-with-panels-fastest mapping: short-scoreboard bound, 3-way
    // bank conflicts on the genc rows.)
    if (gw) {
        const int np8 = (Pp + 7) >> 3;
        for (int item = threadIdx.x; item < AP_PTS * np8; item += AP_THREADS) {
            const int n = item % AP_PTS, p8 = item / AP_PTS;
            if (n >= nn) continue;
            float a[8];
#pragma unroll
            for (int pi = 0; pi < 8; ++pi) a[pi] = 0.f;
            const float *fs = feat_s + n * Fp;
            const int prow0 = 8 * p8;
            for (int f4 = 0; f4 < nf4; ++f4) {
                const float4 fv = *reinterpret_cast<const float4 *>(fs + 4 * f4);
#pragma unroll
                for (int pi = 0; pi < 8; ++pi) {
                    if (prow0 + pi < Pp) {
                        const float4 gv = *reinterpret_cast<const float4 *>(ge_s + (prow0 + pi) * Fp + 4 * f4);
                        a[pi] = fmaf(gv.x, fv.x, a[pi]); a[pi] = fmaf(gv.y, fv.y, a[pi]);
                        a[pi] = fmaf(gv.z, fv.z, a[pi]); a[pi] = fmaf(gv.w, fv.w, a[pi]);
                    }
                }
            }
#pragma unroll
            for (int pi = 0; pi < 8; ++pi)
                if (prow0 + pi < P) gw[((int64_t)b * N + n0 + n) * P + prow0 + pi] = a[pi] * scale;
        }
    }
    if (gfeat) {
        __syncthreads();                 // feat_s is free now: it receives the gfeat tile for a coalesced store
        // gfeat[n .. n+1, f4 .. f4+3] = scale * sum_p w[n, p] * genc[p, f]: work item = (2 points, 4 features)
        for (int item = threadIdx.x; item < (AP_PTS / 2) * nf4; item += AP_THREADS) {
            const int f4 = item % nf4, n2 = item / nf4;
            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
            const float *w0 = w_s + (2 * n2) * Pp, *w1 = w0 + Pp, *gs = ge_s + 4 * f4;
            for (int p4 = 0; p4 < np4; ++p4) {
                const float4 wa = *reinterpret_cast<const float4 *>(w0 + 4 * p4);
                const float4 wb = *reinterpret_cast<const float4 *>(w1 + 4 * p4);
                const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int pi = 0; pi < 4; ++pi) {
                    const float4 gv = *reinterpret_cast<const float4 *>(gs + (4 * p4 + pi) * Fp);
                    a0[0] = fmaf(wav[pi], gv.x, a0[0]); a0[1] = fmaf(wav[pi], gv.y, a0[1]);
                    a0[2] = fmaf(wav[pi], gv.z, a0[2]); a0[3] = fmaf(wav[pi], gv.w, a0[3]);
                    a1[0] = fmaf(wbv[pi], gv.x, a1[0]); a1[1] = fmaf(wbv[pi], gv.y, a1[1]);
                    a1[2] = fmaf(wbv[pi], gv.z, a1[2]); a1[3] = fmaf(wbv[pi], gv.w, a1[3]);
                }
            }
            *reinterpret_cast<float4 *>(feat_s + (2 * n2) * Fp + 4 * f4) = make_float4(a0[0], a0[1], a0[2], a0[3]);
            *reinterpret_cast<float4 *>(feat_s + (2 * n2 + 1) * Fp + 4 * f4) = make_float4(a1[0], a1[1], a1[2], a1[3]);
        }
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int n = warp; n < nn; n += AP_THREADS / 32) {
            float *dst = gfeat + ((int64_t)b * N + n0 + n) * ldgf;
            for (int f = lane; f < F; f += 32) {
                const float v = feat_s[n * Fp + f] * scale;
                dst[f] = accumulate ? (dst[f] + v) : v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Global pooling over the points of each cloud (torch_geometric.nn.global_{mean,max,add}_pool on the dense equal-size
// layout the reference always builds, nn/net_blocks.py:145-150,182-187).  block (32, 8): lane = feature column (coalesced rows),
// threadIdx.y strides over the points; one CTA per (32-column slab, cloud).  max also records the arg-max point for the backward
// (first maximum wins, like scatter-max).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) global_pool_fwd_kernel(const float *__restrict__ x, int ldx, int N, int F, int mode,
                                                              float *__restrict__ out, int32_t *__restrict__ arg) {
    __shared__ float sv[8][32];
    __shared__ int si[8][32];
    const int f = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
    float acc = (mode == NT_POOL_MAX) ? -INFINITY : 0.f;
    int best = 0;
    if (f < F) {
        const float *xp = x + (int64_t)b * N * ldx + f;
        for (int n = threadIdx.y; n < N; n += 8) {
            const float v = __ldg(xp + (int64_t)n * ldx);
            if (mode == NT_POOL_MAX) {
                if (v > acc) { acc = v; best = n; }
            } else {
                acc += v;
            }
        }
    }
    sv[threadIdx.y][threadIdx.x] = acc;
    si[threadIdx.y][threadIdx.x] = best;
    __syncthreads();
    if (threadIdx.y == 0 && f < F) {
        for (int i = 1; i < 8; ++i) {
            const float v = sv[i][threadIdx.x];
            if (mode == NT_POOL_MAX) {
                const int n = si[i][threadIdx.x];
                if (v > acc || (v == acc && n < best)) { acc = v; best = n; }
            } else {
                acc += v;
            }
        }
        if (mode == NT_POOL_MEAN) acc /= (float)N;
        out[(int64_t)b * F + f] = acc;
        if (arg) arg[(int64_t)b * F + f] = best;
    }
}

__global__ void __launch_bounds__(256) global_pool_bwd_kernel(const float *__restrict__ g, const int32_t *__restrict__ arg, int N,
                                                              int F, int mode, float *__restrict__ gx, int ldgx) {
    const int f = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
    if (f >= F) return;
    const float gv = g[(int64_t)b * F + f] * (mode == NT_POOL_MEAN ? 1.0f / (float)N : 1.0f);
    const int a = (mode == NT_POOL_MAX) ? arg[(int64_t)b * F + f] : -1;
    float *gp = gx + (int64_t)b * N * ldgx + f;
    for (int n = blockIdx.z * 8 + threadIdx.y; n < N; n += 8 * gridDim.z)
        gp[(int64_t)n * ldgx] = (mode == NT_POOL_MAX) ? (n == a ? gv : 0.f) : gv;
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

static size_t ap_fwd_smem(int P, int F) { return ((size_t)AP_PTS * ((F + 3) & ~3) + (size_t)AP_PTS * ((P + 3) & ~3)) * sizeof(float); }
static size_t ap_bwd_smem(int P, int F) {
    const size_t Fp = (F + 3) & ~3, Pp = (P + 3) & ~3;
    return ((size_t)AP_PTS * Fp + Pp * Fp + (size_t)AP_PTS * Pp) * sizeof(float);
}

extern "C" int nt_attn_pool_fwd(const float *w, const float *feat, int ldf, int B, int N, int P, int F, float scale,
                                float *enc, void *stream) {
    NT_REQUIRE(w && feat && enc && B >= 0 && N >= 1 && P >= 1 && P <= AP_MAXP && F >= 1 && ldf >= F,
               "nt_attn_pool_fwd: bad arguments (P <= 32)");
    NT_REQUIRE(B <= 65535, "nt_attn_pool_fwd: at most 65535 clouds per call");
    if (B == 0) return 0;
    const size_t smem = ap_fwd_smem(P, F);
    NT_REQUIRE(((F + 3) / 4) * ((P + 3) / 4) <= 2 * AP_THREADS && smem <= 200 * 1024, "nt_attn_pool_fwd: P*F too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attn_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail("nt_attn_pool_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    dim3 grid((N + AP_PTS * AP_FWD_CHUNKS - 1) / (AP_PTS * AP_FWD_CHUNKS), B);
    attn_pool_fwd_kernel<<<grid, AP_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(w, feat, ldf, N, P, F, scale, enc);
    return check_launch("nt_attn_pool_fwd");
}

extern "C" int nt_attn_pool_bwd(const float *genc, const float *w, const float *feat, int ldf, int B, int N, int P,
                                int F, float scale, float *gw, float *gfeat, int ldgf, int accumulate_gfeat,
                                void *stream) {
    NT_REQUIRE(genc && w && feat && B >= 0 && N >= 1 && P >= 1 && P <= AP_MAXP && F >= 1 && ldf >= F,
               "nt_attn_pool_bwd: bad arguments (P <= 32)");
    NT_REQUIRE(!gfeat || ldgf >= F, "nt_attn_pool_bwd: bad gfeat stride");
    NT_REQUIRE(B <= 65535, "nt_attn_pool_bwd: at most 65535 clouds per call");
    if (B == 0) return 0;
    const size_t smem = ap_bwd_smem(P, F);
    NT_REQUIRE(smem <= 200 * 1024, "nt_attn_pool_bwd: P*F too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attn_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail("nt_attn_pool_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    dim3 grid((N + AP_PTS - 1) / AP_PTS, B);
    attn_pool_bwd_kernel<<<grid, AP_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(genc, w, feat, ldf, N, P, F, scale,
                                                                                            gw, gfeat, ldgf, accumulate_gfeat);
    return check_launch("nt_attn_pool_bwd");
}

extern "C" int nt_global_pool_fwd(const float *x, int ldx, int B, int N, int F, int mode, float *out, int32_t *argmax,
                                  void *stream) {
    NT_REQUIRE(x && out && B >= 0 && N >= 1 && F >= 1 && ldx >= F, "nt_global_pool_fwd: bad arguments");
    NT_REQUIRE(mode == NT_POOL_MEAN || mode == NT_POOL_MAX || mode == NT_POOL_ADD, "nt_global_pool_fwd: unknown mode");
    NT_REQUIRE(B <= 65535, "nt_global_pool_fwd: at most 65535 clouds per call");
    if (B == 0) return 0;
    dim3 grid((F + 31) / 32, B), block(32, 8);
    global_pool_fwd_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, N, F, mode, out, argmax);
    return check_launch("nt_global_pool_fwd");
}

extern "C" int nt_global_pool_bwd(const float *g, const int32_t *argmax, int B, int N, int F, int mode, float *gx, int ldgx,
                                  void *stream) {
    NT_REQUIRE(g && gx && B >= 0 && N >= 1 && F >= 1 && ldgx >= F, "nt_global_pool_bwd: bad arguments");
    NT_REQUIRE(mode == NT_POOL_MEAN || mode == NT_POOL_MAX || mode == NT_POOL_ADD, "nt_global_pool_bwd: unknown mode");
    NT_REQUIRE(mode != NT_POOL_MAX || argmax, "nt_global_pool_bwd: max pooling needs the arg-max of the forward");
    NT_REQUIRE(B <= 65535, "nt_global_pool_bwd: at most 65535 clouds per call");
    if (B == 0) return 0;
    dim3 grid((F + 31) / 32, B, (N + 255) / 256 > 64 ? 64 : (N + 255) / 256), block(32, 8);
    global_pool_bwd_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, argmax, N, F, mode, gx, ldgx);
    return check_launch("nt_global_pool_bwd");
}
