// Brute-force kNN graph build (distance + top-k fused), sm_100a.
//
// Replaces torch_cluster.knn as reached from DynamicEdgeConv.forward (reference nn/net_blocks.py:127-135,174).
// Bit-exactness contract (oracle/knn_oracle.c): squared distance = sequential fp32 chain
//     acc = fma(c_d - q_d, c_d - q_d, acc),  d = 0..D-1
// and the result is the k lexicographically smallest (distance, index) pairs, ascending.
//
// Mapping: one thread owns one query and its private sorted top-K list in registers; a CTA of 64 queries of one
// cloud sweeps the cloud's candidates in tiles of TC rows staged in shared memory.  Every lane reads the SAME
// candidate word (shared-memory broadcast), so the kernel is bound by the FP32 pipe: 2 instructions (FADD + FFMA)
// per pair-dimension, the minimum the bit-exact direct form allows.  TC independent chains per thread give the ILP.
// Zero padding (candidate and query padded to a multiple of 4 dims with 0) is exact: fma(0,0,acc) == acc.
#include "common.cuh"
#include <stdlib.h>

namespace nt {

constexpr int KNN_TC = 32;        // candidates per tile (independent fma chains per thread)
constexpr int KNN_DC = 16;        // dims per query-register chunk
constexpr int KNN_PRUNE_EVERY = 2;  // re-evaluate the warp-level alive mask every 2 chunks (32 dims)

template <int K>
__device__ __forceinline__ void topk_insert(float (&ld)[K], int (&li)[K], float d, int c) {
    // caller guarantees d < ld[K-1]; insert before the first entry that is strictly greater (stable for ties)
#pragma unroll
    for (int e = K - 1; e >= 0; --e) {
        bool prev_gt = (e > 0) ? (ld[e > 0 ? e - 1 : 0] > d) : false;
        bool cur_gt = ld[e] > d;
        float nd = prev_gt ? ld[e > 0 ? e - 1 : 0] : (cur_gt ? d : ld[e]);
        int ni = prev_gt ? li[e > 0 ? e - 1 : 0] : (cur_gt ? c : li[e]);
        ld[e] = nd;
        li[e] = ni;
    }
}

// One dim-chunk (NG float4 groups) of the fma chains of the candidates whose bit is set in `alive` (warp-uniform).
// Candidates are handled in pairs so two independent chains interleave inside every branch.
template <int NG, bool PRUNE>
__device__ __forceinline__ void chain_chunk(float (&acc)[KNN_TC], unsigned alive, const float *__restrict__ tile, int Dp,
                                            int d0, const float (&qv)[KNN_DC]) {
#pragma unroll
    for (int c = 0; c < KNN_TC; c += 2) {
        if (!PRUNE || ((alive >> c) & 3u)) {
            const float4 *r0 = reinterpret_cast<const float4 *>(tile + c * Dp + d0);
            const float4 *r1 = reinterpret_cast<const float4 *>(tile + (c + 1) * Dp + d0);
            float a0 = acc[c], a1 = acc[c + 1];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4 u = r0[g], w = r1[g];
                float d, e;
                d = __fsub_rn(u.x, qv[4 * g + 0]); e = __fsub_rn(w.x, qv[4 * g + 0]); a0 = __fmaf_rn(d, d, a0); a1 = __fmaf_rn(e, e, a1);
                d = __fsub_rn(u.y, qv[4 * g + 1]); e = __fsub_rn(w.y, qv[4 * g + 1]); a0 = __fmaf_rn(d, d, a0); a1 = __fmaf_rn(e, e, a1);
                d = __fsub_rn(u.z, qv[4 * g + 2]); e = __fsub_rn(w.z, qv[4 * g + 2]); a0 = __fmaf_rn(d, d, a0); a1 = __fmaf_rn(e, e, a1);
                d = __fsub_rn(u.w, qv[4 * g + 3]); e = __fsub_rn(w.w, qv[4 * g + 3]); a0 = __fmaf_rn(d, d, a0); a1 = __fmaf_rn(e, e, a1);
            }
            acc[c] = a0; acc[c + 1] = a1;
        }
    }
}

__device__ __forceinline__ void load_query_chunk(const float *__restrict__ xq, int d0, int groups, int D, bool q_ok, bool qvec,
                                                 float (&qv)[KNN_DC]) {
#pragma unroll
    for (int g = 0; g < KNN_DC / 4; ++g) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < groups && q_ok) {
            const int d = d0 + 4 * g;
            if (qvec && d + 3 < D) {
                v = __ldg(reinterpret_cast<const float4 *>(xq + d));
            } else if (qvec && d + 2 == D) {
                const float2 t = __ldg(reinterpret_cast<const float2 *>(xq + d));
                v.x = t.x; v.y = t.y;
            } else {
                if (d + 0 < D) v.x = __ldg(xq + d + 0);
                if (d + 1 < D) v.y = __ldg(xq + d + 1);
                if (d + 2 < D) v.z = __ldg(xq + d + 2);
                if (d + 3 < D) v.w = __ldg(xq + d + 3);
            }
        }
        qv[4 * g + 0] = v.x; qv[4 * g + 1] = v.y; qv[4 * g + 2] = v.z; qv[4 * g + 3] = v.w;
    }
}

// Exactness of the pruning: every term of the chain is >= 0 and fmaf rounding is monotone, so the partial sums never
// decrease.  A candidate whose partial sum is already >= the current k-th best distance of EVERY lane of the warp can never
// satisfy the strict insertion test `dist < kth` for any of them, so its remaining dimensions are skipped; the (partial)
// value it keeps still fails that test.  Thresholds only tighten, so a stale alive bit is merely conservative.
// DT > 0: the feature count is a compile-time constant (rows 16-byte aligned, checked at launch) so every bounds test of
// the query / tile loads folds away -- with a runtime D those guards were ~25 % of the issued instructions (ncu, round 1).
template <int K, int KNN_THREADS, bool PRUNE, bool PREFETCH, int DT>
__global__ void __launch_bounds__(KNN_THREADS, (K <= 16) ? 512 / KNN_THREADS : 256 / KNN_THREADS)
knn_kernel(const float *__restrict__ x, int N, int D_rt, int ldx, int k, int32_t *__restrict__ idx,
           float *__restrict__ part_d, int32_t *__restrict__ part_i, int cand_per_split) {
    extern __shared__ __align__(16) float tile[];   // [KNN_TC][Dp]
    const int D = DT ? DT : D_rt;
    const int Dp = (D + 3) & ~3;
    const int b = blockIdx.y;
    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool q_ok = q < N;
    const float *cloud = x + (size_t)b * N * ldx;
    const float *xq = cloud + (size_t)(q_ok ? q : 0) * ldx;
    const bool qvec = DT ? true : (((ldx & 3) == 0) && aligned16(x));
    const bool q_ld = DT ? true : q_ok;     // with DT idle lanes read row 0 (in bounds) instead of branching

    float ld[K];
    int li[K];
#pragma unroll
    for (int e = 0; e < K; ++e) { ld[e] = 1e10f; li[e] = -1; }

    const int full_chunks = Dp / KNN_DC;
    const int tail_groups = (Dp - full_chunks * KNN_DC) >> 2;
    const int n_chunks = full_chunks + (tail_groups ? 1 : 0);

    float qv[KNN_DC], qn[KNN_DC];
    if (PREFETCH) load_query_chunk(xq, 0, (0 < full_chunks) ? KNN_DC / 4 : tail_groups, D, q_ld, qvec, qn);

    // blockIdx.z = slice of the candidate range (see knn_merge_kernel): finer work items even out the SM load
    const int c_begin = blockIdx.z * cand_per_split;
    const int c_end = min(N, c_begin + cand_per_split);
    for (int c0 = c_begin; c0 < c_end; c0 += KNN_TC) {
        __syncthreads();
        for (int i = threadIdx.x; i < KNN_TC * Dp; i += KNN_THREADS) {
            int c = i / Dp, d = i - c * Dp;
            float v = 0.f;
            if (c0 + c < c_end && d < D) v = __ldg(cloud + (size_t)(c0 + c) * ldx + d);
            tile[i] = v;
        }
        __syncthreads();

        float acc[KNN_TC];
#pragma unroll
        for (int c = 0; c < KNN_TC; ++c) acc[c] = 0.f;
        unsigned alive = (c0 + KNN_TC <= c_end) ? 0xffffffffu : ((1u << (c_end - c0)) - 1u);

        for (int ch = 0; ch < n_chunks; ++ch) {
            const int d0 = ch * KNN_DC;
            const int groups = (ch < full_chunks) ? (KNN_DC / 4) : tail_groups;
            if (PREFETCH) {
#pragma unroll
                for (int i = 0; i < KNN_DC; ++i) qv[i] = qn[i];
                // prefetch the query values of the NEXT chunk (wrapping to chunk 0 for the next tile)
                const int nch = (ch + 1 < n_chunks) ? ch + 1 : 0;
                load_query_chunk(xq, nch * KNN_DC, (nch < full_chunks) ? KNN_DC / 4 : tail_groups, D, q_ld, qvec, qn);
            } else {
                load_query_chunk(xq, d0, groups, D, q_ld, qvec, qv);
            }
            switch (groups) {
                case 4: chain_chunk<4, PRUNE>(acc, alive, tile, Dp, d0, qv); break;
                case 3: chain_chunk<3, PRUNE>(acc, alive, tile, Dp, d0, qv); break;
                case 2: chain_chunk<2, PRUNE>(acc, alive, tile, Dp, d0, qv); break;
                default: chain_chunk<1, PRUNE>(acc, alive, tile, Dp, d0, qv); break;
            }
            if (PRUNE && (ch % KNN_PRUNE_EVERY) == KNN_PRUNE_EVERY - 1 && ch + 1 < n_chunks) {
                const float thr = q_ok ? ld[K - 1] : -1.f;          // idle lanes never keep a candidate alive
                unsigned keep = 0;
#pragma unroll
                for (int c = 0; c < KNN_TC; ++c)
                    if (__any_sync(0xffffffffu, acc[c] < thr)) keep |= (1u << c);
                alive &= keep;
            }
        }

#pragma unroll
        for (int c = 0; c < KNN_TC; ++c) {
            if (c0 + c < c_end && acc[c] < ld[K - 1]) topk_insert<K>(ld, li, acc[c], c0 + c);
        }
    }

    if (q_ok) {
        if (gridDim.z == 1) {
            int32_t *o = idx + ((size_t)b * N + q) * k;
#pragma unroll
            for (int e = 0; e < K; ++e)
                if (e < k) o[e] = li[e];
        } else {
            const size_t o = (((size_t)b * N + q) * gridDim.z + blockIdx.z) * K;
#pragma unroll
            for (int e = 0; e < K; ++e) { part_d[o + e] = ld[e]; part_i[o + e] = li[e]; }
        }
    }
}

// exact merge of S sorted partial lists per query: slices are visited in ascending candidate order, so the stable
// strict-greater insertion reproduces the sequential scan (equal distances keep the lower index first)
template <int K>
__global__ void knn_merge_kernel(const float *__restrict__ part_d, const int32_t *__restrict__ part_i, int64_t M, int S, int k,
                                 int32_t *__restrict__ idx) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= M) return;
    float ld[K];
    int li[K];
#pragma unroll
    for (int e = 0; e < K; ++e) { ld[e] = 1e10f; li[e] = -1; }
    for (int s = 0; s < S; ++s) {
        const size_t o = ((size_t)q * S + s) * K;
        for (int e = 0; e < K; ++e) {
            const float d = part_d[o + e];
            if (d < ld[K - 1]) topk_insert<K>(ld, li, d, part_i[o + e]);
            else break;                                      // the slice list is ascending
        }
    }
#pragma unroll
    for (int e = 0; e < K; ++e)
        if (e < k) idx[q * k + e] = li[e];
}

template <int K, int THREADS, bool PRUNE, bool PREFETCH, int DT>
static int launch_knn_dt(const float *x, int B, int N, int D, int ldx, int k, int32_t *idx, void *workspace, int S,
                         cudaStream_t st) {
    const int Dp = (D + 3) & ~3;
    size_t smem = (size_t)KNN_TC * Dp * sizeof(float);
    if (smem > 200 * 1024) return fail("nt_knn: feature dimension %s too large for the staging tile (D=%ld)", "", D);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(knn_kernel<K, THREADS, PRUNE, PREFETCH, DT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail("nt_knn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    if (!workspace || S < 1) S = 1;
    const int per = ((N + S - 1) / S + KNN_TC - 1) / KNN_TC * KNN_TC;
    S = (N + per - 1) / per;
    const int64_t M = (int64_t)B * N;
    float *part_d = reinterpret_cast<float *>(workspace);
    int32_t *part_i = reinterpret_cast<int32_t *>(part_d + (size_t)M * S * K);
    dim3 grid((N + THREADS - 1) / THREADS, B, S);
    knn_kernel<K, THREADS, PRUNE, PREFETCH, DT><<<grid, THREADS, smem, st>>>(x, N, D, ldx, k, idx, part_d, part_i, per);
    int rc = check_launch("nt_knn");
    if (rc || S == 1) return rc;
    knn_merge_kernel<K><<<(unsigned)((M + 255) / 256), 256, 0, st>>>(part_d, part_i, M, S, k, idx);
    return check_launch("nt_knn(merge)");
}

// Candidate split: the one-thread-per-query grid has only B*N/THREADS CTAs (256 at C2 for 296 resident slots: 108 SMs get
// two CTAs, 40 get one).  Splitting the scan into S slices makes ~7 work items per slot, which the hardware CTA scheduler
// balances dynamically; capped so a slice keeps at least 4 tiles.
static int knn_auto_split(int B, int N, int threads) {
    const long ctas = (long)B * ((N + threads - 1) / threads);
    const long slots = 148L * (512 / threads);
    int S = 1;
    while (S < 8 && ctas * S < 6 * slots && N / (2 * S) >= 4 * KNN_TC) S *= 2;
    return S;
}

template <int K, int THREADS, bool PRUNE, bool PREFETCH>
static int launch_knn_cfg(const float *x, int B, int N, int D, int ldx, int k, int32_t *idx, void *workspace, int S,
                          cudaStream_t st) {
    // specialised feature counts of the shipped configs (EConv_feature = 150); needs 16-byte aligned rows
    const bool aligned = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    if (S <= 0) S = knn_auto_split(B, N, THREADS);
    if (D == 150 && aligned && !PRUNE && !PREFETCH)
        return launch_knn_dt<K, THREADS, false, false, 150>(x, B, N, D, ldx, k, idx, workspace, S, st);
    return launch_knn_dt<K, THREADS, PRUNE, PREFETCH, 0>(x, B, N, D, ldx, k, idx, workspace, S, st);
}

template <int K>
static int launch_knn(const float *x, int B, int N, int D, int ldx, int k, int32_t *idx, void *ws, cudaStream_t st) {
    return launch_knn_cfg<K, 256, false, false>(x, B, N, D, ldx, k, idx, ws, 0, st);      // auto candidate split
}

// tensor-core filter + exact re-rank path (knn_tc.cu)
bool knn_tc_eligible(int N, int D, const void *workspace);
int64_t knn_tc_workspace_bytes(int B, int N, int k);
int knn_tc_run(const float *x, int B, int N, int D, int ldx, int k, int32_t *idx, void *workspace, cudaStream_t st);
// xyz graphs (knn3.cu)
int knn3_run(const float *x, int B, int N, int ldx, int k, int32_t *idx, cudaStream_t st);

}  // namespace nt

extern "C" int64_t nt_knn_workspace_bytes(int B, int N, int k) {
    if (B < 1 || N < 1 || k < 1) return 0;
    const int K = k <= 5 ? 5 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32));
    const int64_t split_bytes = (int64_t)B * N * 8 /*max split*/ * K * 8 /*dist + index*/;
    const int64_t tc_bytes = nt::knn_tc_workspace_bytes(B, N, k);
    return split_bytes > tc_bytes ? split_bytes : tc_bytes;
}

extern "C" int nt_knn(const float *x, int B, int N, int D, int ldx, int k, int32_t *idx, void *workspace, void *stream) {
    using namespace nt;
    NT_REQUIRE(x && idx, "nt_knn: null pointer");
    NT_REQUIRE(B >= 0 && N >= 0 && D >= 1 && ldx >= D, "nt_knn: bad shape");
    NT_REQUIRE(k >= 1 && k <= 32, "nt_knn: k must be in [1, 32]");
    NT_REQUIRE(k <= N || N == 0, "nt_knn: k must not exceed the number of points per cloud");
    NT_REQUIRE(B <= 65535, "nt_knn: at most 65535 clouds per call");
    if (B == 0 || N == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // feature-space graphs (D >= 8): bf16 tcgen05 filter + exact fp32 re-rank, bit-identical to the direct form below
    if (knn_tc_eligible(N, D, workspace)) return knn_tc_run(x, B, N, D, ldx, k, idx, workspace, st);
    // coordinate-space graphs (D = 3): register top-k over a shared-memory copy of the cloud (knn3.cu)
    if (D == 3) return knn3_run(x, B, N, ldx, k, idx, st);
    // everything else (4 <= D < 8, or no workspace): the generic direct-form kernel of this file
    if (k <= 5) return launch_knn<5>(x, B, N, D, ldx, k, idx, workspace, st);
    if (k <= 8) return launch_knn<8>(x, B, N, D, ldx, k, idx, workspace, st);
    if (k <= 16) return launch_knn<16>(x, B, N, D, ldx, k, idx, workspace, st);
    return launch_knn<32>(x, B, N, D, ldx, k, idx, workspace, st);
}
