// Tensor-core weight-gradient GEMM (sm_100a, tcgen05 + TMEM):   out[m, n] (+)= sum_r A[r, m] * Bop[r, n]
//
// The reduction runs over ROWS (points / edges, up to millions) and the output is small (<= 400 x 300), so both operands
// are "MN-major" for the tensor core: a row r of A / Bop is contiguous along m / n.  Each CTA owns a contiguous slice of
// rows, streams it through a 3-stage shared-memory ring in 16-row stages (TF32x3 error-compensated split, see gemm_tc.cu)
// and accumulates a full [<=256 x <=256] partial product in TMEM (2 M-tiles of 128 lanes).  Partials go to a workspace
// ([splits, m_pad, n_pad] fp32, coalesced) and a second kernel reduces them in double -- deterministic, no atomics.
//
// Bop = plain matrix, or the EdgeConv edge activation relu(P[centre] + Q[nbr]) gathered on the fly, optionally centred
// per column (Bop - mu) for the BatchNorm-backward moments (see nt_linear_bn_bwd).
#include "gemm_params.cuh"
#include "tc_common.cuh"

namespace nt {
using namespace tc;

constexpr int TN_THREADS = 192;
constexpr int TN_RB = 16;                       // rows (K of the MMA) per stage = 2 tf32 k-steps of 8
constexpr int TN_STAGES = 3;
constexpr int TN_SBO = TN_RB * 16 + 16;         // 272 B between 4-element m/n groups (+16 B pad: conflict-free STS)
constexpr int TN_MAX_M = 256, TN_MAX_N = 256;

struct TNTCParams {
    const float *a; int lda; int m;              // A: [rows, m]
    const float *b; int ldb; int n;              // plain Bop
    EdgeSrc e; int b_edge;                       // gathered Bop
    const float *mu;                             // optional centring
    int64_t rows; int64_t rows_per_split;
    int m_pad, n_pad;                            // m_pad in {128, 256}, n_pad multiple of 16
    int m0, n0;                                  // tile origin inside the full [m, n] output (grid.y / grid.z)
    float *partial;                              // [splits, m_pad, n_pad]
};

__device__ __forceinline__ void tn_store_chunk(uint8_t *hi_plane, uint8_t *lo_plane, int group, int r, const float (&v)[4]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
    const int off = group * TN_SBO + r * 16;
    *reinterpret_cast<uint4 *>(hi_plane + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(lo_plane + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(TN_THREADS, 1) gemm_tn_tc_kernel(TNTCParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ga = p.m_pad / 4, gb = p.n_pad / 4;                 // 16-byte groups per row of A / B
    const size_t a_plane = (size_t)ga * TN_SBO, b_plane = (size_t)gb * TN_SBO;
    const size_t stage_bytes = 2 * a_plane + 2 * b_plane;
    uint8_t *tail = smem + TN_STAGES * stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty = full + TN_STAGES;
    uint64_t *tmem_full = empty + TN_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_tiles = p.m_pad / 128;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < m_tiles * p.n_pad) tmem_cols <<= 1;

    if (warp == 4 && lane == 0) {
        for (int s = 0; s < TN_STAGES; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_split;
    const int64_t r_end = min(p.rows, r_begin + p.rows_per_split);
    const int n_stages = (int)((max((int64_t)0, r_end - r_begin) + TN_RB - 1) / TN_RB);

    if (warp < 4) {
        // =========================== producers: 8 threads per row, groups strided by 8 ===========================
        const int rl = tid >> 3;                 // row inside the stage, 0..15
        const int g0 = tid & 7;
        const bool veca = ((p.lda & 3) == 0) && aligned16(p.a) && ((p.m0 & 3) == 0);
        const bool vecb = p.b_edge ? (((p.e.ldpq & 3) == 0) && ((p.e.qoff & 3) == 0) && aligned16(p.e.pq) && ((p.n0 & 3) == 0))
                                   : (((p.ldb & 3) == 0) && aligned16(p.b) && ((p.n0 & 3) == 0));
        constexpr int MAXA = TN_MAX_M / 4 / 8, MAXB = TN_MAX_N / 4 / 8;     // 8 groups per thread each
        float va[MAXA][4], vb[MAXB][4];

        auto load4 = [&](const float *row, int col, int width, bool vec, float (&v)[4]) {
            if (vec && col + 3 < width) {
                float4 t = __ldg(reinterpret_cast<const float4 *>(row + col));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (col + e < width) ? __ldg(row + col + e) : 0.f;
            }
        };
        auto fetch = [&](int st) {
            const int64_t r = r_begin + (int64_t)st * TN_RB + rl;
            const bool ok = r < r_end;
            const float *arow = ok ? p.a + r * p.lda : nullptr;
            const float *bp = nullptr, *bq = nullptr;
            if (ok) {
                if (p.b_edge) edge_row_ptrs(p.e, r, bp, bq);
                else bp = p.b + r * p.ldb;
            }
#pragma unroll
            for (int i = 0; i < MAXA; ++i) {
                const int g = g0 + 8 * i;
#pragma unroll
                for (int e = 0; e < 4; ++e) va[i][e] = 0.f;
                if (ok && g < ga) load4(arow, p.m0 + g * 4, p.m, veca, va[i]);
            }
#pragma unroll
            for (int i = 0; i < MAXB; ++i) {
                const int g = g0 + 8 * i;
#pragma unroll
                for (int e = 0; e < 4; ++e) vb[i][e] = 0.f;
                if (ok && g < gb) {
                    const int col = p.n0 + g * 4;
                    load4(bp, col, p.n, vecb, vb[i]);
                    if (p.b_edge) {
                        if (bq) {
                            float q[4];
                            load4(bq, col, p.n, vecb, q);
#pragma unroll
                            for (int e = 0; e < 4; ++e) vb[i][e] += q[e];
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) vb[i][e] = fmaxf(vb[i][e], 0.f);
                    }
                    if (p.mu) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (col + e < p.n) vb[i][e] -= __ldg(p.mu + col + e);
                    }
                }
            }
        };

        if (n_stages > 0) fetch(0);
        for (int st = 0; st < n_stages; ++st) {
            const int s = st % TN_STAGES, use = st / TN_STAGES;
            mbar_wait(&empty[s], (use & 1) ^ 1);
            uint8_t *a_hi = smem + s * stage_bytes, *a_lo = a_hi + a_plane, *b_hi = a_lo + a_plane, *b_lo = b_hi + b_plane;
#pragma unroll
            for (int i = 0; i < MAXA; ++i) {
                const int g = g0 + 8 * i;
                if (g < ga) tn_store_chunk(a_hi, a_lo, g, rl, va[i]);
            }
#pragma unroll
            for (int i = 0; i < MAXB; ++i) {
                const int g = g0 + 8 * i;
                if (g < gb) tn_store_chunk(b_hi, b_lo, g, rl, vb[i]);
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
            if (st + 1 < n_stages) fetch(st + 1);
        }

        // =========================== epilogue: TMEM -> coalesced partial tile ===========================
        float *tw = reinterpret_cast<float *>(smem) + warp * (32 * 33);          // stage memory is free now
        float *dst_tile = p.partial + (size_t)blockIdx.x * p.m_pad * p.n_pad;
        if (n_stages > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        for (int mt = 0; mt < m_tiles; ++mt) {
            for (int c0 = 0; c0 < p.n_pad; c0 += 32) {
                float acc[32];
                if (n_stages > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * p.n_pad + c0), acc);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) tw[lane * 33 + i] = acc[i];
                __syncwarp();
                if (c0 + lane < p.n_pad) {
                    float *dst = dst_tile + (size_t)(mt * 128 + warp * 32) * p.n_pad + c0 + lane;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) dst[(size_t)rr * p.n_pad] = tw[rr * 33 + lane];
                }
                __syncwarp();
            }
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, (uint32_t)p.n_pad, 1, 1);        // both operands MN-major
            for (int st = 0; st < n_stages; ++st) {
                const int s = st % TN_STAGES, use = st / TN_STAGES;
                mbar_wait(&full[s], use & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * stage_bytes), a_lo = a_hi + (uint32_t)a_plane;
                const uint32_t b_hi = a_lo + (uint32_t)a_plane, b_lo = b_hi + (uint32_t)b_plane;
#pragma unroll
                for (int ks = 0; ks < TN_RB / 8; ++ks) {
                    const uint64_t dbh = make_smem_desc(b_hi + ks * 128, 128, TN_SBO);
                    const uint64_t dbl = make_smem_desc(b_lo + ks * 128, 128, TN_SBO);
                    for (int mt = 0; mt < m_tiles; ++mt) {
                        const uint32_t moff = (uint32_t)mt * 32u * TN_SBO;               // 32 groups of 4 = 128 m-elements
                        const uint64_t dah = make_smem_desc(a_hi + moff + ks * 128, 128, TN_SBO);
                        const uint64_t dal = make_smem_desc(a_lo + moff + ks * 128, 128, TN_SBO);
                        const uint32_t d = tmem_base + (uint32_t)(mt * p.n_pad);
                        umma_tf32(d, dah, dbh, idesc, (st | ks) ? 1u : 0u);
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dal, dbh, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);
            }
            if (n_stages > 0) umma_commit(tmem_full);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, tmem_cols);
}

// out[m, n] (+)= sum_s partial[s, m - m0, n - n0]     (double accumulation; OutT = float or double)
template <typename OutT>
__global__ void tn_reduce_kernel(const float *__restrict__ partial, int splits, int m_pad, int n_pad, int m0, int n0,
                                 int m, int n, OutT *__restrict__ out, int ldo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_pad * n_pad) return;
    const int mm = i / n_pad, nn = i - mm * n_pad;
    if (m0 + mm >= m || n0 + nn >= n) return;
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += (double)partial[(size_t)s * m_pad * n_pad + i];
    OutT *dst = out + (size_t)(m0 + mm) * ldo + n0 + nn;
    *dst = (OutT)((double)*dst + acc);
}

static int tn_splits(int64_t rows) {
    int64_t s = (rows + 8 * TN_RB - 1) / (8 * TN_RB);          // at least 8 stages per CTA
    if (s > 148) s = 148;
    if (s < 1) s = 1;
    return (int)s;
}

int gemm_tn_tc(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows, const EdgeSrc &e, int b_edge,
               const float *mu, void *out, int out_double, int ldo, float *workspace, cudaStream_t st) {
    const int splits = tn_splits(rows);
    int64_t rps = (rows + splits - 1) / splits;
    rps = ((rps + TN_RB - 1) / TN_RB) * TN_RB;
    for (int m0 = 0; m0 < m; m0 += TN_MAX_M) {
        for (int n0 = 0; n0 < n; n0 += TN_MAX_N) {
            TNTCParams p{};
            p.a = a; p.lda = lda; p.m = m; p.b = b; p.ldb = ldb; p.n = n; p.e = e; p.b_edge = b_edge; p.mu = mu;
            p.rows = rows; p.rows_per_split = rps; p.m0 = m0; p.n0 = n0;
            p.m_pad = (m - m0 > 128) ? 256 : 128;
            p.n_pad = ((min(n - n0, TN_MAX_N) + 15) / 16) * 16;
            p.partial = workspace;
            const size_t stage_bytes = 2 * (size_t)(p.m_pad / 4) * TN_SBO + 2 * (size_t)(p.n_pad / 4) * TN_SBO;
            const size_t smem = TN_STAGES * stage_bytes + 128;
            static bool configured = false;
            if (!configured) {
                cudaError_t err = cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (err != cudaSuccess) return fail("nt_gemm_tn(tc): cudaFuncSetAttribute: %s", cudaGetErrorString(err));
                configured = true;
            }
            if (smem > 227 * 1024) return fail("nt_gemm_tn(tc): tile does not fit shared memory%s", "");
            gemm_tn_tc_kernel<<<splits, TN_THREADS, smem, st>>>(p);
            int rc = check_launch("nt_gemm_tn(tc)");
            if (rc) return rc;
            const int total = p.m_pad * p.n_pad;
            if (out_double)
                tn_reduce_kernel<double><<<(total + 255) / 256, 256, 0, st>>>(workspace, splits, p.m_pad, p.n_pad, m0, n0, m, n,
                                                                            reinterpret_cast<double *>(out), ldo);
            else
                tn_reduce_kernel<float><<<(total + 255) / 256, 256, 0, st>>>(workspace, splits, p.m_pad, p.n_pad, m0, n0, m, n,
                                                                           reinterpret_cast<float *>(out), ldo);
            rc = check_launch("nt_gemm_tn(reduce)");
            if (rc) return rc;
        }
    }
    return 0;
}

}  // namespace nt

extern "C" int64_t nt_gemm_tn_workspace_bytes(void) {
    return (int64_t)148 * nt::TN_MAX_M * nt::TN_MAX_N * (int64_t)sizeof(float);
}
