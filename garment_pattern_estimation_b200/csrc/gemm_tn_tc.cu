// Tensor-core weight-gradient GEMM (sm_100a, tcgen05 + TMEM):   out[m, n] (+)= sum_r A[r, m] * Bop[r, n]
//
// The reduction runs over ROWS (points / edges, up to millions) and the output is small (<= 400 x 300).  In memory a row r
// of A / Bop is contiguous along m / n, i.e. the operands are MN-major for the tensor core; the producers TRANSPOSE them
// on the way into shared memory (lane = m, four scalar loads from four consecutive rows -> one 16-byte K-chunk), so the
// MMAs see the same K-major no-swizzle core-matrix layout as gemm_tc.cu.  (tcgen05 kind::tf32 with MN-major no-swizzle
// descriptors returned all-zero accumulators on this hardware/toolchain; the transposing producer avoids that path.)
// Each CTA owns a contiguous slice of rows, streams it through a 3-stage shared-memory ring in 16-row stages (TF32x3
// error-compensated split) and accumulates a full [<=256 x <=256] partial product in TMEM (2 M-tiles of 128 lanes).  Partials go to a workspace
// ([splits, m_pad, n_pad] fp32, coalesced) and a second kernel reduces them in double -- deterministic, no atomics.
//
// Bop = plain matrix, or the EdgeConv edge activation relu(P[centre] + Q[nbr]) gathered on the fly, optionally centred
// per column (Bop - mu) for the BatchNorm-backward moments (see nt_linear_bn_bwd).
#include "gemm_params.cuh"
#include "tc_common.cuh"

namespace nt {
using namespace tc;

constexpr int TN_PRODUCER_WARPS = 8;             // warps 0-7: producers + epilogue; warp 8: MMA issuer; warp 9: TMEM
constexpr int TN_THREADS = (TN_PRODUCER_WARPS + 2) * 32;
constexpr int TN_RB = 16;                       // rows (K of the MMA) per stage = 2 tf32 k-steps of 8
constexpr int TN_STAGES = 3;
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

// one 16-byte K-chunk (4 consecutive rows r of one output row m / n) of the hi and lo planes
__device__ __forceinline__ void tn_store_chunk(uint8_t *hi_plane, uint8_t *lo_plane, int chunk, int rows_pad, int row,
                                               const float (&v)[4]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
    const int off = (chunk * rows_pad + row) * 16;
    *reinterpret_cast<uint4 *>(hi_plane + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(lo_plane + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool EDGE>
__global__ void __launch_bounds__(TN_THREADS, 1) gemm_tn_tc_kernel(TNTCParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // plane = [4 K-chunks][rows_pad][16 B]  (K-major core matrices: LBO = rows_pad*16, SBO = 128)
    const size_t a_plane = (size_t)4 * p.m_pad * 16, b_plane = (size_t)4 * p.n_pad * 16;
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

    if (warp == TN_PRODUCER_WARPS && lane == 0) {
        for (int s = 0; s < TN_STAGES; ++s) { mbar_init(&full[s], TN_PRODUCER_WARPS * 32); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == TN_PRODUCER_WARPS + 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_split;
    const int64_t r_end = min(p.rows, r_begin + p.rows_per_split);
    const int n_stages = (int)((max((int64_t)0, r_end - r_begin) + TN_RB - 1) / TN_RB);

    if (warp < TN_PRODUCER_WARPS) {
        // ============ producers: (warp & 3) = K-chunk (4 rows), (warp >> 2) = which half of the 32-row groups, lane = row ============
        const int j = warp & 3;                                          // rows 4j .. 4j+3 of the 16-row stage
        const int half = warp >> 2;
        constexpr int MAXA = TN_MAX_M / 32 / 2, MAXB = TN_MAX_N / 32 / 2;  // 32-row groups per thread
        const int ia = p.m_pad / 32, ib = (p.n_pad + 31) / 32;
        if constexpr (EDGE) {
            struct Regs { float a[MAXA][4]; float b[MAXB][4]; float q[MAXB][4]; };   // raw loads; combined in consume()
            Regs v0, v1, v2;                                                 // prefetch ring, depth 3 (static addressing)

            auto fetch = [&](int st, Regs &v) {
                const int64_t rbase = r_begin + (int64_t)st * TN_RB + 4 * j;
                const float *arow[4], *bp[4], *bq[4];
    #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t r = rbase + i;
                    arow[i] = bp[i] = bq[i] = nullptr;
                    if (st < n_stages && r < r_end) {
                        arow[i] = p.a + r * p.lda;
                        if (p.b_edge) edge_row_ptrs(p.e, r, bp[i], bq[i]);
                        else bp[i] = p.b + r * p.ldb;
                    }
                }
    #pragma unroll
                for (int t = 0; t < MAXA; ++t) {
                    const int grp = half + 2 * t;
                    const int col = p.m0 + grp * 32 + lane;
    #pragma unroll
                    for (int i = 0; i < 4; ++i) v.a[t][i] = (grp < ia && arow[i] && col < p.m) ? __ldg(arow[i] + col) : 0.f;
                }
    #pragma unroll
                for (int t = 0; t < MAXB; ++t) {
                    const int grp = half + 2 * t;
                    const int col = p.n0 + grp * 32 + lane;
                    const bool c_ok = grp < ib && (grp * 32 + lane) < p.n_pad && col < p.n;
    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // invalid rows / columns must end up exactly 0 after relu(b + q) - mu: encode them as b = mu, q = 0
                        v.b[t][i] = (c_ok && bp[i]) ? __ldg(bp[i] + col) : 0.f;
                        v.q[t][i] = (c_ok && bp[i] && p.b_edge && bq[i]) ? __ldg(bq[i] + col) : 0.f;
                    }
                }
            };
            auto consume = [&](int st, Regs &v) {
                const int s = st % TN_STAGES, use = st / TN_STAGES;
                mbar_wait(&empty[s], (use & 1) ^ 1);
                uint8_t *a_hi = smem + s * stage_bytes, *a_lo = a_hi + a_plane, *b_hi = a_lo + a_plane, *b_lo = b_hi + b_plane;
    #pragma unroll
                for (int t = 0; t < MAXA; ++t) {
                    const int grp = half + 2 * t;
                    if (grp < ia) tn_store_chunk(a_hi, a_lo, j, p.m_pad, grp * 32 + lane, v.a[t]);
                }
    #pragma unroll
                for (int t = 0; t < MAXB; ++t) {
                    const int grp = half + 2 * t;
                    if (grp < ib && (grp * 32 + lane) < p.n_pad) {
                        const int col = p.n0 + grp * 32 + lane;
                        const bool c_ok = col < p.n;
                        const float mu = (c_ok && p.mu) ? __ldg(p.mu + col) : 0.f;
                        float x[4];
    #pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int64_t r = r_begin + (int64_t)st * TN_RB + 4 * j + i;
                            float y = v.b[t][i];
                            if (p.b_edge) y = fmaxf(y + v.q[t][i], 0.f);
                            x[i] = (c_ok && r < r_end) ? y - mu : 0.f;
                        }
                        tn_store_chunk(b_hi, b_lo, j, p.n_pad, grp * 32 + lane, x);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&full[s]);
            };
            fetch(0, v0); fetch(1, v1); fetch(2, v2);
            for (int st = 0; st < n_stages; st += 3) {
                consume(st, v0); fetch(st + 3, v0);
                if (st + 1 < n_stages) { consume(st + 1, v1); fetch(st + 4, v1); }
                if (st + 2 < n_stages) { consume(st + 2, v2); fetch(st + 5, v2); }
            }
        } else {
            // ---- plain operands: everything loop-invariant hoisted (column masks, centring values, base pointers); a stage
            // whose 16 rows all exist takes a path without per-row tests (ncu, round 1: two thirds of this kernel's issued
            // instructions were predicate / address arithmetic)
            struct Regs { float a[MAXA][4]; float b[MAXB][4]; };
            Regs v0, v1, v2;
            bool a_ld[MAXA], a_st[MAXA], b_ld[MAXB], b_st[MAXB];
            float muv[MAXB];
#pragma unroll
            for (int t = 0; t < MAXA; ++t) {
                const int grp = half + 2 * t;
                a_st[t] = grp < ia;
                a_ld[t] = a_st[t] && (p.m0 + grp * 32 + lane) < p.m;
            }
#pragma unroll
            for (int t = 0; t < MAXB; ++t) {
                const int grp = half + 2 * t;
                b_st[t] = grp < ib && (grp * 32 + lane) < p.n_pad;
                b_ld[t] = b_st[t] && (p.n0 + grp * 32 + lane) < p.n;
                muv[t] = (b_ld[t] && p.mu) ? __ldg(p.mu + p.n0 + grp * 32 + lane) : 0.f;
            }
            const float *abase = p.a + (r_begin + 4 * j) * (int64_t)p.lda + p.m0 + half * 32 + lane;
            const float *bbase = p.b + (r_begin + 4 * j) * (int64_t)p.ldb + p.n0 + half * 32 + lane;
            const int n_full = (int)(max((int64_t)0, r_end - r_begin) / TN_RB);       // stages with all 16 rows present
            auto fetch = [&](int st, Regs &v) {
                if (st >= n_stages) return;
                const float *ar = abase + (int64_t)st * TN_RB * p.lda;
                const float *br = bbase + (int64_t)st * TN_RB * p.ldb;
                if (st < n_full) {
#pragma unroll
                    for (int t = 0; t < MAXA; ++t)
#pragma unroll
                        for (int i = 0; i < 4; ++i) v.a[t][i] = a_ld[t] ? __ldg(ar + (int64_t)i * p.lda + t * 64) : 0.f;
#pragma unroll
                    for (int t = 0; t < MAXB; ++t)
#pragma unroll
                        for (int i = 0; i < 4; ++i) v.b[t][i] = b_ld[t] ? __ldg(br + (int64_t)i * p.ldb + t * 64) : muv[t];
                } else {
                    const int64_t rbase = r_begin + (int64_t)st * TN_RB + 4 * j;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const bool r_ok = rbase + i < r_end;
#pragma unroll
                        for (int t = 0; t < MAXA; ++t) v.a[t][i] = (a_ld[t] && r_ok) ? __ldg(ar + (int64_t)i * p.lda + t * 64) : 0.f;
#pragma unroll
                        for (int t = 0; t < MAXB; ++t) v.b[t][i] = (b_ld[t] && r_ok) ? __ldg(br + (int64_t)i * p.ldb + t * 64) : muv[t];
                    }
                }
            };
            auto consume = [&](int st, Regs &v) {
                const int s = st % TN_STAGES, use = st / TN_STAGES;
                mbar_wait(&empty[s], (use & 1) ^ 1);
                uint8_t *a_hi = smem + s * stage_bytes, *a_lo = a_hi + a_plane, *b_hi = a_lo + a_plane, *b_lo = b_hi + b_plane;
#pragma unroll
                for (int t = 0; t < MAXA; ++t)
                    if (a_st[t]) tn_store_chunk(a_hi, a_lo, j, p.m_pad, (half + 2 * t) * 32 + lane, v.a[t]);
#pragma unroll
                for (int t = 0; t < MAXB; ++t) {
                    if (b_st[t]) {
                        float x[4];                      // missing rows / columns were loaded as mu: exactly 0 after centring
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = v.b[t][i] - muv[t];
                        tn_store_chunk(b_hi, b_lo, j, p.n_pad, (half + 2 * t) * 32 + lane, x);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&full[s]);
            };
            fetch(0, v0); fetch(1, v1); fetch(2, v2);
            for (int st = 0; st < n_stages; st += 3) {
                consume(st, v0); fetch(st + 3, v0);
                if (st + 1 < n_stages) { consume(st + 1, v1); fetch(st + 4, v1); }
                if (st + 2 < n_stages) { consume(st + 2, v2); fetch(st + 5, v2); }
            }
        }

        // =========================== epilogue: TMEM -> coalesced partial tile ===========================
        float *tw = reinterpret_cast<float *>(smem) + warp * (32 * 33);          // stage memory is free now
        const int quad = warp & 3;                                               // TMEM lane quadrant this warp may read
        float *dst_tile = p.partial + (size_t)blockIdx.x * p.m_pad * p.n_pad;
        if (n_stages > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        for (int mt = 0; mt < m_tiles; ++mt) {
            for (int c0 = half * 32; c0 < p.n_pad; c0 += 64) {                   // the two halves interleave the column chunks
                float acc[32];
                if (n_stages > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * p.n_pad + c0), acc);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) tw[lane * 33 + i] = acc[i];
                __syncwarp();
                if (c0 + lane < p.n_pad) {
                    float *dst = dst_tile + (size_t)(mt * 128 + quad * 32) * p.n_pad + c0 + lane;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) dst[(size_t)rr * p.n_pad] = tw[rr * 33 + lane];
                }
                __syncwarp();
            }
        }
    } else if (warp == TN_PRODUCER_WARPS) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, (uint32_t)p.n_pad, 0, 0);        // K-major after the transposing producer
            const uint32_t lbo_a = (uint32_t)p.m_pad * 16, lbo_b = (uint32_t)p.n_pad * 16;
            for (int st = 0; st < n_stages; ++st) {
                const int s = st % TN_STAGES, use = st / TN_STAGES;
                mbar_wait(&full[s], use & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * stage_bytes), a_lo = a_hi + (uint32_t)a_plane;
                const uint32_t b_hi = a_lo + (uint32_t)a_plane, b_lo = b_hi + (uint32_t)b_plane;
#pragma unroll
                for (int ks = 0; ks < TN_RB / 8; ++ks) {                                  // one MMA = K 8 = 2 chunks
                    const uint64_t dbh = make_smem_desc(b_hi + ks * 2 * lbo_b, lbo_b, 128);
                    const uint64_t dbl = make_smem_desc(b_lo + ks * 2 * lbo_b, lbo_b, 128);
                    for (int mt = 0; mt < m_tiles; ++mt) {
                        const uint32_t moff = (uint32_t)mt * 128u * 16u;                 // 128 output rows further down the plane
                        const uint64_t dah = make_smem_desc(a_hi + moff + ks * 2 * lbo_a, lbo_a, 128);
                        const uint64_t dal = make_smem_desc(a_lo + moff + ks * 2 * lbo_a, lbo_a, 128);
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
    if (warp == TN_PRODUCER_WARPS + 1) tmem_dealloc(tmem_base, tmem_cols);
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
            const size_t stage_bytes = 2 * (size_t)4 * p.m_pad * 16 + 2 * (size_t)4 * p.n_pad * 16;
            const size_t smem = TN_STAGES * stage_bytes + 128;
            static bool configured = false;
            if (!configured) {
                cudaError_t err = cudaFuncSetAttribute(gemm_tn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (err == cudaSuccess)
                    err = cudaFuncSetAttribute(gemm_tn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (err != cudaSuccess) return fail("nt_gemm_tn(tc): cudaFuncSetAttribute: %s", cudaGetErrorString(err));
                configured = true;
            }
            if (smem > 227 * 1024) return fail("nt_gemm_tn(tc): tile does not fit shared memory%s", "");
            if (b_edge) gemm_tn_tc_kernel<true><<<splits, TN_THREADS, smem, st>>>(p);
            else gemm_tn_tc_kernel<false><<<splits, TN_THREADS, smem, st>>>(p);
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
