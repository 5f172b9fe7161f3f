// Fused row GEMMs for the MLP() blocks (reference nn/net_blocks.py:43-47) -- fp32 CUDA-core tile engine.
//
//   nt_gemm_nt : out = epilogue( producer(A) . W^T )      rows x K  times  n_out x K
//   nt_gemm_tn : out += A^T . producer(B)                 weight gradients (split over rows, fp32 atomics)
//
// Producers build the A (or B) tile directly in shared memory -- a plain matrix, or the EdgeConv edge activation
// relu(P[centre] + Q[neighbour]) gathered through the kNN index (DynamicEdgeConv.message, nn/net_blocks.py:127-135;
// the first Linear of the edge MLP is split as W.[x_i, x_j - x_i] = (W_a - W_b) x_i + W_b x_j so it is evaluated per
// POINT, and no [E, 2C] edge tensor is ever materialised).  Epilogues fuse bias, ReLU, BatchNorm statistics,
// the max/min aggregation over the k neighbours, and the BN+ReLU backward.
#include "common.cuh"
#include "gemm_params.cuh"

namespace nt {

constexpr int G_THREADS = 256;
constexpr int G_TM = 128;   // rows per CTA tile
constexpr int G_TN = 64;    // output columns per CTA tile
constexpr int G_TK = 16;    // inner step

template <int PROD, int EPI>
__global__ void __launch_bounds__(G_THREADS) gemm_nt_kernel(NTParams p) {
    // operand tiles and the aggregation tile share one pool (the latter is only live after the main loop)
    constexpr bool kNeedTile = (EPI == NT_EPI_RELU_MAXMIN);
    constexpr int kOperandFloats = G_TK * (G_TM + 4) + G_TK * (G_TN + 4);
    constexpr int kTileFloats = kNeedTile ? G_TM * (G_TN + 1) : 0;
    __shared__ __align__(16) float pool[kOperandFloats > kTileFloats ? kOperandFloats : kTileFloats];
    __shared__ float red[2][16][G_TN];                        // column partials (sum, sumsq) per row group
    float (*As)[G_TM + 4] = reinterpret_cast<float (*)[G_TM + 4]>(pool);
    float (*Bs)[G_TN + 4] = reinterpret_cast<float (*)[G_TN + 4]>(pool + G_TK * (G_TM + 4));
    float (*vt)[G_TN + 1] = reinterpret_cast<float (*)[G_TN + 1]>(pool);

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * p.rows_per_tile;
    const int rows_here = (int)min((int64_t)p.rows_per_tile, p.rows - row0);
    const int col0 = blockIdx.y * G_TN;

    // ---- A-tile loader assignment: thread -> (row, 8 consecutive k)
    const int ar = tid >> 1, ak = (tid & 1) * 8;
    const bool ar_ok = ar < rows_here;
    const float *ap = nullptr, *aq = nullptr;
    if (ar_ok) {
        if (PROD == NT_PROD_PLAIN) ap = p.a + (row0 + ar) * (int64_t)p.lda;
        else edge_row_ptrs(p.e, row0 + ar, ap, aq);
    }
    // ---- B-tile loader assignment: thread -> (n, 4 consecutive k)
    const int bn = tid >> 2, bk = (tid & 3) * 4;
    const bool bn_ok = (col0 + bn) < p.n_out;
    const float *wp = p.w + (int64_t)(bn_ok ? col0 + bn : 0) * p.ldw;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += G_TK) {
        float av[8], bv[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int kk = k0 + ak + i;
            float v = 0.f;
            if (ar_ok && kk < p.K) {
                if (PROD == NT_PROD_PLAIN) v = __ldg(ap + kk);
                else {
                    v = __ldg(ap + kk);
                    if (aq) v += __ldg(aq + kk);
                    v = fmaxf(v, 0.f);
                }
            }
            av[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kk = k0 + bk + i;
            bv[i] = (bn_ok && kk < p.K) ? __ldg(wp + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) As[ak + i][ar] = av[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[bk + i][bn] = bv[i];
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < G_TK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    // ---- epilogue
    if (kNeedTile) __syncthreads();      // the aggregation tile aliases the operand tiles
    float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = col0 + tx * 4 + j;
        const bool c_ok = c < p.n_out;
        const float bias = (c_ok && p.bias && EPI != NT_EPI_BNRELU_BWD) ? __ldg(p.bias + c) : 0.f;
        float kk0 = 0.f, kk1 = 0.f, mu = 0.f;
        if (EPI == NT_EPI_BNRELU_BWD && c_ok) { kk0 = __ldg(p.k0 + c); kk1 = __ldg(p.k1 + c); mu = __ldg(p.mu + c); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = ty * 8 + i;
            const bool ok = c_ok && r < rows_here;
            float v = acc[i][j] + bias;
            if (EPI == NT_EPI_BIAS) {
                if (ok) p.out[(row0 + r) * (int64_t)p.ldo + c] = v;
            } else if (EPI == NT_EPI_RELU_STATS || EPI == NT_EPI_RELU_MAXMIN) {
                v = fmaxf(v, 0.f);
                if (ok) {
                    csum[j] += v; csq[j] = fmaf(v, v, csq[j]);
                    if (p.out) p.out[(row0 + r) * (int64_t)p.ldo + c] = v;
                }
                if (kNeedTile) vt[r][tx * 4 + j] = v;
            } else {  // NT_EPI_BNRELU_BWD
                if (ok) {
                    float a;
                    if (p.aux_edge) {
                        const float *pp, *qq;
                        edge_row_ptrs(p.ae, row0 + r, pp, qq);
                        a = __ldg(pp + c);
                        if (qq) a += __ldg(qq + c);
                        a = fmaxf(a, 0.f);
                    } else {
                        a = p.aux[(row0 + r) * (int64_t)p.ldaux + c];
                    }
                    float dz = (a > 0.f) ? (acc[i][j] - kk0 - (a - mu) * kk1) : 0.f;
                    p.out[(row0 + r) * (int64_t)p.ldo + c] = dz;
                    csum[j] += dz;
                }
            }
        }
    }

    if (EPI == NT_EPI_RELU_STATS || EPI == NT_EPI_RELU_MAXMIN || EPI == NT_EPI_BNRELU_BWD) {
        const bool want = (EPI == NT_EPI_BNRELU_BWD) ? (p.colsum != nullptr) : (p.stats != nullptr);
        if (want) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { red[0][ty][tx * 4 + j] = csum[j]; red[1][ty][tx * 4 + j] = csq[j]; }
            __syncthreads();
            if (tid < 2 * G_TN) {
                const int which = tid / G_TN, cc = tid % G_TN;
                float s = 0.f;
#pragma unroll
                for (int g = 0; g < 16; ++g) s += red[which][g][cc];
                const int c = col0 + cc;
                if (c < p.n_out) {
                    if (EPI == NT_EPI_BNRELU_BWD) { if (which == 0) atomicAdd(p.colsum + c, (double)s); }
                    else atomicAdd(p.stats + which * p.n_out + c, (double)s);
                }
            }
        }
    }

    if (EPI == NT_EPI_RELU_MAXMIN) {
        __syncthreads();
        const int kk = p.k_agg;
        const int nodes_here = rows_here / kk;
        const int64_t node0 = row0 / kk;
        for (int t = tid; t < nodes_here * G_TN; t += G_THREADS) {
            const int nd = t / G_TN, cc = t % G_TN;
            const int c = col0 + cc;
            if (c >= p.n_out) continue;
            float mx = vt[nd * kk][cc], mn = mx;
            int ix = 0, in = 0;
            for (int s = 1; s < kk; ++s) {
                float v = vt[nd * kk + s][cc];
                if (v > mx) { mx = v; ix = s; }
                if (v < mn) { mn = v; in = s; }
            }
            const int64_t o = (node0 + nd) * (int64_t)p.n_out + c;
            p.vmax[o] = mx; p.vmin[o] = mn; p.imax[o] = (uint8_t)ix; p.imin[o] = (uint8_t)in;
        }
    }
}

template <int PROD, int EPI>
static int launch_nt(const NTParams &p, cudaStream_t st) {
    dim3 grid((unsigned)((p.rows + p.rows_per_tile - 1) / p.rows_per_tile), (p.n_out + G_TN - 1) / G_TN);
    gemm_nt_kernel<PROD, EPI><<<grid, G_THREADS, 0, st>>>(p);
    return check_launch("nt_gemm_nt");
}

// ------------------------------------------------------------------------------------------------------------
// Weight-gradient GEMM: out[m, n] += sum_r A[r, m] * B[r, n]
// ------------------------------------------------------------------------------------------------------------
constexpr int T_TM = 64, T_TN = 64, T_TK = 16;

struct TNParams {
    const float *a; int lda; int m;
    const float *b; int ldb; int n;
    int64_t rows; int64_t rows_per_split;
    EdgeSrc e; int b_edge;
    const float *mu;          // optional per-column centring of the B operand (BatchNorm mean)
    void *out; int ldo;
};

// OutT = float: fp32 atomics into out.  OutT = double: every CTA's fp32 partial (at most rows_per_split rows) is
// accumulated across CTAs in double -- used for the BN-backward statistics, which are differences of large sums.
template <typename OutT>
__global__ void __launch_bounds__(G_THREADS) gemm_tn_kernel(TNParams p) {
    __shared__ __align__(16) float As[T_TK][T_TM + 4];
    __shared__ __align__(16) float Bs[T_TK][T_TN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * T_TM, n0 = blockIdx.y * T_TN;
    const int64_t r_begin = (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t r_end = min(p.rows, r_begin + p.rows_per_split);
    const int lr = tid >> 4;           // local row 0..15
    const int lc = (tid & 15) * 4;     // 4 consecutive columns

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t r0 = r_begin; r0 < r_end; r0 += T_TK) {
        const int64_t r = r0 + lr;
        const bool r_ok = r < r_end;
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int mm = m0 + lc + i;
            av[i] = (r_ok && mm < p.m) ? __ldg(p.a + r * p.lda + mm) : 0.f;
        }
        if (p.b_edge) {
            const float *pp = nullptr, *qq = nullptr;
            if (r_ok) edge_row_ptrs(p.e, r, pp, qq);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nn = n0 + lc + i;
                float v = 0.f;
                if (r_ok && nn < p.n) {
                    v = __ldg(pp + nn);
                    if (qq) v += __ldg(qq + nn);
                    v = fmaxf(v, 0.f);
                    if (p.mu) v -= __ldg(p.mu + nn);
                }
                bv[i] = v;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nn = n0 + lc + i;
                float v = 0.f;
                if (r_ok && nn < p.n) {
                    v = __ldg(p.b + r * p.ldb + nn);
                    if (p.mu) v -= __ldg(p.mu + nn);
                }
                bv[i] = v;
            }
        }
        __syncthreads();
        *reinterpret_cast<float4 *>(&As[lr][lc]) = make_float4(av[0], av[1], av[2], av[3]);
        *reinterpret_cast<float4 *>(&Bs[lr][lc]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < T_TK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float a[4] = {a0.x, a0.y, a0.z, a0.w};
            float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mm = m0 + ty * 4 + i;
        if (mm >= p.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn < p.n) atomicAdd(reinterpret_cast<OutT *>(p.out) + (int64_t)mm * p.ldo + nn, (OutT)acc[i][j]);
        }
    }
}

}  // namespace nt

static int nt_fill_params(const nt_gemm_args *g, nt::NTParams &p) {
    using namespace nt;
    NT_REQUIRE(g, "nt_gemm_nt: null args");
    NT_REQUIRE(g->rows >= 0 && g->K >= 1 && g->n_out >= 1, "nt_gemm_nt: bad shape");
    NT_REQUIRE(g->w && g->ldw >= g->K, "nt_gemm_nt: bad weight");
    p = NTParams{};
    p.rows = g->rows; p.K = g->K; p.n_out = g->n_out; p.rows_per_tile = G_TM;
    p.a = g->a; p.lda = g->lda;
    p.e = EdgeSrc{g->pq, g->ldpq, g->qoff, g->idx, g->k > 0 ? g->k : 1, g->n_per_cloud > 0 ? g->n_per_cloud : 1};
    p.w = g->w; p.ldw = g->ldw; p.bias = g->bias;
    p.out = g->out; p.ldo = g->ldo; p.stats = g->stats;
    p.vmax = g->vmax; p.vmin = g->vmin; p.imax = g->imax; p.imin = g->imin; p.k_agg = g->k;
    p.aux = g->aux; p.ldaux = g->ldaux; p.aux_edge = g->aux_edge; p.ae = p.e;
    p.k0 = g->k0; p.k1 = g->k1; p.mu = g->mu; p.colsum = g->colsum;
    p.scatter = g->scatter_dpq; p.ldscatter = g->ldscatter;
    if (g->producer == NT_PROD_PLAIN) NT_REQUIRE(g->a && g->lda >= g->K, "nt_gemm_nt: bad plain operand");
    else if (g->producer == NT_PROD_EDGE) {
        NT_REQUIRE(g->pq && g->ldpq >= g->K, "nt_gemm_nt: bad edge operand");
        if (g->idx) NT_REQUIRE(g->k >= 1 && g->n_per_cloud >= 1, "nt_gemm_nt: edge operand needs k and n_per_cloud");
    } else return fail("nt_gemm_nt: unknown producer %s%ld", "", g->producer);
    if (p.scatter) {
        NT_REQUIRE(g->epilogue == NT_EPI_BNRELU_BWD, "nt_gemm_nt: scatter_dpq needs NT_EPI_BNRELU_BWD");
        NT_REQUIRE(g->idx && g->k >= 1 && g->k <= G_TM && g->n_per_cloud >= 1 && g->rows % g->k == 0,
                   "nt_gemm_nt: scatter_dpq needs idx, k, n_per_cloud and rows % k == 0");
        p.rows_per_tile = (G_TM / g->k) * g->k;          // tiles hold whole centre points
    }
    return 0;
}

extern "C" int nt_gemm_nt_scatter_supported(const nt_gemm_args *g) {
    using namespace nt;
    NTParams p;
    if (!g || nt_fill_params(g, p) != 0 || !p.scatter || g->w_split == nullptr) return 0;
    return nt_tc_would_stream(p, g->producer, g->epilogue, g->precision) ? 1 : 0;
}

extern "C" int nt_gemm_nt(const nt_gemm_args *g, void *stream) {
    using namespace nt;
    NTParams p;
    if (g && g->rows == 0 && g->K >= 1 && g->n_out >= 1) return 0;       // nothing to do (operand pointers may be NULL)
    if (int rc = nt_fill_params(g, p)) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool tc = g->w_split != nullptr;
    switch (g->epilogue) {
        case NT_EPI_BIAS:
            NT_REQUIRE(g->out && g->ldo >= g->n_out, "nt_gemm_nt: bad output");
            NT_REQUIRE(g->producer == NT_PROD_PLAIN, "nt_gemm_nt: NT_EPI_BIAS needs the plain producer");
            if (tc) return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
            return launch_nt<NT_PROD_PLAIN, NT_EPI_BIAS>(p, st);
        case NT_EPI_RELU_STATS:
            NT_REQUIRE(g->out == nullptr || g->ldo >= g->n_out, "nt_gemm_nt: bad output");
            if (tc) return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
            return g->producer == NT_PROD_PLAIN ? launch_nt<NT_PROD_PLAIN, NT_EPI_RELU_STATS>(p, st)
                                                : launch_nt<NT_PROD_EDGE, NT_EPI_RELU_STATS>(p, st);
        case NT_EPI_RELU_MAXMIN:
            NT_REQUIRE(g->k >= 1 && g->k <= G_TM, "nt_gemm_nt: aggregation needs 1 <= k <= 128");
            NT_REQUIRE(g->rows % g->k == 0, "nt_gemm_nt: rows must be a multiple of k");
            NT_REQUIRE(g->vmax && g->vmin && g->imax && g->imin, "nt_gemm_nt: aggregation outputs missing");
            p.rows_per_tile = (G_TM / g->k) * g->k;
            if (tc) return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
            return g->producer == NT_PROD_PLAIN ? launch_nt<NT_PROD_PLAIN, NT_EPI_RELU_MAXMIN>(p, st)
                                                : launch_nt<NT_PROD_EDGE, NT_EPI_RELU_MAXMIN>(p, st);
        case NT_EPI_BNRELU_BWD:
            NT_REQUIRE((g->out ? g->ldo >= g->n_out : p.scatter != nullptr) && g->k0 && g->k1 && g->mu, "nt_gemm_nt: bwd operands missing");
            NT_REQUIRE(!p.scatter || tc, "nt_gemm_nt: scatter_dpq needs the tensor-core engine");
            NT_REQUIRE(g->aux_edge ? (g->pq != nullptr) : (g->aux != nullptr), "nt_gemm_nt: aux operand missing");
            NT_REQUIRE(g->producer == NT_PROD_PLAIN, "nt_gemm_nt: NT_EPI_BNRELU_BWD needs the plain producer");
            if (tc) return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
            return launch_nt<NT_PROD_PLAIN, NT_EPI_BNRELU_BWD>(p, st);
        default: return fail("nt_gemm_nt: unknown epilogue %s%ld", "", g->epilogue);
    }
}

static int gemm_tn_impl(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows,
                        const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                        const float *mu, void *out, int out_double, int ldo, void *workspace, void *stream) {
    using namespace nt;
    NT_REQUIRE(a && out && m >= 1 && n >= 1 && lda >= m && ldo >= n && rows >= 0, "nt_gemm_tn: bad arguments");
    NT_REQUIRE(pq ? (ldpq >= n) : (b != nullptr && ldb >= n), "nt_gemm_tn: bad B operand");
    if (rows == 0) return 0;
    TNParams p{};
    p.a = a; p.lda = lda; p.m = m; p.b = b; p.ldb = ldb; p.n = n; p.rows = rows;
    p.e = EdgeSrc{pq, ldpq, qoff, idx, k > 0 ? k : 1, n_per_cloud > 0 ? n_per_cloud : 1};
    p.b_edge = pq != nullptr; p.out = out; p.ldo = ldo; p.mu = mu;
    if (workspace)      // tensor-core engine (the shipped host code always passes a workspace)
        return gemm_tn_tc(a, lda, m, b, ldb, n, rows, p.e, p.b_edge, mu, out, out_double, ldo,
                          reinterpret_cast<float *>(workspace), reinterpret_cast<cudaStream_t>(stream));
    const int tiles = ((m + T_TM - 1) / T_TM) * ((n + T_TN - 1) / T_TN);
    int64_t splits = (4 * 148 + tiles - 1) / tiles;                       // ~4 CTAs per SM in flight
    if (out_double) {                                                     // short fp32 partials for the statistics
        int64_t want = (rows + 1023) / 1024;
        if (want > splits) splits = want;
    }
    int64_t max_splits = (rows + 4 * T_TK - 1) / (4 * T_TK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = (rows + splits - 1) / splits;
    rps = ((rps + T_TK - 1) / T_TK) * T_TK;
    splits = (rows + rps - 1) / rps;
    p.rows_per_split = rps;
    dim3 grid((m + T_TM - 1) / T_TM, (n + T_TN - 1) / T_TN, (unsigned)splits);
    if (out_double) gemm_tn_kernel<double><<<grid, G_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    else gemm_tn_kernel<float><<<grid, G_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("nt_gemm_tn");
}

extern "C" int nt_gemm_tn(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows,
                          const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                          float *out, int ldo, void *workspace, void *stream) {
    return gemm_tn_impl(a, lda, m, b, ldb, n, rows, pq, ldpq, qoff, idx, k, n_per_cloud, nullptr, out, 0, ldo, workspace,
                        stream);
}

extern "C" int nt_gemm_tn_centered(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows,
                                   const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                                   const float *mu, double *out, int ldo, void *workspace, void *stream) {
    NT_REQUIRE(mu != nullptr, "nt_gemm_tn_centered: mu missing");
    return gemm_tn_impl(a, lda, m, b, ldb, n, rows, pq, ldpq, qoff, idx, k, n_per_cloud, mu, out, 1, ldo, workspace, stream);
}
