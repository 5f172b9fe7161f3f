// C-ABI entry points of the fused row GEMMs for the MLP() blocks (reference nn/net_blocks.py:43-47):
//
//   nt_gemm_nt : out = epilogue( producer(A) . W^T )      rows x K  times  n_out x K      (gemm_tc.cu / gemm_tc3.cu, tcgen05)
//   nt_gemm_tn : out += A^T . producer(B)                 weight gradients                 (gemm_tn_tc.cu, tcgen05)
//
// Producers build the A (or B) operand directly in shared memory -- a plain matrix, or the EdgeConv edge activation
// relu(P[centre] + Q[neighbour]) gathered through the kNN index (DynamicEdgeConv.message, nn/net_blocks.py:127-135;
// the first Linear of the edge MLP is split as W.[x_i, x_j - x_i] = (W_a - W_b) x_i + W_b x_j so it is evaluated per
// POINT, and no [E, 2C] edge tensor is ever materialised).  Epilogues fuse bias, ReLU, BatchNorm statistics,
// the max/min aggregation over the k neighbours, and the BN+ReLU backward.
// (Round 1 also shipped an fp32 CUDA-core engine for validation; the float64 comparisons in tests/ replaced it.)
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include "gemm_params.cuh"

namespace nt {
constexpr int G_TM = 128;   // rows per CTA tile
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
    p.engine = g->engine;
    NT_REQUIRE(g->engine == 0 || g->engine == 1 || (g->engine >= 3 && g->engine <= 6), "nt_gemm_nt: engine must be 0 (auto), 1, 3, 4, 5 or 6");
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
    NT_REQUIRE(g->w_split != nullptr, "nt_gemm_nt: w_split missing (prepare the weights with nt_gemm_prepare_weights; there is no "
                                        "CUDA-core or CPU fallback)");
    switch (g->epilogue) {
        case NT_EPI_BIAS:
            NT_REQUIRE(g->out && g->ldo >= g->n_out, "nt_gemm_nt: bad output");
            NT_REQUIRE(g->producer == NT_PROD_PLAIN, "nt_gemm_nt: NT_EPI_BIAS needs the plain producer");
            return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
        case NT_EPI_RELU_STATS:
            NT_REQUIRE(g->out == nullptr || g->ldo >= g->n_out, "nt_gemm_nt: bad output");
            return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
        case NT_EPI_RELU_MAXMIN:
            NT_REQUIRE(g->k >= 1 && g->k <= G_TM, "nt_gemm_nt: aggregation needs 1 <= k <= 128");
            NT_REQUIRE(g->rows % g->k == 0, "nt_gemm_nt: rows must be a multiple of k");
            NT_REQUIRE(g->vmax && g->vmin && g->imax && g->imin, "nt_gemm_nt: aggregation outputs missing");
            p.rows_per_tile = (G_TM / g->k) * g->k;
            return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
        case NT_EPI_BNRELU_BWD:
            NT_REQUIRE((g->out ? g->ldo >= g->n_out : p.scatter != nullptr) && g->k0 && g->k1 && g->mu, "nt_gemm_nt: bwd operands missing");
            NT_REQUIRE(g->aux_edge ? (g->pq != nullptr) : (g->aux != nullptr), "nt_gemm_nt: aux operand missing");
            NT_REQUIRE(g->producer == NT_PROD_PLAIN, "nt_gemm_nt: NT_EPI_BNRELU_BWD needs the plain producer");
            return launch_nt_tc(p, g->producer, g->epilogue, g->precision, g->w_split, st);
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
    NT_REQUIRE(workspace != nullptr, "nt_gemm_tn: workspace missing (nt_gemm_tn_workspace_bytes() bytes; there is no CUDA-core fallback)");
    const EdgeSrc e{pq, ldpq, qoff, idx, k > 0 ? k : 1, n_per_cloud > 0 ? n_per_cloud : 1};
    // plain operands: MN-major BF16x3 engine (rows are read as they lie); NT_TN_PRECISION=tf32x3 keeps the transposing TF32x3 engine
    static const bool force_tf32 = [] { const char *v = getenv("NT_TN_PRECISION"); return v && strcmp(v, "tf32x3") == 0; }();
    if (!pq && !force_tf32)
        return gemm_tn_mn(a, lda, m, b, ldb, n, rows, mu, out, out_double, ldo, reinterpret_cast<float *>(workspace),
                          reinterpret_cast<cudaStream_t>(stream));
    return gemm_tn_tc(a, lda, m, b, ldb, n, rows, e, pq != nullptr, mu, out, out_double, ldo, reinterpret_cast<float *>(workspace),
                      reinterpret_cast<cudaStream_t>(stream));
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
