// Composite EdgeConv layer, TRAINING mode, behind the C ABI (SURVEY 8b "nt_edgeconv_fwd/bwd ... training variants with BN stat
// buffers"; reference nn/net_blocks.py:126-135,172-180 -> torch_geometric DynamicEdgeConv over MLP([2C, H1, H2, H3]), each stage
// Linear -> ReLU -> BatchNorm1d in batch-statistics mode, max aggregation over the k neighbours, optional skip-connection columns):
//
//   out[i] = BN3( max_{j in kNN(i)}  a3 ),  a3 = relu(W3' a2 + b3'), a2 = relu(W2' a1 + b2'), a1 = relu(P[i] + Q[j])
//
// One call = the whole layer: the host orchestration that used to live in the Python autograd function (ops.py) is here, so a
// non-Python host gets "EdgeConv forward / backward" as two calls.  The kernels are the library's own entry points, launched in the
// same order on the caller's stream: nt_gemm_nt (PQ) -> nt_edge_activation -> nt_bn_fold -> nt_gemm_nt(RELU_STATS) -> nt_bn_fold ->
// nt_gemm_nt(RELU_MAXMIN) -> nt_bn_fold -> nt_maxmin_finish, and the mirror image for the backward (see DESIGN.md section 2 for the
// algebra: split first Linear, BatchNorm folded into the next Linear, max/min commuted past the trailing BN, BN backward moments
// out of the weight-gradient GEMM).  The caller owns every byte: `saved` (what the backward needs) and `scratch` (temporaries) are
// sized by the two query functions; nothing is allocated, no hidden synchronisation, no state.
#include "common.cuh"

namespace nt {
namespace {

inline int pad4(int c) { return (c + 3) & ~3; }
inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

// byte offsets of the segments of `saved` / `scratch` (every segment 256-byte aligned)
struct Plan {
    int64_t M, R;
    int C, H[3], k, N, tail;
    int ldpq, ld1, ld2, ld3;
    // saved
    size_t pq, a1, a2, a3, sel, vsel, bn[3], wft[3], wc, saved_bytes;
    // forward scratch
    size_t f_bc, f_stats, f_wf, f_bf, f_split[3], f_vmax, f_vmin, f_imax, f_imin, fwd_bytes;
    // backward scratch
    size_t b_acc, b_dz3, b_dz2, b_dz1, b_vec[3], b_dpq, b_dwc, b_wct, b_split[3], b_tn, bwd_bytes;
    size_t acc_doubles;
};

bool make_plan(const nt_edgeconv_args *g, Plan &p) {
    if (!g || g->M < 0 || g->C < 1 || g->H1 < 1 || g->H2 < 1 || g->H3 < 1 || g->k < 1 || g->k > 128 || g->n_per_cloud < 1 || g->tail < 0)
        return false;
    p.M = g->M; p.k = g->k; p.N = g->n_per_cloud; p.R = g->M * g->k; p.C = g->C; p.tail = g->tail;
    p.H[0] = g->H1; p.H[1] = g->H2; p.H[2] = g->H3;
    p.ldpq = pad4(2 * g->H1); p.ld1 = pad4(g->H1); p.ld2 = pad4(g->H2); p.ld3 = pad4(g->H3);
    size_t o = 0;
    auto seg = [&](size_t bytes) { size_t at = o; o += up256(bytes); return at; };
    p.pq = seg((size_t)p.M * p.ldpq * 4);
    p.a1 = seg((size_t)p.R * p.ld1 * 4);
    p.a2 = seg((size_t)p.R * p.ld2 * 4);
    p.a3 = seg((size_t)p.R * p.ld3 * 4);
    p.sel = seg((size_t)p.M * g->H3);
    p.vsel = seg((size_t)p.M * g->H3 * 4);
    for (int l = 0; l < 3; ++l) p.bn[l] = seg((size_t)4 * p.H[l] * 4);                 // mean | rstd | s | t
    p.wft[0] = 0;
    for (int l = 1; l < 3; ++l) p.wft[l] = seg((size_t)p.H[l - 1] * p.H[l] * 4);       // (W_l diag(s_{l-1}))^T : [H_{l-1}, H_l]
    p.wc = seg((size_t)2 * g->H1 * g->C * 4);                                          // [Wa - Wb ; Wb] : [2 H1, C]
    p.saved_bytes = o;

    o = 0;
    p.f_bc = seg((size_t)2 * g->H1 * 4);
    p.f_stats = seg((size_t)2 * (g->H1 + g->H2 + g->H3) * 8);
    const int wmax = g->H2 * g->H1 > g->H3 * g->H2 ? g->H2 * g->H1 : g->H3 * g->H2;
    p.f_wf = seg((size_t)wmax * 4);
    p.f_bf = seg((size_t)(g->H2 > g->H3 ? g->H2 : g->H3) * 4);
    p.f_split[0] = seg((size_t)nt_gemm_weights_bytes(2 * g->H1, g->C, NT_PREC_TF32X3));
    p.f_split[1] = seg((size_t)nt_gemm_weights_bytes(g->H2, g->H1, NT_PREC_TF32X3));
    p.f_split[2] = seg((size_t)nt_gemm_weights_bytes(g->H3, g->H2, NT_PREC_TF32X3));
    p.f_vmax = seg((size_t)p.M * g->H3 * 4);
    p.f_vmin = seg((size_t)p.M * g->H3 * 4);
    p.f_imax = seg((size_t)p.M * g->H3);
    p.f_imin = seg((size_t)p.M * g->H3);
    p.fwd_bytes = o;

    o = 0;
    p.acc_doubles = (size_t)2 * g->H3 + g->H3 + (size_t)g->H3 * g->H2 + g->H2 + (size_t)g->H2 * g->H1 + g->H1;
    p.b_acc = seg(p.acc_doubles * 8);
    // dz3 is dead once dz2 exists, and dz1 is written after that: the two share one region (1.6 GB at C4: 2.56 M edge rows)
    p.b_dz3 = seg((size_t)p.R * (p.ld3 > p.ld1 ? p.ld3 : p.ld1) * 4);
    p.b_dz2 = seg((size_t)p.R * p.ld2 * 4);
    p.b_dz1 = p.b_dz3;
    p.b_vec[0] = 0;
    for (int l = 1; l < 3; ++l) p.b_vec[l] = seg((size_t)2 * p.H[l - 1] * 4);          // k0 | k1 of the BN in front of Linear l
    p.b_dpq = seg((size_t)p.M * 2 * g->H1 * 4);
    p.b_dwc = seg((size_t)2 * g->H1 * g->C * 4);
    p.b_wct = seg((size_t)2 * g->H1 * g->C * 4);
    p.b_split[0] = seg((size_t)nt_gemm_weights_bytes(g->C, 2 * g->H1, NT_PREC_TF32X3));
    p.b_split[1] = seg((size_t)nt_gemm_weights_bytes(g->H1, g->H2, NT_PREC_TF32X3));
    p.b_split[2] = seg((size_t)nt_gemm_weights_bytes(g->H2, g->H3, NT_PREC_TF32X3));
    p.b_tn = seg((size_t)nt_gemm_tn_workspace_bytes());
    p.bwd_bytes = o;
    return true;
}

// Wc = [W0[:, :C] - W0[:, C:] ; W0[:, C:]]  ([2 H1, C]),  bc = [b0 ; 0]      (W.[x_i, x_j - x_i] = (Wa - Wb) x_i + Wb x_j)
__global__ void compose_wc_kernel(const float *__restrict__ w0, const float *__restrict__ b0, int H1, int C, float *__restrict__ wc,
                                  float *__restrict__ bc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H1 * C) {
        const int o = i / C, c = i - o * C;
        const float wa = w0[(size_t)o * 2 * C + c], wb = w0[(size_t)o * 2 * C + C + c];
        wc[i] = wa - wb;
        wc[(size_t)H1 * C + i] = wb;
    }
    if (i < 2 * H1) bc[i] = i < H1 ? (b0 ? b0[i] : 0.f) : 0.f;
}

__global__ void transpose_kernel(const float *__restrict__ src, int rows, int cols, float *__restrict__ dst) {   // dst [cols, rows]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i - r * cols;
    dst[(size_t)c * rows + r] = src[i];
}

// gradients of the first Linear from the per-point gradient of [Wa - Wb ; Wb]:  dW0 = [dWc_top | dWc_bottom - dWc_top]  ([H1, 2C]);
// plus the double -> float copies of the bias / trailing-BN gradients
__global__ void finish_first_kernel(const float *__restrict__ dwc, int H1, int C, float *__restrict__ gw0, const double *__restrict__ csum1,
                                    float *__restrict__ gb0, const double *__restrict__ sums3, int H3, float *__restrict__ gbeta3,
                                    float *__restrict__ ggamma3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H1 * C) {
        const int o = i / C, c = i - o * C;
        const float top = dwc[i], bot = dwc[(size_t)H1 * C + i];
        gw0[(size_t)o * 2 * C + c] = top;
        gw0[(size_t)o * 2 * C + C + c] = bot - top;
    }
    if (i < H1 && gb0) gb0[i] = (float)csum1[i];
    if (i < H3) {
        if (gbeta3) gbeta3[i] = (float)sums3[i];
        if (ggamma3) ggamma3[i] = (float)sums3[H3 + i];
    }
}

inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

#define NT_TRY(expr)            \
    do {                        \
        int rc_ = (expr);       \
        if (rc_) return rc_;    \
    } while (0)

int check_common(const nt_edgeconv_args *g) {
    NT_REQUIRE(g->x && g->idx && g->ldx >= g->C, "nt_edgeconv: bad input");
    for (int l = 0; l < 3; ++l) NT_REQUIRE(g->W[l] && g->gamma[l] && g->beta[l], "nt_edgeconv: parameters missing");
    NT_REQUIRE(g->saved && g->scratch, "nt_edgeconv: saved / scratch buffers missing (nt_edgeconv_saved_bytes / _scratch_bytes)");
    NT_REQUIRE((reinterpret_cast<uintptr_t>(g->saved) & 255u) == 0 && (reinterpret_cast<uintptr_t>(g->scratch) & 255u) == 0,
               "nt_edgeconv: saved / scratch must be 256-byte aligned");
    NT_REQUIRE(g->tail == 0 || (g->tail_src && g->tail_ld >= g->tail), "nt_edgeconv: bad skip-connection source");
    return 0;
}

}  // namespace
}  // namespace nt

using namespace nt;

extern "C" int64_t nt_edgeconv_saved_bytes(const nt_edgeconv_args *g) {
    Plan p;
    return make_plan(g, p) ? (int64_t)p.saved_bytes : -1;
}

extern "C" int64_t nt_edgeconv_scratch_bytes(const nt_edgeconv_args *g, int backward) {
    Plan p;
    if (!make_plan(g, p)) return -1;
    return (int64_t)(backward ? p.bwd_bytes : p.fwd_bytes);
}

extern "C" int nt_edgeconv_train_fwd(const nt_edgeconv_args *g, void *stream) {
    Plan p;
    NT_REQUIRE(make_plan(g, p), "nt_edgeconv_train_fwd: bad shape");
    if (g->M == 0) return 0;
    NT_TRY(check_common(g));
    NT_REQUIRE(g->out && g->ldo >= g->H3 + g->tail, "nt_edgeconv_train_fwd: bad output");
    NT_REQUIRE(g->M % g->n_per_cloud == 0, "nt_edgeconv_train_fwd: M must be a multiple of n_per_cloud");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t *sv = reinterpret_cast<uint8_t *>(g->saved), *sc = reinterpret_cast<uint8_t *>(g->scratch);
    auto F = [](uint8_t *base, size_t off) { return reinterpret_cast<float *>(base + off); };
    const int C = g->C, H1 = g->H1, H2 = g->H2, H3 = g->H3, k = g->k, N = g->n_per_cloud;
    float *pq = F(sv, p.pq), *a1 = F(sv, p.a1), *a2 = F(sv, p.a2), *a3 = F(sv, p.a3), *wc = F(sv, p.wc);
    float *bc = F(sc, p.f_bc), *wf = F(sc, p.f_wf), *bf = F(sc, p.f_bf);
    double *stats = reinterpret_cast<double *>(sc + p.f_stats);
    double *stats1 = stats, *stats2 = stats + 2 * H1, *stats3 = stats2 + 2 * H2;
    if (cudaMemsetAsync(stats, 0, (size_t)2 * (H1 + H2 + H3) * 8, st) != cudaSuccess) return fail("nt_edgeconv_train_fwd: memset failed%s", "");

    // ---- first Linear per POINT: PQ = x . [Wa - Wb ; Wb]^T + [b ; 0]
    compose_wc_kernel<<<blocks((int64_t)H1 * C > 2 * H1 ? (int64_t)H1 * C : 2 * H1), 256, 0, st>>>(g->W[0], g->b[0], H1, C, wc, bc);
    NT_TRY(check_launch("nt_edgeconv_train_fwd(compose)"));
    NT_TRY(nt_gemm_prepare_weights(wc, C, 2 * H1, C, NT_PREC_TF32X3, sc + p.f_split[0], stream));
    nt_gemm_args a{};
    a.rows = g->M; a.K = C; a.n_out = 2 * H1; a.producer = NT_PROD_PLAIN; a.epilogue = NT_EPI_BIAS;
    a.a = g->x; a.lda = g->ldx; a.w = wc; a.ldw = C; a.bias = bc; a.w_split = sc + p.f_split[0]; a.precision = NT_PREC_TF32X3;
    a.out = pq; a.ldo = p.ldpq;
    NT_TRY(nt_gemm_nt(&a, stream));

    // ---- a1 = relu(P[i] + Q[j]) + its BatchNorm statistics; BN1 folded into Linear 2
    NT_TRY(nt_edge_activation(pq, p.ldpq, H1, g->idx, k, N, p.R, H1, a1, p.ld1, stats1, stream));
    auto fold = [&](int l, double *stats_l, const float *w_next, const float *b_next, int n_next, float *w_ft) {
        float *bn = F(sv, p.bn[l]);
        const int Cn = p.H[l];
        return nt_bn_fold(stats_l, p.R, Cn, g->gamma[l], g->beta[l], g->running_mean[l], g->running_var[l], g->num_batches_tracked[l],
                          g->momentum, g->eps, 1, bn, bn + Cn, bn + 2 * Cn, bn + 3 * Cn, w_next, b_next, n_next, w_next ? wf : nullptr,
                          w_ft, w_next ? bf : nullptr, stream);
    };
    NT_TRY(fold(0, stats1, g->W[1], g->b[1], H2, F(sv, p.wft[1])));

    // ---- a2 = relu(a1 . W2'^T + b2'), statistics in the epilogue; BN2 folded into Linear 3
    NT_TRY(nt_gemm_prepare_weights(wf, H1, H2, H1, NT_PREC_TF32X3, sc + p.f_split[1], stream));
    a = nt_gemm_args{};
    a.rows = p.R; a.K = H1; a.n_out = H2; a.producer = NT_PROD_PLAIN; a.epilogue = NT_EPI_RELU_STATS;
    a.a = a1; a.lda = p.ld1; a.w = wf; a.ldw = H1; a.bias = bf; a.w_split = sc + p.f_split[1]; a.precision = NT_PREC_TF32X3;
    a.out = a2; a.ldo = p.ld2; a.stats = stats2;
    NT_TRY(nt_gemm_nt(&a, stream));
    NT_TRY(fold(1, stats2, g->W[2], g->b[2], H3, F(sv, p.wft[2])));

    // ---- a3 = relu(a2 . W3'^T + b3') with max / min over the k edge rows of every point; BN3 applied to the selected extremum
    NT_TRY(nt_gemm_prepare_weights(wf, H2, H3, H2, NT_PREC_TF32X3, sc + p.f_split[2], stream));
    a = nt_gemm_args{};
    a.rows = p.R; a.K = H2; a.n_out = H3; a.producer = NT_PROD_PLAIN; a.epilogue = NT_EPI_RELU_MAXMIN; a.k = k;
    a.a = a2; a.lda = p.ld2; a.w = wf; a.ldw = H2; a.bias = bf; a.w_split = sc + p.f_split[2]; a.precision = NT_PREC_TF32X3;
    a.out = a3; a.ldo = p.ld3; a.stats = stats3;
    a.vmax = F(sc, p.f_vmax); a.vmin = F(sc, p.f_vmin); a.imax = sc + p.f_imax; a.imin = sc + p.f_imin;
    NT_TRY(nt_gemm_nt(&a, stream));
    NT_TRY(fold(2, stats3, nullptr, nullptr, 0, nullptr));
    const float *bn3 = F(sv, p.bn[2]);
    return nt_maxmin_finish(a.vmax, a.vmin, a.imax, a.imin, bn3 + 2 * H3, bn3 + 3 * H3, g->M, H3, g->out, g->ldo, sv + p.sel, F(sv, p.vsel),
                            g->tail_src, g->tail_ld, g->tail, stream);
}

extern "C" int nt_edgeconv_train_bwd(const nt_edgeconv_args *g, void *stream) {
    Plan p;
    NT_REQUIRE(make_plan(g, p), "nt_edgeconv_train_bwd: bad shape");
    if (g->M == 0) return 0;
    NT_TRY(check_common(g));
    NT_REQUIRE(g->gout && g->ldg >= g->H3, "nt_edgeconv_train_bwd: bad upstream gradient");
    for (int l = 0; l < 3; ++l) NT_REQUIRE(g->gW[l] && g->gb[l] && g->ggamma[l] && g->gbeta[l], "nt_edgeconv_train_bwd: gradient outputs missing");
    NT_REQUIRE(!g->gx || g->ldgx >= g->C, "nt_edgeconv_train_bwd: bad gx");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t *sv = reinterpret_cast<uint8_t *>(g->saved), *sc = reinterpret_cast<uint8_t *>(g->scratch);
    auto F = [](uint8_t *base, size_t off) { return reinterpret_cast<float *>(base + off); };
    const int C = g->C, H1 = g->H1, H3 = g->H3, k = g->k, N = g->n_per_cloud;
    const int64_t R = p.R;
    float *pq = F(sv, p.pq), *wc = F(sv, p.wc);
    (void)pq;
    float *acts[4] = {nullptr, F(sv, p.a1), F(sv, p.a2), F(sv, p.a3)};
    const int lds[4] = {0, p.ld1, p.ld2, p.ld3};
    float *dzs[4] = {nullptr, F(sc, p.b_dz1), F(sc, p.b_dz2), F(sc, p.b_dz3)};
    double *acc = reinterpret_cast<double *>(sc + p.b_acc);
    if (cudaMemsetAsync(acc, 0, p.acc_doubles * 8, st) != cudaSuccess) return fail("nt_edgeconv_train_bwd: memset failed%s", "");
    double *sums3 = acc, *csum = acc + 2 * H3;                  // csum of the layer being processed
    double *next = csum + H3;

    // ---- trailing BN (behind the aggregation): column sums over the points, then dz3 for every edge row
    const float *bn3 = F(sv, p.bn[2]);
    NT_TRY(nt_bn_bwd_reduce(g->gout, g->ldg, F(sv, p.vsel), H3, bn3, bn3 + H3, g->M, H3, sums3, stream));
    NT_TRY(nt_bn_relu_bwd_last(acts[3], lds[3], g->gout, g->ldg, sv + p.sel, k, bn3 + 2 * H3, bn3, bn3 + H3, sums3, R, R, H3, dzs[3], lds[3],
                               csum, stream));

    // ---- Linear 3 and Linear 2 (index l = 2, 1): weight gradient + BN-backward moments, then the data gradient through BN + ReLU
    for (int l = 2; l >= 1; --l) {
        const int Hout = p.H[l], Hin = p.H[l - 1];
        const float *pbn = F(sv, p.bn[l - 1]);                  // mean | rstd | s | t of the BN in front of Linear l
        double *raw = next; next += (size_t)Hout * Hin;
        double *csum_prev = next; next += Hin;
        NT_TRY(nt_gemm_tn_centered(dzs[l + 1], lds[l + 1], Hout, acts[l], lds[l], Hin, R, nullptr, 0, 0, nullptr, 1, 1, pbn, raw, Hin,
                                   sc + p.b_tn, stream));
        float *vec = F(sc, p.b_vec[l]);                         // k0 | k1
        NT_TRY(nt_linear_bn_bwd(raw, csum, Hout, Hin, g->W[l], pbn + 2 * Hin, g->beta[l - 1], pbn + Hin, R, g->gW[l], g->gb[l],
                                g->ggamma[l - 1], g->gbeta[l - 1], vec, vec + Hin, stream));
        const float *wft = F(sv, p.wft[l]);                     // [Hin, Hout]
        NT_TRY(nt_gemm_prepare_weights(wft, Hout, Hin, Hout, NT_PREC_TF32X3, sc + p.b_split[l], stream));
        nt_gemm_args a{};
        a.rows = R; a.K = Hout; a.n_out = Hin; a.producer = NT_PROD_PLAIN; a.epilogue = NT_EPI_BNRELU_BWD;
        a.a = dzs[l + 1]; a.lda = lds[l + 1]; a.w = wft; a.ldw = Hout; a.w_split = sc + p.b_split[l]; a.precision = NT_PREC_TF32X3;
        a.out = dzs[l]; a.ldo = lds[l]; a.aux = acts[l]; a.ldaux = lds[l]; a.k0 = vec; a.k1 = vec + Hin; a.mu = pbn; a.colsum = csum_prev;
        NT_TRY(nt_gemm_nt(&a, stream));
        csum = csum_prev;
    }

    // ---- first Linear (per point): scatter dz1 to dPQ, weight gradient against x, input gradient
    float *dpq = F(sc, p.b_dpq), *dwc = F(sc, p.b_dwc);
    if (cudaMemsetAsync(dpq, 0, (size_t)g->M * 2 * H1 * 4, st) != cudaSuccess || cudaMemsetAsync(dwc, 0, (size_t)2 * H1 * C * 4, st) != cudaSuccess)
        return fail("nt_edgeconv_train_bwd: memset failed%s", "");
    NT_TRY(nt_edge_scatter(dzs[1], lds[1], g->idx, k, N, g->M, H1, dpq, 2 * H1, stream));
    NT_TRY(nt_gemm_tn(dpq, 2 * H1, 2 * H1, g->x, g->ldx, C, g->M, nullptr, 0, 0, nullptr, 1, 1, dwc, C, sc + p.b_tn, stream));
    const int64_t nfin = (int64_t)H1 * C > H3 ? (int64_t)H1 * C : H3;
    finish_first_kernel<<<blocks(nfin > H1 ? nfin : H1), 256, 0, st>>>(dwc, H1, C, g->gW[0], csum, g->gb[0], sums3, H3, g->gbeta[2], g->ggamma[2]);
    NT_TRY(check_launch("nt_edgeconv_train_bwd(finish)"));
    if (g->gx) {
        float *wct = F(sc, p.b_wct);
        transpose_kernel<<<blocks((int64_t)2 * H1 * C), 256, 0, st>>>(wc, 2 * H1, C, wct);
        NT_TRY(check_launch("nt_edgeconv_train_bwd(transpose)"));
        NT_TRY(nt_gemm_prepare_weights(wct, 2 * H1, C, 2 * H1, NT_PREC_TF32X3, sc + p.b_split[0], stream));
        nt_gemm_args a{};
        a.rows = g->M; a.K = 2 * H1; a.n_out = C; a.producer = NT_PROD_PLAIN; a.epilogue = NT_EPI_BIAS;
        a.a = dpq; a.lda = 2 * H1; a.w = wct; a.ldw = 2 * H1; a.w_split = sc + p.b_split[0]; a.precision = NT_PREC_TF32X3;
        a.out = g->gx; a.ldo = g->ldgx;
        NT_TRY(nt_gemm_nt(&a, stream));
    }
    return 0;
}
