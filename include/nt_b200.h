/*
 * nt_b200.h -- C ABI of the B200-native NeuralTailor hot path (libnt_b200.so, sm_100a).
 *
 * The reference (maria-korosteleva/Garment-Pattern-Estimation) owns no native code: its hot path reaches
 * third-party CUDA/ATen kernels through Python.  Every entry point below replaces one of those borrowed
 * operators; the comment above each names the reference call site (file:line under /root/reference) it serves.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated; fp32 row-major;
 *     `ld*` arguments are row strides in elements.
 *   - the caller owns every buffer (PyTorch tensors in the shipped host code); the library never allocates or
 *     frees user-visible memory and keeps no global mutable state (nt_launch_count's counter and the debug hooks aside).
 *   - every function only ENQUEUES work on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - return value: 0 = ok, non-zero = error (message via nt_last_error(), thread-local).  No exceptions cross
 *     the ABI.  The Python host maps non-zero to RuntimeError (reference convention: nn/trainer.py:60).
 *   - "rows" are points (M = B*N) or edges (E = M*k); edge row e belongs to centre point e / k, slot e % k,
 *     and its neighbour is  (e / k / N) * N + idx[e]   (idx holds indices LOCAL to the cloud).
 */
#ifndef NT_B200_H
#define NT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------- */
const char *nt_last_error(void);
int nt_version(void);                          /* ABI version, bumped on incompatible change */
int nt_built_arch(void);                       /* 100 for sm_100a */
/* number of kernels launched by this library in the calling process since load (bench.py's gpu_launches) */
int64_t nt_launch_count(void);
/* sizeof() of an argument struct of this header by name ("nt_gemm_args", "nt_pattern_loss_args", "nt_lstm_sizes_t"; 0 = unknown), so
 * that a foreign-language binding can verify its mirror of the layout when it loads the library */
int nt_sizeof(const char *struct_name);

/* ---- kNN graph: torch_cluster.knn via DynamicEdgeConv.forward (nn/net_blocks.py:127-135,174) ------------
 * For every point of every cloud: the k nearest points of the SAME cloud by squared L2 over D features, self
 * included, ascending by (distance, index); distance = sequential fp32 fma chain over d = 0..D-1 (bit-exact with
 * oracle/knn_oracle.c).  x: [B*N, >=D] with row stride ldx.  idx: [B*N, k] int32, local to the cloud.
 * Requires 1 <= k <= 32 and k <= N.
 * workspace: nt_knn_workspace_bytes(B, N, k) bytes of device memory (partial lists when the candidate scan is split over
 * several CTAs to fill the GPU); NULL = never split. */
int64_t nt_knn_workspace_bytes(int B, int N, int k);
int nt_knn(const float *x, int B, int N, int D, int ldx, int k, int32_t *idx, void *workspace, void *stream);

/* ---- generic fused row GEMM:  out = epilogue( producer(A) . W^T )   (W: [n_out, K] like nn.Linear.weight) --
 * Serves the Linear layers of MLP() (nn/net_blocks.py:43-47) inside DynamicEdgeConv.message (edge rows) and
 * point_segment_mlp / panel_dec_lin / placement_decoder (nn/nets.py:223-233,254-276; point rows).
 *
 * producer (what a row of A is):
 *   NT_PROD_PLAIN : A[r, :] = a[r*lda + :]
 *   NT_PROD_EDGE  : A[r, :] = relu( pq[c*ldpq + :] + pq[j*ldpq + qoff + :] ),  c = r / k, j = neighbour of edge r
 *                   (idx == NULL: A[r, :] = relu(pq[r*ldpq + :]) -- the first activation of a per-point MLP)
 * epilogue:
 *   NT_EPI_BIAS        : out = acc + bias
 *   NT_EPI_RELU_STATS  : v = relu(acc + bias); out = v (if out != NULL); stats[0:n_out] += sum_r v,
 *                        stats[n_out:2n_out] += sum_r v*v (double; training-mode BatchNorm1d statistics)
 *   NT_EPI_RELU_MAXMIN : RELU_STATS plus, per centre point, max and min of v over its k edge rows with the
 *                        first slot attaining them (EdgeConv max aggregation commuted past the trailing BN)
 *   NT_EPI_BNRELU_BWD  : g = acc; a = aux row (plain or EDGE producer); dz = a > 0 ? g - k0 - (a - mu)*k1 : 0;
 *                        out = dz; colsum[0:n_out] += sum_r dz   (backward of Linear<-BN<-ReLU in one pass)
 */
enum { NT_PROD_PLAIN = 0, NT_PROD_EDGE = 1 };
enum { NT_EPI_BIAS = 0, NT_EPI_RELU_STATS = 1, NT_EPI_RELU_MAXMIN = 2, NT_EPI_BNRELU_BWD = 3 };
/* tensor-core operand precision: both split every fp32 operand into hi + lo and issue hi.hi + hi.lo + lo.hi into an fp32
 * TMEM accumulator.  TF32X3 (~1e-6 relative, fp32-SGEMM-like; forward path, keeps the dynamic kNN graph and the ReLU
 * masks aligned with the fp32 reference) costs twice the tensor-pipe time of BF16X3 (~1e-5; gradient GEMMs). */
enum { NT_PREC_BF16X3 = 0, NT_PREC_TF32X3 = 1 };

typedef struct nt_gemm_args {
    /* problem */
    int64_t rows;            /* rows of A / out */
    int K;                   /* inner dimension */
    int n_out;               /* output columns */
    int producer, epilogue;
    /* NT_PROD_PLAIN */
    const float *a; int lda;
    /* NT_PROD_EDGE (also used by the aux operand of NT_EPI_BNRELU_BWD when aux_edge != 0) */
    const float *pq; int ldpq; int qoff;
    const int32_t *idx; int k; int n_per_cloud;
    /* weights: w (fp32, [n_out, K]) and w_split (its hi / lo planes from nt_gemm_prepare_weights) are both required: the GEMMs run
     * on the tcgen05 tensor cores only (no CUDA-core or CPU fallback) */
    const float *w; int ldw; const float *bias; const void *w_split; int precision;
    /* outputs */
    float *out; int ldo;
    double *stats;                                   /* [2*n_out] */
    float *vmax, *vmin; uint8_t *imax, *imin;        /* [rows/k, n_out] */
    /* NT_EPI_BNRELU_BWD */
    const float *aux; int ldaux; int aux_edge;
    const float *k0, *k1, *mu;                       /* [n_out] each */
    double *colsum;                                  /* [n_out] */
    /* NT_EPI_BNRELU_BWD, optional fused scatter of dz into the per-point gradient of the split first EdgeConv Linear (what
     * nt_edge_scatter does in a second pass): scatter_dpq[c, 0:n_out] = sum of dz over the k edge rows of centre c (plain
     * store), scatter_dpq[j, n_out:2*n_out] += dz[e] for the neighbour j = idx[e] of every edge e (atomic; the caller zeroes
     * that half).  Uses idx / k / n_per_cloud; `out` may then be NULL (dz is never written).  Only for calls for which
     * nt_gemm_nt_scatter_supported() returns 1; otherwise nt_gemm_nt fails. */
    float *scatter_dpq; int ldscatter;
    /* kernel engine for THIS call (TF32X3 only): 0 = auto (streaming persistent engine for large aligned calls, one tile per CTA
     * otherwise), 1 = one tile per CTA, 3 / 4 = streaming engine with one / two row tiles per weight stage, 5 = streaming engine with
     * the aux rows of NT_EPI_BNRELU_BWD staged through a shared-memory ring, 6 = second-generation streaming engine (A operand in
     * tensor memory, TMA tensor-map epilogue; what auto picks when the call is eligible).  Results are bit-identical across
     * engines (same operand split, same accumulation order); the field exists for tests and measurements. */
    int engine;
} nt_gemm_args;

int nt_gemm_nt(const nt_gemm_args *args, void *stream);
/* 1 if nt_gemm_nt(args) would run the fused-scatter epilogue (streaming engine, aligned operands), 0 otherwise. */
int nt_gemm_nt_scatter_supported(const nt_gemm_args *args);

/* Tensor-core operand preparation: splits W [n_out, K] (fp32) into bf16 hi/lo planes laid out as UMMA core matrices per
 * (column tile, K block of 32 bf16 / 16 tf32 elements).  w_split must hold nt_gemm_weights_bytes(n_out, K) bytes, 16-byte aligned. */
int64_t nt_gemm_weights_bytes(int n_out, int K, int precision);
int nt_gemm_prepare_weights(const float *w, int ldw, int n_out, int K, int precision, void *w_split, void *stream);

/* Weight-gradient GEMM: out[m, n] += sum_r A[r, m] * Bop[r, n]  (out must be zeroed / initialised by the caller).
 * Bop is a plain matrix (b, ldb) or, when pq != NULL, the EDGE producer above.  Backward of nn.Linear.
 * workspace: nt_gemm_tn_workspace_bytes() bytes of device memory (required): tcgen05 tensor-core engine, TF32x3, per-CTA partial
 * products reduced deterministically in double. */
int64_t nt_gemm_tn_workspace_bytes(void);
int nt_gemm_tn(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows,
               const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
               float *out, int ldo, void *workspace, void *stream);

/* Same contraction with the B operand centred per column (Bop[r, n] - mu[n]) and accumulated in DOUBLE across CTAs
 * (each CTA sums at most 1024 rows in fp32): the moments the BatchNorm backward needs. */
int nt_gemm_tn_centered(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows,
                        const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                        const float *mu, double *out, int ldo, void *workspace, void *stream);

/* ---- BatchNorm1d bookkeeping (nn/net_blocks.py:46; BN placed AFTER ReLU, statistics over all rows) ---------
 * Turns accumulated statistics (or running statistics when training == 0) into the affine y = a*s + t, updates
 * the running buffers (momentum, unbiased variance) in training mode, and folds the affine into the NEXT Linear:
 *   w_f = w_next * diag(s)  [n_next, C],  w_ft = w_f^T  [C, n_next],  b_f = b_next + w_next . t.
 * w_next may be NULL (last layer).  Outputs mean/rstd/s/t are [C]. */
int nt_bn_fold(const double *stats, int64_t count, int C, const float *gamma, const float *beta,
               float *running_mean, float *running_var, int64_t *num_batches_tracked,
               float momentum, float eps, int training,
               float *mean, float *rstd, float *s, float *t,
               const float *w_next, const float *b_next, int n_next, float *w_f, float *w_ft, float *b_f,
               void *stream);

/* Statistics of the first activation of an MLP, a1 = NT_PROD_EDGE row (see above): stats[0:H] += sum a1,
 * stats[H:2H] += sum a1^2 over `rows` rows (double). */
int nt_edge_stats(const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                  int64_t rows, int H, double *stats, void *stream);

/* Materialised first edge activation: out[e, 0:H] = relu(pq[centre(e), 0:H] + pq[nbr(e), qoff:qoff+H]) for `rows` edge rows
 * (DynamicEdgeConv.message input after the algebraic split of the first Linear, nn/net_blocks.py:127-135), with the
 * BatchNorm statistics of nt_edge_stats accumulated in the same pass when stats != NULL.  idx == NULL: plain rows,
 * out[r] = relu(pq[r, 0:H]) (first activation of the per-point MLP, nn/net_blocks.py:43-47). */
int nt_edge_activation(const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                       int64_t rows, int H, float *out, int ldo, double *stats, void *stream);

/* EdgeConv aggregation finish (DynamicEdgeConv aggr='max', nn/net_blocks.py:127-135): out[m, c] =
 * s[c] >= 0 ? s*vmax + t : s*vmin + t; sel = slot that produced it, vsel = the pre-BN value it had (both optional,
 * [M, C], kept for the backward).  Optionally appends `tail` columns copied from tail_src (the skip connection,
 * nn/net_blocks.py:178-180).  out row stride ldo. */
int nt_maxmin_finish(const float *vmax, const float *vmin, const uint8_t *imax, const uint8_t *imin,
                     const float *s, const float *t, int64_t M, int C, float *out, int ldo, uint8_t *sel,
                     float *vsel, const float *tail_src, int tail_ld, int tail, void *stream);

/* y = a*s + t on [rows, C] (trailing BN of a per-point MLP). */
int nt_bn_apply(const float *a, int lda, const float *s, const float *t, int64_t rows, int C,
                float *out, int ldo, void *stream);

/* Column reductions for the trailing BN's backward.  g: [rows_g, C] upstream gradient; v: the BN input the gradient
 * refers to (same shape): sums[0:C] += sum g, sums[C:2C] += sum g * (v - mu) * rstd  (double accumulators).
 * v == NULL: plain column sum only (bias gradient of a Linear). */
int nt_bn_bwd_reduce(const float *g, int ldg, const float *v, int ldv, const float *mu, const float *rstd,
                     int64_t rows, int C, double *sums, void *stream);

/* Backward through the trailing BN + ReLU for every row: dz[r, c] = a > 0 ? s*(gsel - dbeta/cnt - xhat*dgamma/cnt) : 0
 * where gsel = g[r / k, c] if sel == NULL or sel[r / k, c] == r % k, else 0.  a, dz: [rows, C]; colsum += sum dz. */
int nt_bn_relu_bwd_last(const float *a, int lda, const float *g, int ldg, const uint8_t *sel, int k,
                        const float *s, const float *mu, const float *rstd, const double *sums, int64_t count,
                        int64_t rows, int C, float *dz, int lddz, double *colsum, void *stream);

/* Backward bookkeeping of one folded Linear (Linear_l fed by BN_{l} output y = a*s + t):
 * from rawc = dz^T . (a_prev - mean) [n_out, C] (nt_gemm_tn_centered, double), csum = sum_r dz [n_out] (double) and the
 * previous BN's (s, beta, rstd):  dW = s*rawc + csum (x) beta,  db = csum,  dbeta = W^T csum,
 * dgamma[c] = rstd[c] * sum_o W[o,c]*rawc[o,c], and the vectors of the next NT_EPI_BNRELU_BWD epilogue:
 * k0 = s*dbeta/cnt, k1 = s*rstd*dgamma/cnt.  (Centring removes the mean^2/var cancellation of the raw moment.) */
int nt_linear_bn_bwd(const double *rawc, const double *csum, int n_out, int C, const float *w,
                     const float *s, const float *beta, const float *rstd, int64_t count,
                     float *dW, float *db, float *dgamma, float *dbeta, float *k0, float *k1, void *stream);

/* EdgeConv first-layer scatter: dz: [E, H] gradient of the pre-activation P[c] + Q[j];
 * dpq[c, 0:H] = sum over the k slots, dpq[j, H:2H] += dz (atomic).  dpq must be zeroed by the caller. */
int nt_edge_scatter(const float *dz, int lddz, const int32_t *idx, int k, int n_per_cloud, int64_t M, int H,
                    float *dpq, int lddpq, void *stream);

/* ---- attention head ----------------------------------------------------------------------------------------
 * Sparsemax over the last dimension (sparsemax.Sparsemax(dim=1), nn/nets.py:225,255): P <= 32 columns. */
int nt_sparsemax_fwd(const float *z, int64_t rows, int P, float *out, void *stream);
int nt_sparsemax_bwd(const float *out, const float *g, int64_t rows, int P, float *gz, void *stream);

/* The 23-iteration weighted pooling loop of nn/nets.py:263-276 as one contraction:
 * enc[b, p, f] = scale * sum_n w[b, n, p] * feat[b, n, f]   (scale = 1/N for global_mean_pool).  enc zeroed by caller. */
int nt_attn_pool_fwd(const float *w, const float *feat, int ldf, int B, int N, int P, int F, float scale,
                     float *enc, void *stream);
/* gw[b,n,p] = scale * sum_f genc[b,p,f]*feat[b,n,f];  gfeat[b,n,f] (+)= scale * sum_p w[b,n,p]*genc[b,p,f] */
int nt_attn_pool_bwd(const float *genc, const float *w, const float *feat, int ldf, int B, int N, int P, int F,
                     float scale, float *gw, float *gfeat, int ldgf, int accumulate_gfeat, void *stream);

/* Global pooling over the N points of each of B equal-size clouds (torch_geometric.nn.global_{mean,max,add}_pool as used by
 * EdgeConvFeatures, nn/net_blocks.py:145-150,182-187): out[b, f] = mean / max / sum over n of x[b*N + n, f].
 * argmax ([B, F] int32, may be NULL unless the backward of max is needed) receives the first maximal point.
 * Backward: gx[b*N + n, f] = g[b, f] / N (mean), g[b, f] (add), g[b, f] * [n == argmax[b, f]] (max); gx is overwritten. */
#define NT_POOL_MEAN 0
#define NT_POOL_MAX 1
#define NT_POOL_ADD 2
int nt_global_pool_fwd(const float *x, int ldx, int B, int N, int F, int mode, float *out, int32_t *argmax, void *stream);
int nt_global_pool_bwd(const float *g, const int32_t *argmax, int B, int N, int F, int mode, float *gx, int ldgx,
                       void *stream);

/* ---- LSTM panel decoder: nn.LSTM(batch_first=True) inside LSTMDecoderModule.forward (nn/net_blocks.py:373,382-402) ------------
 * L layers (<= 4), hidden H (<= 255), gates in PyTorch order i, f, g, o; the layer-0 input is the SAME vector x[r, :E] (E <= 256)
 * at each of the T steps -- the reference repeats the encoding (net_blocks.py:388).  R rows (sequences) are independent.
 * One persistent cooperative kernel per direction (csrc/lstm.cu): weights stationary in shared memory, tcgen05 BF16x3 products,
 * h exchanged between CTAs through L2 in the tensor-core operand layout.  Needs a device on which L*16 CTAs are co-resident.
 *
 * Buffers (all caller-owned device memory; sizes from nt_lstm_sizes):
 *   weights   : prepared by nt_lstm_prepare_weights from weight_ih_l*, weight_hh_l*, bias_ih_l*, bias_hh_l* (host arrays of L
 *               device pointers); valid until the parameters change.  128-byte aligned.
 *   y         : fp32 [T][R][y_ld] -- the module output, TIME-major (top layer's h_t); columns >= H are scratch.
 *   act       : the hidden states of all layers / steps in the tensor-core operand layout (opaque; 256-byte aligned).  Scratch
 *               in inference; in training it is kept for nt_lstm_bwd together with
 *   cs, gates : saved cell states / activated gates (opaque layout); both NULL in inference.
 *   workspace : scratch, 256-byte aligned; contents are not needed after the call returns.
 * nt_lstm_bwd: dy [T][R][ld_dy] = gradient w.r.t. y.  Writes dx [R][ld_dx] (gradient w.r.t. x, may be NULL) and, per layer,
 * dw_ih [4H, in], dw_hh [4H, H], db_ih = db_hh [4H] (host arrays of L device pointers; entries or whole arrays may be NULL). */
typedef struct nt_lstm_sizes_t {
    int64_t weights_bytes, act_bytes, fwd_workspace_bytes, y_bytes, cs_bytes, gates_bytes, bwd_workspace_bytes;
    int y_ld;
} nt_lstm_sizes_t;
int nt_lstm_sizes(int R, int T, int L, int H, int E, nt_lstm_sizes_t *out);
int nt_lstm_prepare_weights(const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                            const float *const *b_hh, int L, int H, int E, void *weights, void *stream);
int nt_lstm_fwd(const float *x, int ldx, const float *h0, const float *c0, const void *weights, int R, int T, int L, int H,
                int E, float *y, void *act, float *cs, float *gates, void *workspace, void *stream);
int nt_lstm_bwd(const float *dy, int ld_dy, const void *act, const float *cs, const float *gates, const void *weights, int R,
                int T, int L, int H, int E, void *workspace, float *dx, int ld_dx, float *const *dw_ih, float *const *dw_hh,
                float *const *db_ih, float *const *db_hh, void *stream);

/* ---- composite EdgeConv layer, TRAINING mode (DynamicEdgeConv over MLP([2C, H1, H2, H3]) with batch-statistics BatchNorm;
 * nn/net_blocks.py:126-135,172-180; SURVEY 8b "nt_edgeconv_fwd/bwd ... training variants with BN stat buffers") --------------------
 * One call per direction runs the whole layer on `stream` with this library's kernels (forward: split first Linear per point ->
 * nt_edge_activation -> nt_bn_fold -> nt_gemm_nt(RELU_STATS) -> nt_bn_fold -> nt_gemm_nt(RELU_MAXMIN) -> nt_bn_fold ->
 * nt_maxmin_finish; backward: the mirror image, down to the gradients of x and of every parameter).  The caller allocates
 * `saved` (nt_edgeconv_saved_bytes(): written by the forward, read by the backward -- activations a1..a3, PQ, BatchNorm vectors,
 * folded transposed weights) and `scratch` (nt_edgeconv_scratch_bytes(args, backward): temporaries, reusable once the stream has
 * passed the call); both 256-byte aligned device memory.  Row layouts: x [M, C] (row stride ldx), idx [M, k] int32 LOCAL to the
 * cloud (nt_knn), out / gout [M, H3 + tail], W[l] contiguous [H_l, in_l] (in_0 = 2C), gradients contiguous and OVERWRITTEN.
 * The running BatchNorm buffers are updated in place like nn.BatchNorm1d (momentum; num_batches_tracked += 1).  TF32x3 products. */
typedef struct nt_edgeconv_args {
    int64_t M;                                   /* points = clouds * n_per_cloud */
    int C, H1, H2, H3, k, n_per_cloud;
    const float *x; int ldx;
    const int32_t *idx;
    const float *tail_src; int tail_ld, tail;    /* skip-connection columns copied to out[:, H3:H3+tail] (tail = 0: none) */
    const float *W[3], *b[3], *gamma[3], *beta[3];
    float *running_mean[3], *running_var[3]; int64_t *num_batches_tracked[3];
    float momentum, eps;
    float *out; int ldo;                         /* forward result */
    void *saved, *scratch;
    /* backward only */
    const float *gout; int ldg;
    float *gx; int ldgx;                         /* [M, C] or NULL */
    float *gW[3], *gb[3], *ggamma[3], *gbeta[3];
} nt_edgeconv_args;
int64_t nt_edgeconv_saved_bytes(const nt_edgeconv_args *args);                  /* -1: bad shape */
int64_t nt_edgeconv_scratch_bytes(const nt_edgeconv_args *args, int backward);
int nt_edgeconv_train_fwd(const nt_edgeconv_args *args, void *stream);
int nt_edgeconv_train_bwd(const nt_edgeconv_args *args, void *stream);

/* ---- fused inference EdgeConv layer (DynamicEdgeConv in eval mode; nn/net_blocks.py:126-135,172-180) ------------------------------
 * out[i, 0:C] = s*max|min_{j in kNN(i)} relu(W3' relu(W2' relu(P[i] + Q[j]) + b2') + b3') + t  in ONE kernel (gather, two chained
 * tcgen05 GEMMs with the intermediate kept in TMEM / shared memory, max over the k rows of a point, trailing BatchNorm affine,
 * optional skip-connection columns out[i, C:C+tail] = tail_src[i]).  pq = [P | Q] per point (row stride ldpq, Q at column H1) from
 * the split first Linear; w2_split / w3_split = nt_gemm_prepare_weights(precision) of the BatchNorm-folded weights W2' [H2, H1],
 * W3' [C, H2] (nt_bn_fold with training = 0); precision = NT_PREC_TF32X3 (fp32-class, what the shipped host code uses) or
 * NT_PREC_BF16X3; s_out / t_out = the folded affine of the last BatchNorm.
 * Sizes: nt_edgeconv_eval_supported(H1, H2, C, k, ldpq) must return 1. */
int nt_edgeconv_eval_supported(int H1, int H2, int C, int k, int ldpq);
int nt_edgeconv_eval_fwd(const float *pq, int ldpq, int H1, const int32_t *idx, int k, int n_per_cloud, int64_t M,
                         const void *w2_split, const float *b2, int H2, const void *w3_split, const float *b3, int C,
                         int precision, const float *s_out, const float *t_out, const float *tail_src, int tail_ld, int tail,
                         float *out, int ldo, void *stream);

/* ---- PointNet++ set abstraction (nn/net_blocks.py:10-88: torch_geometric.nn.fps / radius / PointConv) --------------------------
 * pos: [B*N, >=D] fp32, row stride ld, B equal-size clouds, D <= 8 coordinates.  Index outputs are LOCAL to the cloud.
 * nt_fps: farthest point sampling, n_samples = ceil(ratio * N) per cloud, starting from point 0 of every cloud (torch_cluster's
 *   random_start=False; the library default, a random first point, cannot be pinned); ties -> lowest index.  idx: [B, n_samples].
 * nt_radius: for every centre (centres: [B, M] local point indices) the first max_nbr points of its cloud, ascending index, with
 *   squared distance < r*r.  nbr: [B*M, max_nbr] (-1 padded), count: [B*M].
 * nt_point_edges_count / _fill: the edge list PointConv builds in the reference's bipartite call (radius edges whose global source
 *   index equals their target index removed, then one edge i -> i appended per centre i < min(B*N, B*M) -- the library's
 *   index-based self-loop handling), with the message input msg[e, :D] = pos[src] - pos[centre(dst)].  keep_count: [B*M] kept
 *   radius edges per centre; offsets: their exclusive prefix sum (caller); n_radius_edges = their total; E = n_radius_edges +
 *   min(B*N, B*M).  src / dst: [E] int64 GLOBAL rows.
 * nt_scatter_max_fwd: out[t, f] = max over edges with dst == t of v[e, f] (0 for targets without edges), arg[t, f] = lowest such
 *   edge (INT64_MAX if none); key_scratch: T*F int32.  nt_scatter_max_bwd: gv[arg[t, f], f] = g[t, f] (gv zeroed by the caller). */
int nt_fps(const float *pos, int ld, int B, int N, int D, int n_samples, int32_t *idx, void *stream);
int nt_radius(const float *pos, int ld, int B, int N, int D, const int32_t *centres, int M, float r, int max_nbr, int32_t *nbr,
              int32_t *count, void *stream);
int nt_point_edges_count(const int32_t *nbr, const int32_t *count, int B, int N, int M, int max_nbr, int32_t *keep_count,
                         void *stream);
int nt_point_edges_fill(const float *pos, int ld, int D, const int32_t *centres, const int32_t *nbr, const int32_t *count,
                        const int64_t *offsets, int B, int N, int M, int max_nbr, int64_t n_radius_edges, int64_t *src,
                        int64_t *dst, float *msg, int ldm, void *stream);
int nt_scatter_max_fwd(const float *v, int ldv, const int64_t *dst, int64_t E, int F, int64_t T, float *out, int64_t *arg,
                       int32_t *key_scratch, void *stream);
int nt_scatter_max_bwd(const float *g, const int64_t *arg, int64_t T, int F, float *gv, int ldg, void *stream);

/* ---- stage-2 input batching: NNSewingPattern.all_edge_pairs (nn/data/pattern_converter.py:458-499) -----------------------------
 * edges: [P, Lmax, F] 3D edge features per panel (device).  Block b enumerates, in the reference's loop order, the panel pairs
 * (blk_i[b] < blk_j[b]) that both have edges; the block holds rows(i) x blk_cols[b] pairs, row-major, starting at pair blk_off[b]
 * (device arrays built by the host from the per-panel edge counts).  pairs: [n_pairs, 2F] = cat(edge of panel i, edge of panel j);
 * mapping (optional): [n_pairs, 4] = (panel i, edge, panel j, edge). */
int nt_edge_pairs(const float *edges, int Lmax, int F, const int32_t *blk_i, const int32_t *blk_j, const int32_t *blk_cols,
                  const int64_t *blk_off, int n_blocks, int64_t n_pairs, float *pairs, int32_t *mapping, void *stream);

/* ---- training step: loss and optimizer (nn/trainer.py:96-101) ---------------------------------------------------------------------
 * The four loss terms of the shipped attention config (models/att/att.yaml:124) -- nn.MSELoss on outlines / rotations /
 * translations (nn/metrics/composed_loss.py:301-321) and PanelLoopLoss (nn/metrics/losses.py:19-51: for every panel with
 * num_edges >= 3, the squared sum of its first num_edges edge vectors minus the padding vector, averaged over ALL panels x 2) --
 * in one pass.  Predictions may be strided views (element strides; the last dimension must be dense), ground truth is contiguous.
 * fwd: acc5 = 5 doubles of scratch, out5 = {total, pattern_loss, loop_loss, rotation_loss, translation_loss} (device floats).
 * bwd: gradients w.r.t. the three predictions (contiguous), multiplied by the device scalar *grad_scale. */
typedef struct nt_pattern_loss_args {
    const float *outlines; int64_t outl_stride_b, outl_stride_p, outl_stride_e;     /* [B, P, Lp, D] */
    const float *rotations; int64_t rot_stride_b, rot_stride_p;                      /* [B, P, Dr] */
    const float *translations; int64_t tr_stride_b, tr_stride_p;                     /* [B, P, Dt] */
    const float *gt_outlines, *gt_rotations, *gt_translations;
    const int64_t *num_edges;                                                        /* [B, P] */
    int B, P, Lp, D, Dr, Dt;
    float pad_x, pad_y, loop_weight;
    int use_shape, use_loop, use_rotation, use_translation;
} nt_pattern_loss_args;
int nt_pattern_loss_fwd(const nt_pattern_loss_args *args, double *acc5, float *out5, void *stream);
int nt_pattern_loss_bwd(const nt_pattern_loss_args *args, const float *grad_scale, float *g_outlines, float *g_rotations,
                        float *g_translations, void *stream);

/* torch.optim.Adam (nn/trainer.py:64,98-99; amsgrad off) on one flat buffer of n parameters: g = grads * grad_scale
 * (+ weight_decay * p), moment updates, bias-corrected update, and -- if zero_grad -- grads = 0 (optimizer.zero_grad).
 * lr: device scalar (a stepping scheduler only rewrites it).  state2: two device floats, zero-initialised by the caller once:
 * [0] = number of steps taken (advanced by the kernel), [1] = internal.  CUDA-graph safe. */
int nt_adam_step(float *params, float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, const float *lr, float beta1,
                 float beta2, float eps, float weight_decay, float grad_scale, int zero_grad, float *state2, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NT_B200_H */
