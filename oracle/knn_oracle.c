/*
 * ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the brute-force k-nearest-neighbour search the reference reaches through
 *   nn/net_blocks.py:127-135,174  ->  torch_geometric.nn.DynamicEdgeConv  ->  torch_cluster.knn
 * The arithmetic lives in the third-party package torch-cluster (unpinned in the reference's
 * requirements.txt; docs/Installation.md:46-48 names the torch-1.12.0+cu116 wheels).  Its source is
 * not under /root/reference, so this file restates the published CUDA algorithm (knn_cuda.cu,
 * `knn_kernel`):
 *   - one query at a time, candidates of the SAME cloud scanned in ascending index order;
 *   - squared L2 distance accumulated sequentially over the feature dimensions in the input dtype
 *     (fp32); nvcc's default -fmad=true contracts `dist += d*d` into one fused multiply-add, so the
 *     chain is   acc = fmaf(x_d - y_d, x_d - y_d, acc),  d = 0..D-1;
 *   - a candidate is inserted before the first stored entry whose distance is STRICTLY greater, so
 *     equal distances keep the lower index first; the result is the k lexicographically smallest
 *     (distance, index) pairs in ascending order; the query itself (distance 0) is a candidate.
 * Parity status: UNPINNED against torch-cluster binaries (the package cannot be installed here);
 * pinned only against this published algorithm.  See DESIGN.md "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#define NT_ORACLE_MAX_K 128

typedef struct {
    const float *x; int64_t N, D, k, lo, hi; int32_t *idx_out; float *dist_out;
} nt_knn_job;

/* Sequential scan for the queries [lo, hi) (flattened cloud*N + query). */
static void *nt_knn_range(void *arg)
{
    const nt_knn_job *jb = (const nt_knn_job *)arg;
    const int64_t N = jb->N, D = jb->D, k = jb->k;
    float best_d[NT_ORACLE_MAX_K];
    int32_t best_i[NT_ORACLE_MAX_K];
    for (int64_t bq = jb->lo; bq < jb->hi; ++bq) {
        const int64_t b = bq / N, q = bq % N;
        const float *cloud = jb->x + b * N * D;
        const float *xq = cloud + q * D;
        for (int64_t e = 0; e < k; ++e) { best_d[e] = 1e10f; best_i[e] = -1; }
        for (int64_t c = 0; c < N; ++c) {
            const float *xc = cloud + c * D;
            float acc = 0.0f;
            for (int64_t d = 0; d < D; ++d) {
                float diff = xc[d] - xq[d];
                acc = fmaf(diff, diff, acc);
            }
            for (int64_t e = 0; e < k; ++e) {
                if (best_d[e] > acc) {
                    for (int64_t m = k - 1; m > e; --m) { best_d[m] = best_d[m - 1]; best_i[m] = best_i[m - 1]; }
                    best_d[e] = acc; best_i[e] = (int32_t)c;
                    break;
                }
            }
        }
        int32_t *io = jb->idx_out + bq * k;
        for (int64_t e = 0; e < k; ++e) io[e] = best_i[e];
        if (jb->dist_out) {
            float *dd = jb->dist_out + bq * k;
            for (int64_t e = 0; e < k; ++e) dd[e] = best_d[e];
        }
    }
    return NULL;
}

/* x: [B, N, D] fp32 row-major.  idx_out: [B, N, k] int32, index LOCAL to the cloud.
 * dist_out (optional, may be NULL): [B, N, k] fp32 squared distances.
 * nthreads: host threads to split the (independent) queries over; each query's scan stays
 * strictly sequential, so the result does not depend on the thread count.
 * Returns 0 on success, non-zero on bad arguments.  Slots beyond min(k, N) are filled with -1
 * (torch_cluster drops such edges). */
int nt_oracle_knn(const float *x, int64_t B, int64_t N, int64_t D, int64_t k,
                  int32_t *idx_out, float *dist_out, int nthreads)
{
    if (!x || !idx_out || B < 0 || N < 0 || D <= 0 || k <= 0 || k > NT_ORACLE_MAX_K) return 1;
    const int64_t total = B * N;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > total) nthreads = total > 0 ? (int)total : 1;
    nt_knn_job jobs[256];
    pthread_t tid[256];
    const int64_t chunk = (total + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        int64_t lo = t * chunk, hi = lo + chunk; if (hi > total) hi = total; if (lo > hi) lo = hi;
        jobs[t] = (nt_knn_job){x, N, D, k, lo, hi, idx_out, dist_out};
    }
    if (nthreads == 1) { nt_knn_range(&jobs[0]); return 0; }
    for (int t = 0; t < nthreads; ++t)
        if (pthread_create(&tid[t], NULL, nt_knn_range, &jobs[t]) != 0) return 2;
    for (int t = 0; t < nthreads; ++t) pthread_join(tid[t], NULL);
    return 0;
}

/* Squared distance of a single pair with the same fmaf chain (used by tests to probe ties). */
float nt_oracle_sqdist(const float *a, const float *b, int64_t D)
{
    float acc = 0.0f;
    for (int64_t d = 0; d < D; ++d) { float diff = a[d] - b[d]; acc = fmaf(diff, diff, acc); }
    return acc;
}

/* ------------------------------------------------------------------------------------------------------------------
 * PointNet++ sampling / grouping (nn/net_blocks.py:19-21 -> torch_geometric.nn.fps / radius -> torch_cluster).
 * Restated from the published CUDA algorithms of torch-cluster (fps_cuda.cu, radius_cuda.cu); UNPINNED like the kNN above.
 *
 * fps: iterative farthest point sampling inside every cloud.  torch_cluster starts from a RANDOM point by default
 * (random_start=True), which cannot be pinned; this restatement (and the B200 kernel) start from point 0 of the cloud, the
 * library's random_start=False behaviour.  Each step keeps, per point, the squared distance to the nearest selected point
 * (fp32, the same fmaf chain as above) and selects the point with the largest one; ties -> lowest index.
 * n_samples = ceil(ratio * N) is computed by the caller.
 * ------------------------------------------------------------------------------------------------------------------ */
int nt_oracle_fps(const float *pos, int64_t B, int64_t N, int64_t D, int64_t n_samples, int32_t *idx_out)
{
    if (!pos || !idx_out || B < 0 || N < 1 || D < 1 || n_samples < 1 || n_samples > N) return 1;
    float *mind = (float *)malloc((size_t)N * sizeof(float));
    if (!mind) return 2;
    for (int64_t b = 0; b < B; ++b) {
        const float *cloud = pos + b * N * D;
        for (int64_t i = 0; i < N; ++i) mind[i] = INFINITY;
        int64_t cur = 0;
        for (int64_t s = 0; s < n_samples; ++s) {
            idx_out[b * n_samples + s] = (int32_t)cur;
            const float *pc = cloud + cur * D;
            float best = -1.0f;
            int64_t best_i = 0;
            for (int64_t i = 0; i < N; ++i) {
                float acc = 0.0f;
                for (int64_t d = 0; d < D; ++d) { float diff = cloud[i * D + d] - pc[d]; acc = fmaf(diff, diff, acc); }
                if (acc < mind[i]) mind[i] = acc;
                if (mind[i] > best) { best = mind[i]; best_i = i; }
            }
            cur = best_i;
        }
    }
    free(mind);
    return 0;
}

/* radius: for every centre (given as local point indices [B, M]) the first `max_nbr` points of ITS cloud, in ascending index
 * order, whose squared distance is STRICTLY below r*r (radius_cuda.cu: `if (dist < r)` with r squared by the caller).
 * nbr_out: [B, M, max_nbr] local indices, -1 past count_out[b, m]. */
int nt_oracle_radius(const float *pos, int64_t B, int64_t N, int64_t D, const int32_t *centres, int64_t M, float r,
                     int64_t max_nbr, int32_t *nbr_out, int32_t *count_out)
{
    if (!pos || !centres || !nbr_out || !count_out || B < 0 || N < 1 || D < 1 || M < 1 || max_nbr < 1) return 1;
    const float r2 = r * r;
    for (int64_t b = 0; b < B; ++b) {
        const float *cloud = pos + b * N * D;
        for (int64_t m = 0; m < M; ++m) {
            const float *pc = cloud + (int64_t)centres[b * M + m] * D;
            int32_t *out = nbr_out + (b * M + m) * max_nbr;
            int64_t cnt = 0;
            for (int64_t i = 0; i < N && cnt < max_nbr; ++i) {
                float acc = 0.0f;
                for (int64_t d = 0; d < D; ++d) { float diff = cloud[i * D + d] - pc[d]; acc = fmaf(diff, diff, acc); }
                if (acc < r2) out[cnt++] = (int32_t)i;
            }
            count_out[b * M + m] = (int32_t)cnt;
            for (; cnt < max_nbr; ++cnt) out[cnt] = -1;
        }
    }
    return 0;
}
