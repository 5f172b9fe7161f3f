// PointNet++ set abstraction (reference nn/net_blocks.py:10-88: torch_geometric fps -> radius -> PointConv with max aggregation).
// SURVEY.md section 8 row a14.  Integer / index work on small clouds: one CTA per cloud with the cloud staged in shared memory.
//
//   nt_fps        farthest point sampling, deterministic start (point 0 of the cloud; torch_cluster random_start=False),
//                 squared distances with the same sequential fmaf chain as the kNN kernels -> bit-exact with oracle/knn_oracle.c
//   nt_radius     per centre: the first max_nbr points of its cloud (ascending index) with squared distance < r^2
//   nt_point_edges   edge list of PointConv in the reference's bipartite call, including the library's index-based self-loop
//                    handling (remove src == dst, append i -> i), and the message input pos_j - pos_i per edge
//   nt_scatter_max_fwd / _bwd   max aggregation over a general edge -> target map (first edge attaining the maximum wins)
#include "common.cuh"

namespace nt {

constexpr int FPS_THREADS = 512;

__device__ __forceinline__ float sqdist3(const float *a, const float *b, int D) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d) {
        const float diff = a[d] - b[d];
        acc = __fmaf_rn(diff, diff, acc);
    }
    return acc;
}

// one CTA per cloud; positions [N, D] and the running minimum distances in shared memory
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const float *__restrict__ pos, int ld, int N, int D, int n_samples,
                                                          int32_t *__restrict__ idx_out) {
    extern __shared__ float sm[];
    float *p = sm;                       // [N][D]
    float *mind = sm + (size_t)N * D;    // [N]
    __shared__ float red_v[FPS_THREADS / 32];
    __shared__ int red_i[FPS_THREADS / 32];
    __shared__ int cur_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cloud = pos + (int64_t)b * N * ld;
    for (int i = tid; i < N * D; i += FPS_THREADS) p[i] = cloud[(int64_t)(i / D) * ld + (i % D)];
    for (int i = tid; i < N; i += FPS_THREADS) mind[i] = INFINITY;
    if (tid == 0) cur_s = 0;
    __syncthreads();
    for (int s = 0; s < n_samples; ++s) {
        const int cur = cur_s;
        if (tid == 0) idx_out[(int64_t)b * n_samples + s] = cur;
        float best = -1.f;
        int best_i = 0x7fffffff;
        for (int i = tid; i < N; i += FPS_THREADS) {
            const float dd = sqdist3(p + (size_t)i * D, p + (size_t)cur * D, D);
            const float m = fminf(mind[i], dd);
            mind[i] = m;
            if (m > best) { best = m; best_i = i; }          // ascending i within a thread: first maximum kept
        }
        // arg-max with ties to the lowest index
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) { red_v[warp] = best; red_i[warp] = best_i; }
        __syncthreads();
        if (warp == 0) {
            best = lane < FPS_THREADS / 32 ? red_v[lane] : -2.f;
            best_i = lane < FPS_THREADS / 32 ? red_i[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
                if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
            }
            if (lane == 0) cur_s = best_i;
        }
        __syncthreads();
    }
}

// one thread per centre; the cloud is read through the read-only path (N * D floats per cloud, L1/L2 resident)
__global__ void radius_kernel(const float *__restrict__ pos, int ld, int B, int N, int D, const int32_t *__restrict__ centres, int M,
                              float r2, int max_nbr, int32_t *__restrict__ nbr, int32_t *__restrict__ count) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)B * M) return;
    const int b = (int)(q / M);
    const float *cloud = pos + (int64_t)b * N * ld;
    float c[8];
    const float *pc = cloud + (int64_t)centres[q] * ld;
    for (int d = 0; d < D; ++d) c[d] = __ldg(pc + d);
    int32_t *out = nbr + q * max_nbr;
    int cnt = 0;
    for (int i = 0; i < N && cnt < max_nbr; ++i) {
        float acc = 0.f;
        for (int d = 0; d < D; ++d) {
            const float diff = __ldg(cloud + (int64_t)i * ld + d) - c[d];
            acc = __fmaf_rn(diff, diff, acc);
        }
        if (acc < r2) out[cnt++] = i;
    }
    count[q] = cnt;
    for (int j = cnt; j < max_nbr; ++j) out[j] = -1;
}

// Edge list of PointConv (PyG PointNetConv.forward with add_self_loops=True on the bipartite pair (points, centres)):
// grouped radius edges minus those with source index == target index, then one edge (i -> i) per centre i < min(B*N, B*M).
// offsets[q] = exclusive prefix sum over centres of the kept radius edges (computed by the caller from `keep_count`).
__global__ void point_edges_count_kernel(const int32_t *__restrict__ nbr, const int32_t *__restrict__ count, int B, int N, int M,
                                         int max_nbr, int32_t *__restrict__ keep_count) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)B * M) return;
    const int b = (int)(q / M);
    int kept = 0;
    for (int j = 0; j < count[q]; ++j) kept += ((int64_t)b * N + nbr[q * max_nbr + j]) != q;
    keep_count[q] = kept;
}

__global__ void point_edges_fill_kernel(const float *__restrict__ pos, int ld, int D, const int32_t *__restrict__ centres,
                                        const int32_t *__restrict__ nbr, const int32_t *__restrict__ count,
                                        const int64_t *__restrict__ offsets, int B, int N, int M, int max_nbr, int64_t n_radius_edges,
                                        int64_t *__restrict__ src, int64_t *__restrict__ dst, float *__restrict__ msg, int ldm) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_centres = (int64_t)B * M, n_points = (int64_t)B * N;
    if (q >= n_centres) return;
    const int b = (int)(q / M);
    const float *pc = pos + ((int64_t)b * N + centres[q]) * ld;
    int64_t e = offsets[q];
    for (int j = 0; j < count[q]; ++j) {
        const int64_t s = (int64_t)b * N + nbr[q * max_nbr + j];
        if (s == q) continue;
        src[e] = s; dst[e] = q;
        for (int d = 0; d < D; ++d) msg[e * ldm + d] = pos[s * ld + d] - pc[d];
        ++e;
    }
    if (q < (n_points < n_centres ? n_points : n_centres)) {          // the appended "self loop": point q -> centre q
        const int64_t e2 = n_radius_edges + q;
        src[e2] = q; dst[e2] = q;
        for (int d = 0; d < D; ++d) msg[e2 * ldm + d] = pos[q * ld + d] - pc[d];
    }
}

// ---- scatter-max: out[t, f] = max over edges e with dst[e] = t of v[e, f]; arg[t, f] = the lowest such e attaining it -----------
__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void scatter_max_init_kernel(int *__restrict__ key, int64_t *__restrict__ arg, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = float_to_ordered(-INFINITY); arg[i] = INT64_MAX; }
}
__global__ void scatter_max_pass1_kernel(const float *__restrict__ v, int ldv, const int64_t *__restrict__ dst, int64_t E, int F,
                                         int *__restrict__ key) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * F) return;
    const int64_t e = i / F;
    const int f = (int)(i % F);
    atomicMax(key + dst[e] * F + f, float_to_ordered(v[e * ldv + f]));
}
__global__ void scatter_max_pass2_kernel(const float *__restrict__ v, int ldv, const int64_t *__restrict__ dst, int64_t E, int F,
                                         const int *__restrict__ key, int64_t *__restrict__ arg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * F) return;
    const int64_t e = i / F;
    const int f = (int)(i % F);
    if (float_to_ordered(v[e * ldv + f]) == key[dst[e] * F + f])
        atomicMin(reinterpret_cast<unsigned long long *>(arg + dst[e] * F + f), (unsigned long long)e);
}
__global__ void scatter_max_finish_kernel(const int *__restrict__ key, const int64_t *__restrict__ arg, int64_t n, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = arg[i] == INT64_MAX ? 0.f : ordered_to_float(key[i]);       // targets without edges -> 0 (PyG)
}
__global__ void scatter_max_bwd_kernel(const float *__restrict__ g, const int64_t *__restrict__ arg, int64_t T, int F,
                                       float *__restrict__ gv, int ldg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * F) return;
    const int64_t e = arg[i];
    if (e != INT64_MAX) gv[e * ldg + (i % F)] = g[i];          // one writer per (edge, feature): arg is unique per target
}

// ---- stage-2 input: all pairs of 3D edges from different panels (NNSewingPattern.all_edge_pairs, nn/data/pattern_converter.py:458) ----
// block b = panel pair (i < j) in the reference's loop order; pair p of the block = (row r of panel i, row c of panel j), r-major.
__global__ void edge_pairs_kernel(const float *__restrict__ edges, int Lmax, int F, const int32_t *__restrict__ blk_i,
                                  const int32_t *__restrict__ blk_j, const int32_t *__restrict__ blk_cols,
                                  const int64_t *__restrict__ blk_off, int n_blocks, int64_t n_pairs, float *__restrict__ pairs,
                                  int32_t *__restrict__ mapping) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pairs) return;
    int lo = 0, hi = n_blocks - 1;                       // last block with blk_off <= q
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (blk_off[mid] <= q) lo = mid; else hi = mid - 1;
    }
    const int i = blk_i[lo], j = blk_j[lo], cols = blk_cols[lo];
    const int64_t within = q - blk_off[lo];
    const int r = (int)(within / cols), c = (int)(within % cols);
    const float *ei = edges + ((int64_t)i * Lmax + r) * F, *ej = edges + ((int64_t)j * Lmax + c) * F;
    float *o = pairs + q * 2 * F;
    for (int f = 0; f < F; ++f) { o[f] = ei[f]; o[F + f] = ej[f]; }
    if (mapping) { mapping[q * 4] = i; mapping[q * 4 + 1] = r; mapping[q * 4 + 2] = j; mapping[q * 4 + 3] = c; }
}

}  // namespace nt

using namespace nt;

extern "C" int nt_edge_pairs(const float *edges, int Lmax, int F, const int32_t *blk_i, const int32_t *blk_j, const int32_t *blk_cols,
                             const int64_t *blk_off, int n_blocks, int64_t n_pairs, float *pairs, int32_t *mapping, void *stream) {
    NT_REQUIRE(edges && blk_i && blk_j && blk_cols && blk_off && pairs && Lmax >= 1 && F >= 1 && n_blocks >= 1 && n_pairs >= 1,
               "nt_edge_pairs: bad arguments");
    edge_pairs_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        edges, Lmax, F, blk_i, blk_j, blk_cols, blk_off, n_blocks, n_pairs, pairs, mapping);
    return check_launch("nt_edge_pairs");
}

extern "C" int nt_fps(const float *pos, int ld, int B, int N, int D, int n_samples, int32_t *idx, void *stream) {
    NT_REQUIRE(pos && idx && B >= 1 && N >= 1 && D >= 1 && D <= 8 && ld >= D, "nt_fps: bad arguments");
    NT_REQUIRE(n_samples >= 1 && n_samples <= N, "nt_fps: need 1 <= n_samples <= N");
    const size_t smem = ((size_t)N * D + N) * sizeof(float);
    NT_REQUIRE(smem <= 200 * 1024, "nt_fps: cloud too large for the shared-memory kernel (N * (D + 1) floats must fit 200 KB)");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return fail("nt_fps: cudaFuncSetAttribute failed%s", "");
        configured = true;
    }
    fps_kernel<<<B, FPS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(pos, ld, N, D, n_samples, idx);
    return check_launch("nt_fps");
}

extern "C" int nt_radius(const float *pos, int ld, int B, int N, int D, const int32_t *centres, int M, float r, int max_nbr,
                         int32_t *nbr, int32_t *count, void *stream) {
    NT_REQUIRE(pos && centres && nbr && count && B >= 1 && N >= 1 && D >= 1 && D <= 8 && ld >= D && M >= 1 && max_nbr >= 1,
               "nt_radius: bad arguments");
    const int64_t total = (int64_t)B * M;
    radius_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pos, ld, B, N, D, centres, M, r * r,
                                                                                                     max_nbr, nbr, count);
    return check_launch("nt_radius");
}

extern "C" int nt_point_edges_count(const int32_t *nbr, const int32_t *count, int B, int N, int M, int max_nbr, int32_t *keep_count,
                                    void *stream) {
    NT_REQUIRE(nbr && count && keep_count && B >= 1 && N >= 1 && M >= 1 && max_nbr >= 1, "nt_point_edges_count: bad arguments");
    const int64_t total = (int64_t)B * M;
    point_edges_count_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(nbr, count, B, N, M,
                                                                                                                max_nbr, keep_count);
    return check_launch("nt_point_edges_count");
}

extern "C" int nt_point_edges_fill(const float *pos, int ld, int D, const int32_t *centres, const int32_t *nbr, const int32_t *count,
                                   const int64_t *offsets, int B, int N, int M, int max_nbr, int64_t n_radius_edges, int64_t *src,
                                   int64_t *dst, float *msg, int ldm, void *stream) {
    NT_REQUIRE(pos && centres && nbr && count && offsets && src && dst && msg && ldm >= D && D >= 1 && D <= 8,
               "nt_point_edges_fill: bad arguments");
    const int64_t total = (int64_t)B * M;
    point_edges_fill_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        pos, ld, D, centres, nbr, count, offsets, B, N, M, max_nbr, n_radius_edges, src, dst, msg, ldm);
    return check_launch("nt_point_edges_fill");
}

extern "C" int nt_scatter_max_fwd(const float *v, int ldv, const int64_t *dst, int64_t E, int F, int64_t T, float *out, int64_t *arg,
                                  int32_t *key_scratch, void *stream) {
    NT_REQUIRE(v && dst && out && arg && key_scratch && E >= 0 && F >= 1 && T >= 1 && ldv >= F, "nt_scatter_max_fwd: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t n = T * F;
    scatter_max_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key_scratch, arg, n);
    if (int rc = check_launch("nt_scatter_max_fwd(init)")) return rc;
    if (E > 0) {
        scatter_max_pass1_kernel<<<(unsigned)((E * F + 255) / 256), 256, 0, st>>>(v, ldv, dst, E, F, key_scratch);
        if (int rc = check_launch("nt_scatter_max_fwd(max)")) return rc;
        scatter_max_pass2_kernel<<<(unsigned)((E * F + 255) / 256), 256, 0, st>>>(v, ldv, dst, E, F, key_scratch, arg);
        if (int rc = check_launch("nt_scatter_max_fwd(arg)")) return rc;
    }
    scatter_max_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key_scratch, arg, n, out);
    return check_launch("nt_scatter_max_fwd(finish)");
}

extern "C" int nt_scatter_max_bwd(const float *g, const int64_t *arg, int64_t T, int F, float *gv, int ldg, void *stream) {
    NT_REQUIRE(g && arg && gv && T >= 1 && F >= 1 && ldg >= F, "nt_scatter_max_bwd: bad arguments (gv must be zeroed by the caller)");
    const int64_t n = T * F;
    scatter_max_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, arg, T, F, gv, ldg);
    return check_launch("nt_scatter_max_bwd");
}
