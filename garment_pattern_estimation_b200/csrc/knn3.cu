// kNN on xyz coordinates (D = 3): the graph of the first EdgeConv layer (reference nn/net_blocks.py:127-129,174 ->
// torch_cluster.knn on `pos`).  8 flop per pair, so the scan is bound by instruction issue and by the top-k bookkeeping, not by
// the distance arithmetic: one thread per query, the whole cloud staged in shared memory as float4 (x, y, z, -) and read with
// broadcast 16-byte loads, the k best kept sorted in registers.  Bit-exact with oracle/knn_oracle.c:
//   d = fmaf(dz, dz, fmaf(dy, dy, dx * dx)),  diff = candidate - query,  candidates in ascending index order,
//   a candidate enters only when it is STRICTLY closer than the current k-th best (ties keep the lower index first).
#include "common.cuh"

namespace nt {

constexpr int K3_THREADS = 128;
constexpr int K3_TILE = 2048;            // candidates staged per pass (32 KB)

template <int K>
__global__ void __launch_bounds__(K3_THREADS) knn3_kernel(const float *__restrict__ x, int N, int ldx, int k,
                                                          int32_t *__restrict__ idx) {
    __shared__ float4 cand[K3_TILE];
    const int b = blockIdx.y;
    const float *cloud = x + (int64_t)b * N * ldx;
    const int q = blockIdx.x * K3_THREADS + threadIdx.x;
    const bool q_ok = q < N;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (q_ok) { qx = cloud[(int64_t)q * ldx]; qy = cloud[(int64_t)q * ldx + 1]; qz = cloud[(int64_t)q * ldx + 2]; }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int e = 0; e < K; ++e) { bd[e] = 1e10f; bi[e] = -1; }

    for (int c0 = 0; c0 < N; c0 += K3_TILE) {
        const int cn = min(K3_TILE, N - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += K3_THREADS) {
            const float *pc = cloud + (int64_t)(c0 + i) * ldx;
            cand[i] = make_float4(pc[0], pc[1], pc[2], 0.f);
        }
        __syncthreads();
        auto dist = [&](const float4 p) {
            const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        };
        auto offer = [&](int c, float d) {
            if (d < bd[K - 1]) {
                bd[K - 1] = d; bi[K - 1] = c0 + c;
#pragma unroll
                for (int m = K - 1; m > 0; --m) {
                    if (bd[m - 1] > bd[m]) {            // strict: an equal, earlier candidate stays in front
                        const float td = bd[m - 1]; bd[m - 1] = bd[m]; bd[m] = td;
                        const int tix = bi[m - 1]; bi[m - 1] = bi[m]; bi[m] = tix;
                    }
                }
            }
        };
        int c = 0;
        for (; c + 4 <= cn; c += 4) {
            // four distances, ONE test against the current k-th best (the common case rejects all four); candidates that
            // pass are still offered one by one in index order, so the result is that of the sequential scan
            const float d0 = dist(cand[c]), d1 = dist(cand[c + 1]), d2 = dist(cand[c + 2]), d3 = dist(cand[c + 3]);
            if (fminf(fminf(d0, d1), fminf(d2, d3)) < bd[K - 1]) {
                offer(c, d0); offer(c + 1, d1); offer(c + 2, d2); offer(c + 3, d3);
            }
        }
        for (; c < cn; ++c) offer(c, dist(cand[c]));
    }
    if (q_ok) {
        int32_t *o = idx + ((int64_t)b * N + q) * k;
#pragma unroll
        for (int e = 0; e < K; ++e)
            if (e < k) o[e] = bi[e];
    }
}

template <int K>
static int launch_knn3(const float *x, int B, int N, int ldx, int k, int32_t *idx, cudaStream_t st) {
    dim3 grid((N + K3_THREADS - 1) / K3_THREADS, B);
    knn3_kernel<K><<<grid, K3_THREADS, 0, st>>>(x, N, ldx, k, idx);
    return check_launch("nt_knn(xyz)");
}

int knn3_run(const float *x, int B, int N, int ldx, int k, int32_t *idx, cudaStream_t st) {
    if (k <= 5) return launch_knn3<5>(x, B, N, ldx, k, idx, st);
    if (k <= 8) return launch_knn3<8>(x, B, N, ldx, k, idx, st);
    if (k <= 16) return launch_knn3<16>(x, B, N, ldx, k, idx, st);
    return launch_knn3<32>(x, B, N, ldx, k, idx, st);
}

}  // namespace nt
