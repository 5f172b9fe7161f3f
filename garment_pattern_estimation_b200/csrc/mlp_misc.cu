// BatchNorm bookkeeping, aggregation finish, and the element-wise / column-reduction pieces of the MLP() backward
// (reference nn/net_blocks.py:43-47: Linear -> ReLU -> BatchNorm1d, BN after ReLU and also on the last layer).
// All of these are bandwidth-trivial next to the GEMMs; they exist so that no [rows, C] tensor makes an extra trip
// through HBM for BN normalisation (the affine is folded into the next Linear's weights instead).
#include "common.cuh"

namespace nt {

// ---------------------------------------------------------------------------------------------------------
// bn_fold: statistics -> (mean, rstd, s, t), running-buffer update, fold into the next Linear.
// grid = n_next + 1 blocks; block b < n_next folds output row b, the last block writes the per-channel vectors.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bn_channel(const double *stats, int64_t count, int C, int c, const float *gamma,
                                           const float *beta, const float *rmean, const float *rvar, float eps,
                                           int training, float &mean, float &var_b, float &rstd, float &s, float &t) {
    if (training) {
        double m = stats[c] / (double)count;
        double v = stats[C + c] / (double)count - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var_b = (float)v;
        rstd = (float)(1.0 / sqrt(v + (double)eps));
    } else {
        mean = rmean[c]; var_b = rvar[c];
        rstd = 1.0f / sqrtf(var_b + eps);
    }
    s = gamma[c] * rstd;
    t = beta[c] - mean * s;
}

__global__ void bn_fold_kernel(const double *stats, int64_t count, int C, const float *gamma, const float *beta,
                               float *rmean, float *rvar, int64_t *nbt, float momentum, float eps, int training,
                               float *mean_o, float *rstd_o, float *s_o, float *t_o,
                               const float *w_next, const float *b_next, int n_next,
                               float *w_f, float *w_ft, float *b_f) {
    __shared__ float red[32];
    const int o = blockIdx.x;
    if (o < n_next) {
        float part = 0.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float mean, var_b, rstd, s, t;
            bn_channel(stats, count, C, c, gamma, beta, rmean, rvar, eps, training, mean, var_b, rstd, s, t);
            float w = w_next[(int64_t)o * C + c];
            float wf = w * s;
            w_f[(int64_t)o * C + c] = wf;
            if (w_ft) w_ft[(int64_t)c * n_next + o] = wf;
            part = fmaf(w, t, part);
        }
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += red[i];
            b_f[o] = (b_next ? b_next[o] : 0.f) + tot;
        }
    } else {
        // NOTE: this block READS the running buffers before updating them; the fold blocks above only read them in
        // eval mode, where they are not written.
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float mean, var_b, rstd, s, t;
            bn_channel(stats, count, C, c, gamma, beta, rmean, rvar, eps, training, mean, var_b, rstd, s, t);
            mean_o[c] = mean; rstd_o[c] = rstd; s_o[c] = s; t_o[c] = t;
            if (training && rmean && rvar) {
                float unbiased = count > 1 ? var_b * ((float)count / (float)(count - 1)) : var_b;
                rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
                rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
            }
        }
        if (training && nbt && threadIdx.x == 0) *nbt += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void maxmin_finish_kernel(const float *vmax, const float *vmin, const uint8_t *imax, const uint8_t *imin,
                                     const float *s, const float *t, int64_t M, int C, float *out, int ldo,
                                     uint8_t *sel, float *vsel, const float *tail_src, int tail_ld, int tail) {
    const int W = C + tail;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * W) return;
    int64_t m = i / W;
    int c = (int)(i - m * W);
    if (c < C) {
        float sc = s[c];
        bool up = sc >= 0.f;
        int64_t o = m * C + c;
        float v = up ? vmax[o] : vmin[o];
        out[m * ldo + c] = fmaf(v, sc, t[c]);
        if (sel) sel[o] = up ? imax[o] : imin[o];
        if (vsel) vsel[o] = v;
    } else {
        out[m * ldo + c] = tail_src[m * tail_ld + (c - C)];
    }
}

__global__ void bn_apply_kernel(const float *a, int lda, const float *s, const float *t, int64_t rows, int C,
                                float *out, int ldo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    int64_t r = i / C;
    int c = (int)(i - r * C);
    out[r * ldo + c] = fmaf(a[r * lda + c], s[c], t[c]);
}

// ---------------------------------------------------------------------------------------------------------
// column reductions: block (32, 8); blockIdx.y selects a 32-column slab, blockIdx.x a row chunk
// ---------------------------------------------------------------------------------------------------------
constexpr int CR_ROWS = 512;   // rows per block

__global__ void bn_bwd_reduce_kernel(const float *g, int ldg, const float *v, int ldv, const float *mu,
                                     const float *rstd, int64_t rows, int C, double *sums) {
    __shared__ float r0[8][33], r1[8][33];
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int64_t rbeg = (int64_t)blockIdx.x * CR_ROWS;
    const int64_t rend = min(rows, rbeg + CR_ROWS);
    float a0 = 0.f, a1 = 0.f;
    if (c < C) {
        const float m = v ? mu[c] : 0.f, rs = v ? rstd[c] : 0.f;
        for (int64_t r = rbeg + threadIdx.y; r < rend; r += 8) {
            float gv = g[r * ldg + c];
            a0 += gv;
            if (v) a1 = fmaf(gv, (v[r * ldv + c] - m) * rs, a1);
        }
    }
    r0[threadIdx.y][threadIdx.x] = a0; r1[threadIdx.y][threadIdx.x] = a1;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0 += r0[i][threadIdx.x]; s1 += r1[i][threadIdx.x]; }
        atomicAdd(sums + c, (double)s0);
        if (v) atomicAdd(sums + C + c, (double)s1);
    }
}

// statistics of the FIRST activation a1 = relu(P[centre] + Q[neighbour]) (no GEMM in front of it at edge level)
__global__ void edge_stats_kernel(const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                                  int64_t rows, int H, double *stats) {
    __shared__ float r0[8][33], r1[8][33];
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int64_t rbeg = (int64_t)blockIdx.x * CR_ROWS;
    const int64_t rend = min(rows, rbeg + CR_ROWS);
    float a0 = 0.f, a1 = 0.f;
    if (c < H) {
        for (int64_t r = rbeg + threadIdx.y; r < rend; r += 8) {
            float v;
            if (idx) {
                const int64_t centre = r / k;
                const int64_t j = (centre / n_per_cloud) * (int64_t)n_per_cloud + idx[r];
                v = pq[centre * ldpq + c] + pq[j * ldpq + qoff + c];
            } else {
                v = pq[r * ldpq + c];
            }
            v = fmaxf(v, 0.f);
            a0 += v;
            a1 = fmaf(v, v, a1);
        }
    }
    r0[threadIdx.y][threadIdx.x] = a0; r1[threadIdx.y][threadIdx.x] = a1;
    __syncthreads();
    if (threadIdx.y == 0 && c < H) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0 += r0[i][threadIdx.x]; s1 += r1[i][threadIdx.x]; }
        atomicAdd(stats + c, (double)s0);
        atomicAdd(stats + H + c, (double)s1);
    }
}

// Materialises the first edge activation a1[e, :] = relu(P[centre(e), :] + Q[nbr(e), :]) (float4 per thread, block = 32
// column quads x 8 row lanes) and accumulates its BatchNorm statistics on the way.  Every later consumer (the next GEMM's
// A operand, the weight-gradient GEMM, the BN/ReLU backward epilogue) then streams a1 instead of re-gathering two random
// rows of PQ per edge (ncu, round 1: the gathered variants of those kernels ran 2x slower than the plain ones).
constexpr int EA_ROWS = 128;   // rows per block

__global__ void __launch_bounds__(256) edge_activation_kernel(const float *__restrict__ pq, int ldpq, int qoff,
                                                              const int32_t *__restrict__ idx, int k, int n_per_cloud,
                                                              int64_t rows, int H, float *__restrict__ out, int ldo,
                                                              double *__restrict__ stats) {
    __shared__ float red[8][32][8];
    const int c0 = (blockIdx.y * 32 + threadIdx.x) * 4;
    const int64_t rbeg = (int64_t)blockIdx.x * EA_ROWS;
    const int64_t rend = min(rows, rbeg + EA_ROWS);
    const bool vec = ((ldpq & 3) == 0) && ((qoff & 3) == 0) && ((ldo & 3) == 0) && aligned16(pq) && aligned16(out) && c0 + 3 < H;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < H) {
        for (int64_t r = rbeg + threadIdx.y; r < rend; r += 8) {
            // idx == nullptr: plain rows, a_1 = relu(pq[r]) (no neighbour term)
            const int64_t centre = idx ? r / k : r;
            const int64_t j = idx ? (centre / n_per_cloud) * (int64_t)n_per_cloud + __ldg(idx + r) : 0;
            const float *pp = pq + centre * ldpq + c0, *qp = idx ? pq + j * ldpq + qoff + c0 : nullptr;
            float v[4];
            if (vec) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(pp));
                const float4 b = qp ? __ldg(reinterpret_cast<const float4 *>(qp)) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[0] = fmaxf(a.x + b.x, 0.f); v[1] = fmaxf(a.y + b.y, 0.f); v[2] = fmaxf(a.z + b.z, 0.f); v[3] = fmaxf(a.w + b.w, 0.f);
                *reinterpret_cast<float4 *>(out + r * ldo + c0) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[e] = (c0 + e < H) ? fmaxf(__ldg(pp + e) + (qp ? __ldg(qp + e) : 0.f), 0.f) : 0.f;
                    if (c0 + e < H) out[r * ldo + c0 + e] = v[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) { s1[e] += v[e]; s2[e] = fmaf(v[e], v[e], s2[e]); }
        }
    }
    if (!stats) return;
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[threadIdx.y][threadIdx.x][e] = s1[e]; red[threadIdx.y][threadIdx.x][4 + e] = s2[e]; }
    __syncthreads();
    if (threadIdx.y == 0 && c0 < H) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (c0 + e >= H) break;
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { t1 += red[i][threadIdx.x][e]; t2 += red[i][threadIdx.x][4 + e]; }
            atomicAdd(stats + c0 + e, (double)t1);
            atomicAdd(stats + H + c0 + e, (double)t2);
        }
    }
}

// Each thread owns 4 consecutive columns (one float4 when the row stride allows it) and walks its row chunk 8 rows apart;
// block = (32 column-quads, 8 row lanes) -> a warp reads 512 contiguous bytes of a row.
constexpr int BL_ROWS = 256;   // rows per block

__global__ void __launch_bounds__(256) bn_relu_bwd_last_kernel(const float *__restrict__ a, int lda, const float *__restrict__ g,
                                                               int ldg, const uint8_t *__restrict__ sel, int k,
                                                               const float *__restrict__ s, const float *__restrict__ mu,
                                                               const float *__restrict__ rstd, const double *__restrict__ sums,
                                                               int64_t count, int64_t rows, int C, float *__restrict__ dz,
                                                               int lddz, double *__restrict__ colsum) {
    __shared__ float red[8][32][4];
    const int c0 = (blockIdx.y * 32 + threadIdx.x) * 4;
    const int64_t rbeg = (int64_t)blockIdx.x * BL_ROWS;
    const int64_t rend = min(rows, rbeg + BL_ROWS);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
        const float inv = 1.0f / (float)count;
        float sc[4], m[4], rs[4], dbeta[4], dgamma[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = min(c0 + j, C - 1);
            sc[j] = s[c]; m[j] = mu[c]; rs[j] = rstd[c];
            dbeta[j] = (float)sums[c] * inv; dgamma[j] = (float)sums[C + c] * inv;
        }
        const bool vec = (c0 + 3 < C) && ((lda & 3) == 0) && ((lddz & 3) == 0) && aligned16(a) && aligned16(dz);
        for (int64_t r = rbeg + threadIdx.y; r < rend; r += 8) {
            float av[4];
            if (vec) {
                const float4 t = *reinterpret_cast<const float4 *>(a + r * lda + c0);
                av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) av[j] = (c0 + j < C) ? a[r * lda + c0 + j] : 0.f;
            }
            const int64_t node = r / k;
            const uint8_t slot = (uint8_t)(r - node * k);
            float out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                out[j] = 0.f;
                if (c0 + j < C && av[j] > 0.f) {
                    float gs = 0.f;
                    if (!sel || sel[node * C + c0 + j] == slot) gs = g[node * ldg + c0 + j];
                    out[j] = sc[j] * (gs - dbeta[j] - (av[j] - m[j]) * rs[j] * dgamma[j]);
                }
                acc[j] += out[j];
            }
            if (vec) {
                *reinterpret_cast<float4 *>(dz + r * lddz + c0) = make_float4(out[0], out[1], out[2], out[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < C) dz[r * lddz + c0 + j] = out[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j >= C) continue;
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x][j];
            atomicAdd(colsum + c0 + j, (double)t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// one WARP per input channel c: lanes stride over the n_out output rows, double partial sums reduced by shuffles
__global__ void __launch_bounds__(256) linear_bn_bwd_kernel(const double *__restrict__ rawc, const double *__restrict__ csum,
                                                            int n_out, int C, const float *__restrict__ w,
                                                            const float *__restrict__ s, const float *__restrict__ beta,
                                                            const float *__restrict__ rstd, int64_t count, float *__restrict__ dW,
                                                            float *__restrict__ db, float *__restrict__ dgamma,
                                                            float *__restrict__ dbeta, float *__restrict__ k0, float *__restrict__ k1) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (blockIdx.x == 0) {
        for (int o = threadIdx.x; o < n_out; o += blockDim.x) db[o] = (float)csum[o];
    }
    if (c >= C) return;
    const double sc = s[c], bc = beta[c];
    double db_acc = 0.0, dg_acc = 0.0;
    for (int o = lane; o < n_out; o += 32) {
        const double cs = csum[o];
        const double rw = rawc[(int64_t)o * C + c];
        const double wv = w[(int64_t)o * C + c];
        // y_prev = a*s + t with a = (a - mean) + mean  =>  dW = s * rawc + csum * (s*mean + t) = s * rawc + csum * beta
        dW[(int64_t)o * C + c] = (float)(rw * sc + cs * bc);
        db_acc += wv * cs;
        dg_acc += wv * rw;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        db_acc += __shfl_xor_sync(0xffffffffu, db_acc, off);
        dg_acc += __shfl_xor_sync(0xffffffffu, dg_acc, off);
    }
    if (lane == 0) {
        dg_acc *= (double)rstd[c];
        dbeta[c] = (float)db_acc; dgamma[c] = (float)dg_acc;
        const double inv = 1.0 / (double)count;
        k0[c] = (float)(sc * db_acc * inv);
        k1[c] = (float)(sc * (double)rstd[c] * dg_acc * inv);
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void edge_scatter_kernel(const float *dz, int lddz, const int32_t *idx, int k, int n_per_cloud,
                                    int64_t M, int H, float *dpq, int lddpq) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * H) return;
    int64_t m = i / H;
    int h = (int)(i - m * H);
    const int64_t base = (m / n_per_cloud) * (int64_t)n_per_cloud;
    float sum = 0.f;
    for (int sl = 0; sl < k; ++sl) {
        const int64_t e = m * k + sl;
        const float v = dz[e * lddz + h];
        sum += v;
        if (v != 0.f) atomicAdd(dpq + (base + idx[e]) * lddpq + H + h, v);
    }
    dpq[m * lddpq + h] = sum;
}

// ---------------------------------------------------------------------------------------------------------
// Node-centric float4 versions of the three edge-sized element-wise kernels (used when every row is 16-byte aligned).
// A thread owns 4 consecutive columns of one centre point and walks its k edge rows with up to 4 independent 16-byte
// loads in flight; per-node operands (centre row of PQ, upstream gradient, selected slots) are read once instead of k
// times and no 64-bit division is left in the loops.  Block = (32 column quads, 8 node lanes).
// Block = (column quads of a row, node lanes): when a row has at most 64 quads the block's x extent IS the number of quads (every lane
// live: H = 200 -> 50 x 5 threads, C = 150 -> 38 x 6), otherwise 32 quads per block and blockIdx.y walks the row (the first
// version always did: 22 % / 41 % of the lanes idle at H = 200 / C = 150).  npb = nodes per block.

__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__global__ void __launch_bounds__(256) edge_activation_v4_kernel(const float *__restrict__ pq, int ldpq, int qoff,
                                                                 const int32_t *__restrict__ idx, int k, int n_per_cloud,
                                                                 int64_t nodes, int H, float *__restrict__ out, int ldo,
                                                                 double *__restrict__ stats, int npb) {
    __shared__ float red[8][64][8];
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    const int64_t nbeg = (int64_t)blockIdx.x * npb;
    const int64_t nend = min(nodes, nbeg + npb);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < H) {
        for (int64_t node = nbeg + threadIdx.y; node < nend; node += blockDim.y) {
            const float4 pc = ld4(pq + node * ldpq + c0);
            const int64_t base = (node / n_per_cloud) * (int64_t)n_per_cloud;
            const int32_t *ip = idx + node * k;
            float *op = out + node * k * (int64_t)ldo + c0;
            for (int s0 = 0; s0 < k; s0 += 4) {
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (s0 + u < k) q[u] = ld4(pq + (base + __ldg(ip + s0 + u)) * ldpq + qoff + c0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (s0 + u < k) {
                        float4 v;
                        v.x = fmaxf(pc.x + q[u].x, 0.f); v.y = fmaxf(pc.y + q[u].y, 0.f);
                        v.z = fmaxf(pc.z + q[u].z, 0.f); v.w = fmaxf(pc.w + q[u].w, 0.f);
                        *reinterpret_cast<float4 *>(op + (int64_t)(s0 + u) * ldo) = v;
                        s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
                        s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
                        s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
                    }
                }
            }
        }
    }
    if (!stats) return;
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[threadIdx.y][threadIdx.x][e] = s1[e]; red[threadIdx.y][threadIdx.x][4 + e] = s2[e]; }
    __syncthreads();
    if (threadIdx.y == 0 && c0 < H) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t1 = 0.f, t2 = 0.f;
            for (int i = 0; i < (int)blockDim.y; ++i) { t1 += red[i][threadIdx.x][e]; t2 += red[i][threadIdx.x][4 + e]; }
            atomicAdd(stats + c0 + e, (double)t1);
            atomicAdd(stats + H + c0 + e, (double)t2);
        }
    }
}

__global__ void __launch_bounds__(256) bn_relu_bwd_last_v4_kernel(const float *__restrict__ a, int lda, const float *__restrict__ g,
                                                                  int ldg, const uint8_t *__restrict__ sel, int k,
                                                                  const float *__restrict__ s, const float *__restrict__ mu,
                                                                  const float *__restrict__ rstd, const double *__restrict__ sums,
                                                                  int64_t count, int64_t nodes, int C, float *__restrict__ dz,
                                                                  int lddz, double *__restrict__ colsum, int npb) {
    __shared__ float red[8][64][4];
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;     // C % 4 == 0 is NOT required: the tail quad is masked per column
    const int64_t nbeg = (int64_t)blockIdx.x * npb;
    const int64_t nend = min(nodes, nbeg + npb);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
        const float inv = 1.0f / (float)count;
        float sc[4], k0[4], k1[4], m[4];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = min(c0 + j, C - 1);
            ok[j] = c0 + j < C;
            sc[j] = s[c]; m[j] = mu[c];
            k0[j] = (float)sums[c] * inv;                        // dbeta / count
            k1[j] = rstd[c] * ((float)sums[C + c] * inv);        // rstd * dgamma / count
        }
        for (int64_t node = nbeg + threadIdx.y; node < nend; node += blockDim.y) {
            float gv[4];                                         // upstream gradient rows may have any stride: scalar loads
#pragma unroll
            for (int j = 0; j < 4; ++j) gv[j] = ok[j] ? __ldg(g + node * ldg + c0 + j) : 0.f;
            int slot_sel[4] = {-1, -1, -1, -1};                  // -1: every slot receives the gradient (no aggregation)
            if (sel) {
#pragma unroll
                for (int j = 0; j < 4; ++j) slot_sel[j] = ok[j] ? (int)sel[node * C + c0 + j] : 0;
            }
            const float *ap = a + node * k * (int64_t)lda + c0;
            float *op = dz + node * k * (int64_t)lddz + c0;
            for (int s0 = 0; s0 < k; s0 += 4) {
                float4 av[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (s0 + u < k) av[u] = ld4(ap + (int64_t)(s0 + u) * lda);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (s0 + u < k) {
                        const float x[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
                        float o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float gs = (slot_sel[j] < 0 || slot_sel[j] == s0 + u) ? gv[j] : 0.f;
                            // same association as the scalar kernel: sc * (gs - dbeta' - ((x - m) * rstd) * dgamma')
                            o[j] = (ok[j] && x[j] > 0.f) ? sc[j] * (gs - k0[j] - (x[j] - m[j]) * k1[j]) : 0.f;
                            acc[j] += o[j];
                        }
                        *reinterpret_cast<float4 *>(op + (int64_t)(s0 + u) * lddz) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j >= C) continue;
            float t = 0.f;
            for (int i = 0; i < (int)blockDim.y; ++i) t += red[i][threadIdx.x][j];
            atomicAdd(colsum + c0 + j, (double)t);
        }
    }
}

// ---- two nodes per thread and iteration, every load of both nodes issued before the first use (k <= 8) ----------------------
// The one-node loops above expose three dependent memory latencies per iteration (index -> gathered row -> next node); ncu showed
// both kernels latency-bound (long scoreboard, 26-44 % of the DRAM peak).  Here a thread keeps up to 2 x (1 + k) 16-byte loads in
// flight.  (Instantiated for k <= 5, the shipped k_neighbors; larger k keeps the one-node kernels.)

template <int NX_K>
__global__ void __launch_bounds__(256, 2) edge_activation_x2_kernel(const float *__restrict__ pq, int ldpq, int qoff,
                                                                 const int32_t *__restrict__ idx, int k, int n_per_cloud,
                                                                 int64_t nodes, int H, float *__restrict__ out, int ldo,
                                                                 double *__restrict__ stats, int npb) {
    __shared__ float red[8][64][8];
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    const int64_t nbeg = (int64_t)blockIdx.x * npb;
    const int64_t nend = min(nodes, nbeg + npb);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < H) {
        for (int64_t nodeA = nbeg + threadIdx.y; nodeA < nend; nodeA += 2 * blockDim.y) {
            const int64_t nodeB = nodeA + blockDim.y;
            const bool hasB = nodeB < nend;
            const int32_t *ipA = idx + nodeA * k, *ipB = idx + (hasB ? nodeB : nodeA) * k;
            int jA[NX_K], jB[NX_K];
#pragma unroll
            for (int u = 0; u < NX_K; ++u)
                if (u < k) { jA[u] = __ldg(ipA + u); jB[u] = __ldg(ipB + u); }
            const float4 pcA = ld4(pq + nodeA * ldpq + c0);
            const float4 pcB = ld4(pq + (hasB ? nodeB : nodeA) * ldpq + c0);
            const int64_t baseA = (nodeA / n_per_cloud) * (int64_t)n_per_cloud;
            const int64_t baseB = ((hasB ? nodeB : nodeA) / n_per_cloud) * (int64_t)n_per_cloud;
            float4 qA[NX_K], qB[NX_K];
#pragma unroll
            for (int u = 0; u < NX_K; ++u)
                if (u < k) {
                    qA[u] = ld4(pq + (baseA + jA[u]) * ldpq + qoff + c0);
                    qB[u] = ld4(pq + (baseB + jB[u]) * ldpq + qoff + c0);
                }
            float *opA = out + nodeA * k * (int64_t)ldo + c0;
            float *opB = out + nodeB * k * (int64_t)ldo + c0;
#pragma unroll
            for (int u = 0; u < NX_K; ++u)
                if (u < k) {
                    float4 v;
                    v.x = fmaxf(pcA.x + qA[u].x, 0.f); v.y = fmaxf(pcA.y + qA[u].y, 0.f);
                    v.z = fmaxf(pcA.z + qA[u].z, 0.f); v.w = fmaxf(pcA.w + qA[u].w, 0.f);
                    *reinterpret_cast<float4 *>(opA + (int64_t)u * ldo) = v;
                    s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
                    s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
                    s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
                    if (hasB) {
                        v.x = fmaxf(pcB.x + qB[u].x, 0.f); v.y = fmaxf(pcB.y + qB[u].y, 0.f);
                        v.z = fmaxf(pcB.z + qB[u].z, 0.f); v.w = fmaxf(pcB.w + qB[u].w, 0.f);
                        *reinterpret_cast<float4 *>(opB + (int64_t)u * ldo) = v;
                        s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
                        s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
                        s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
                    }
                }
        }
    }
    if (!stats) return;
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[threadIdx.y][threadIdx.x][e] = s1[e]; red[threadIdx.y][threadIdx.x][4 + e] = s2[e]; }
    __syncthreads();
    if (threadIdx.y == 0 && c0 < H) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t1 = 0.f, t2 = 0.f;
            for (int i = 0; i < (int)blockDim.y; ++i) { t1 += red[i][threadIdx.x][e]; t2 += red[i][threadIdx.x][4 + e]; }
            atomicAdd(stats + c0 + e, (double)t1);
            atomicAdd(stats + H + c0 + e, (double)t2);
        }
    }
}

template <int NX_K>
__global__ void __launch_bounds__(256, 2) bn_relu_bwd_last_x2_kernel(const float *__restrict__ a, int lda, const float *__restrict__ g,
                                                                  int ldg, const uint8_t *__restrict__ sel, int k,
                                                                  const float *__restrict__ s, const float *__restrict__ mu,
                                                                  const float *__restrict__ rstd, const double *__restrict__ sums,
                                                                  int64_t count, int64_t nodes, int C, float *__restrict__ dz,
                                                                  int lddz, double *__restrict__ colsum, int npb) {
    __shared__ float red[8][64][4];
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    const int64_t nbeg = (int64_t)blockIdx.x * npb;
    const int64_t nend = min(nodes, nbeg + npb);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
        const float inv = 1.0f / (float)count;
        float sc[4], k0[4], k1[4], m[4];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = min(c0 + j, C - 1);
            ok[j] = c0 + j < C;
            sc[j] = s[c]; m[j] = mu[c];
            k0[j] = (float)sums[c] * inv;
            k1[j] = rstd[c] * ((float)sums[C + c] * inv);
        }
        for (int64_t nodeA = nbeg + threadIdx.y; nodeA < nend; nodeA += 2 * blockDim.y) {
            const bool hasB = nodeA + blockDim.y < nend;
            const int64_t nodeB = hasB ? nodeA + blockDim.y : nodeA;   // loads of a missing node B alias node A (never stored)
            float gv[2][4];
            int ss[2][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                gv[0][j] = ok[j] ? __ldg(g + nodeA * ldg + c0 + j) : 0.f;
                gv[1][j] = ok[j] ? __ldg(g + nodeB * ldg + c0 + j) : 0.f;
                ss[0][j] = ss[1][j] = -1;                            // -1: every slot receives the gradient (no aggregation)
                if (sel) {
                    ss[0][j] = ok[j] ? (int)sel[nodeA * C + c0 + j] : 0;
                    ss[1][j] = ok[j] ? (int)sel[nodeB * C + c0 + j] : 0;
                }
            }
            float4 av[2][NX_K];
#pragma unroll
            for (int u = 0; u < NX_K; ++u)
                if (u < k) {
                    av[0][u] = ld4(a + (nodeA * k + u) * (int64_t)lda + c0);
                    av[1][u] = ld4(a + (nodeB * k + u) * (int64_t)lda + c0);
                }
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                if (w == 1 && !hasB) break;
                float *op = dz + (w == 0 ? nodeA : nodeB) * k * (int64_t)lddz + c0;
#pragma unroll
                for (int u = 0; u < NX_K; ++u)
                    if (u < k) {
                        const float x[4] = {av[w][u].x, av[w][u].y, av[w][u].z, av[w][u].w};
                        float o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float gs = (ss[w][j] < 0 || ss[w][j] == u) ? gv[w][j] : 0.f;
                            // same association as the scalar kernel: sc * (gs - dbeta' - ((x - m) * rstd) * dgamma')
                            o[j] = (ok[j] && x[j] > 0.f) ? sc[j] * (gs - k0[j] - (x[j] - m[j]) * k1[j]) : 0.f;
                            acc[j] += o[j];
                        }
                        *reinterpret_cast<float4 *>(op + (int64_t)u * lddz) = make_float4(o[0], o[1], o[2], o[3]);
                    }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j >= C) continue;
            float t = 0.f;
            for (int i = 0; i < (int)blockDim.y; ++i) t += red[i][threadIdx.x][j];
            atomicAdd(colsum + c0 + j, (double)t);
        }
    }
}

__global__ void __launch_bounds__(256) edge_scatter_v4_kernel(const float *__restrict__ dz, int lddz, const int32_t *__restrict__ idx,
                                                              int k, int n_per_cloud, int64_t M, int H, float *__restrict__ dpq,
                                                              int lddpq, int npb) {
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if (c0 >= H) return;
    const int64_t nbeg = (int64_t)blockIdx.x * npb;
    const int64_t nend = min(M, nbeg + npb);
    for (int64_t node = nbeg + threadIdx.y; node < nend; node += blockDim.y) {
        const int64_t base = (node / n_per_cloud) * (int64_t)n_per_cloud;
        const int32_t *ip = idx + node * k;
        const float *zp = dz + node * k * (int64_t)lddz + c0;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < k; s0 += 4) {
            float4 v[4];
            int j[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (s0 + u < k) { v[u] = ld4(zp + (int64_t)(s0 + u) * lddz); j[u] = __ldg(ip + s0 + u); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s0 + u < k) {
                    sum.x += v[u].x; sum.y += v[u].y; sum.z += v[u].z; sum.w += v[u].w;
                    if (v[u].x != 0.f || v[u].y != 0.f || v[u].z != 0.f || v[u].w != 0.f)
                        atomicAdd(reinterpret_cast<float4 *>(dpq + (base + j[u]) * lddpq + H + c0), v[u]);
                }
            }
        }
        *reinterpret_cast<float4 *>(dpq + node * lddpq + c0) = sum;
    }
}

static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }
// launch geometry of the quad kernels above for rows of `cols` columns
// Nodes per block: at least four per node lane, and for large inputs enough that the grid stays at ~4 blocks per SM -- every block ends
// with one double atomicAdd per column (BatchNorm statistics / column sums), and those serialise per address in L2: at C2 the first
// version ran 3277 blocks = 3277 atomics on each of 400 addresses per launch.
static inline void quad_geometry(int cols, int64_t nodes, dim3 &block, unsigned &grid_y, int &npb) {
    const int quads = (cols + 3) / 4;
    int ny = 8;
    if (quads <= 64) {
        ny = 256 / quads;
        if (ny > 8) ny = 8;
        block = dim3(quads, ny); grid_y = 1;
    } else {
        block = dim3(32, 8); grid_y = (quads + 31) / 32;
    }
    npb = 4 * ny;
    const int64_t want = (nodes + 591) / 592;                          // 148 SMs x 4 blocks
    if (want > npb) npb = (int)((want + 2 * ny - 1) / (2 * ny)) * (2 * ny);
}

}  // namespace nt

using namespace nt;

extern "C" int nt_bn_fold(const double *stats, int64_t count, int C, const float *gamma, const float *beta,
                          float *running_mean, float *running_var, int64_t *num_batches_tracked, float momentum,
                          float eps, int training, float *mean, float *rstd, float *s, float *t,
                          const float *w_next, const float *b_next, int n_next, float *w_f, float *w_ft, float *b_f,
                          void *stream) {
    NT_REQUIRE(C >= 1 && gamma && beta && mean && rstd && s && t, "nt_bn_fold: bad arguments");
    NT_REQUIRE(training ? (stats != nullptr && count >= 1) : (running_mean && running_var),
               "nt_bn_fold: statistics missing");
    if (!w_next) n_next = 0;
    NT_REQUIRE(n_next == 0 || (w_f && b_f), "nt_bn_fold: fold outputs missing");
    bn_fold_kernel<<<n_next + 1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        stats, count, C, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, training, mean,
        rstd, s, t, w_next, b_next, n_next, w_f, w_ft, b_f);
    return check_launch("nt_bn_fold");
}

extern "C" int nt_maxmin_finish(const float *vmax, const float *vmin, const uint8_t *imax, const uint8_t *imin,
                                const float *s, const float *t, int64_t M, int C, float *out, int ldo, uint8_t *sel,
                                float *vsel, const float *tail_src, int tail_ld, int tail, void *stream) {
    NT_REQUIRE(vmax && vmin && imax && imin && s && t && out && M >= 0 && C >= 1, "nt_maxmin_finish: bad arguments");
    NT_REQUIRE(tail >= 0 && ldo >= C + tail && (tail == 0 || tail_src), "nt_maxmin_finish: bad tail/ldo");
    if (M == 0) return 0;
    const int64_t n = M * (C + tail);
    maxmin_finish_kernel<<<blocks_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        vmax, vmin, imax, imin, s, t, M, C, out, ldo, sel, vsel, tail_src, tail_ld, tail);
    return check_launch("nt_maxmin_finish");
}

extern "C" int nt_bn_apply(const float *a, int lda, const float *s, const float *t, int64_t rows, int C, float *out,
                           int ldo, void *stream) {
    NT_REQUIRE(a && s && t && out && rows >= 0 && C >= 1 && lda >= C && ldo >= C, "nt_bn_apply: bad arguments");
    if (rows == 0) return 0;
    bn_apply_kernel<<<blocks_for(rows * C, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, lda, s, t, rows,
                                                                                                   C, out, ldo);
    return check_launch("nt_bn_apply");
}

extern "C" int nt_bn_bwd_reduce(const float *g, int ldg, const float *v, int ldv, const float *mu, const float *rstd,
                                int64_t rows, int C, double *sums, void *stream) {
    NT_REQUIRE(g && sums && rows >= 0 && C >= 1 && ldg >= C, "nt_bn_bwd_reduce: bad arguments");
    NT_REQUIRE(!v || (mu && rstd && ldv >= C), "nt_bn_bwd_reduce: bad BN operands");
    if (rows == 0) return 0;
    dim3 grid(blocks_for(rows, CR_ROWS), (C + 31) / 32), block(32, 8);
    bn_bwd_reduce_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, ldg, v, ldv, mu, rstd, rows, C,
                                                                                   sums);
    return check_launch("nt_bn_bwd_reduce");
}

extern "C" int nt_edge_stats(const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                             int64_t rows, int H, double *stats, void *stream) {
    NT_REQUIRE(pq && stats && rows >= 0 && H >= 1 && ldpq >= H, "nt_edge_stats: bad arguments");
    NT_REQUIRE(!idx || (k >= 1 && n_per_cloud >= 1), "nt_edge_stats: edge operand needs k and n_per_cloud");
    if (rows == 0) return 0;
    dim3 grid(blocks_for(rows, CR_ROWS), (H + 31) / 32), block(32, 8);
    edge_stats_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pq, ldpq, qoff, idx, k, n_per_cloud,
                                                                                rows, H, stats);
    return check_launch("nt_edge_stats");
}

extern "C" int nt_edge_activation(const float *pq, int ldpq, int qoff, const int32_t *idx, int k, int n_per_cloud,
                                  int64_t rows, int H, float *out, int ldo, double *stats, void *stream) {
    NT_REQUIRE(pq && out && rows >= 0 && H >= 1 && ldpq >= H && ldo >= H, "nt_edge_activation: bad arguments");
    NT_REQUIRE(!idx || (k >= 1 && n_per_cloud >= 1), "nt_edge_activation: edge operand needs k and n_per_cloud");
    if (rows == 0) return 0;
    dim3 block(32, 8);
    if (idx && (H & 3) == 0 && (ldpq & 3) == 0 && (qoff & 3) == 0 && (ldo & 3) == 0 && aligned16(pq) && aligned16(out) && rows % k == 0) {
        unsigned gy; int npb;
        quad_geometry(H, rows / k, block, gy, npb);
        dim3 grid(blocks_for(rows / k, npb), gy);
        if (k <= 5)
            edge_activation_x2_kernel<5><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pq, ldpq, qoff, idx, k, n_per_cloud,
                                                                                                   rows / k, H, out, ldo, stats, npb);
        else
            edge_activation_v4_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pq, ldpq, qoff, idx, k, n_per_cloud,
                                                                                                rows / k, H, out, ldo, stats, npb);
    } else {
        dim3 grid(blocks_for(rows, EA_ROWS), (H + 127) / 128);
        edge_activation_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pq, ldpq, qoff, idx, k, n_per_cloud,
                                                                                         rows, H, out, ldo, stats);
    }
    return check_launch("nt_edge_activation");
}

extern "C" int nt_bn_relu_bwd_last(const float *a, int lda, const float *g, int ldg, const uint8_t *sel, int k,
                                   const float *s, const float *mu, const float *rstd, const double *sums,
                                   int64_t count, int64_t rows, int C, float *dz, int lddz, double *colsum,
                                   void *stream) {
    NT_REQUIRE(a && g && s && mu && rstd && sums && dz && k >= 1 && count >= 1 && C >= 1, "nt_bn_relu_bwd_last: bad arguments");
    NT_REQUIRE(lda >= C && ldg >= C && lddz >= C && rows % k == 0, "nt_bn_relu_bwd_last: bad strides");
    if (rows == 0) return 0;
    dim3 block(32, 8);
    const int Cp = (C + 3) & ~3;      // the float4 kernel touches whole column quads: rows must be padded to a multiple of 4
    if ((lda & 3) == 0 && (lddz & 3) == 0 && lda >= Cp && lddz >= Cp && aligned16(a) && aligned16(dz)) {
        unsigned gy; int npb;
        quad_geometry(C, rows / k, block, gy, npb);
        dim3 grid(blocks_for(rows / k, npb), gy);
        if (k <= 5)
            bn_relu_bwd_last_x2_kernel<5><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
                a, lda, g, ldg, sel, k, s, mu, rstd, sums, count, rows / k, C, dz, lddz, colsum, npb);
        else
            bn_relu_bwd_last_v4_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
                a, lda, g, ldg, sel, k, s, mu, rstd, sums, count, rows / k, C, dz, lddz, colsum, npb);
    } else {
        dim3 grid(blocks_for(rows, BL_ROWS), (C + 127) / 128);
        bn_relu_bwd_last_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            a, lda, g, ldg, sel, k, s, mu, rstd, sums, count, rows, C, dz, lddz, colsum);
    }
    return check_launch("nt_bn_relu_bwd_last");
}

extern "C" int nt_linear_bn_bwd(const double *rawc, const double *csum, int n_out, int C, const float *w,
                                const float *s, const float *beta, const float *rstd, int64_t count,
                                float *dW, float *db, float *dgamma, float *dbeta, float *k0, float *k1, void *stream) {
    NT_REQUIRE(rawc && csum && dW && db && n_out >= 1 && C >= 1, "nt_linear_bn_bwd: bad arguments");
    NT_REQUIRE(w && s && beta && rstd && dgamma && dbeta && k0 && k1 && count >= 1,
               "nt_linear_bn_bwd: BN operands missing");
    linear_bn_bwd_kernel<<<(C + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        rawc, csum, n_out, C, w, s, beta, rstd, count, dW, db, dgamma, dbeta, k0, k1);
    return check_launch("nt_linear_bn_bwd");
}

extern "C" int nt_edge_scatter(const float *dz, int lddz, const int32_t *idx, int k, int n_per_cloud, int64_t M, int H,
                               float *dpq, int lddpq, void *stream) {
    NT_REQUIRE(dz && idx && dpq && k >= 1 && n_per_cloud >= 1 && M >= 0 && H >= 1 && lddz >= H && lddpq >= 2 * H,
               "nt_edge_scatter: bad arguments");
    if (M == 0) return 0;
    if ((H & 3) == 0 && (lddz & 3) == 0 && (lddpq & 3) == 0 && aligned16(dz) && aligned16(dpq)) {
        dim3 block;
        unsigned gy; int npb;
        quad_geometry(H, 0, block, gy, npb);          // no per-block atomics here: small blocks (larger ones measured 0.111 vs 0.085 ms)
        dim3 grid(blocks_for(M, npb), gy);
        edge_scatter_v4_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dz, lddz, idx, k, n_per_cloud, M, H,
                                                                                         dpq, lddpq, npb);
    } else {
        edge_scatter_kernel<<<blocks_for(M * H, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            dz, lddz, idx, k, n_per_cloud, M, H, dpq, lddpq);
    }
    return check_launch("nt_edge_scatter");
}
