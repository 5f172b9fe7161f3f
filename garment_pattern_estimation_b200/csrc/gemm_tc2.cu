// Persistent, warp-specialised version of the fused row GEMM (same contract as gemm_tc.cu / nt_gemm_nt), sm_100a.
//
// One CTA per SM loops over row tiles; three pipelines run concurrently inside it:
//   warps 0-7   A-operand producers (256 threads: thread = row x half of the K chunks) with a depth-3 register prefetch ring
//               that runs ACROSS tile boundaries, 4 shared-memory stages (W tile by cp.async.bulk, A tile by st.shared);
//   warp 12     single-thread tcgen05.mma issue into one of TWO TMEM accumulators (2 x 256 columns);
//   warps 8-11  epilogue (TMEM lane quadrant = warp % 4): drains accumulator t while the tensor core already works on
//               tile t+1 and the producers on tile t+1 / t+2.
// The one-tile-per-CTA kernel exposes the gather latency and the whole epilogue once per tile with only 12 resident warps
// (ncu: tensor pipe 13-22 % active, stalls = long scoreboard + barrier); here 13 warps of ONE CTA cover all three phases.
#include "gemm_tc_shared.cuh"

namespace nt {

constexpr int P2_PRODUCER_WARPS = 8;
constexpr int P2_EPI_WARP0 = 8;                 // warps 8..11 (warp % 4 = TMEM lane quadrant)
constexpr int P2_MMA_WARP = 12;
constexpr int P2_THREADS = 13 * 32;
constexpr int P2_STAGES = 4;                    // ring capacity; `stages` (3 or 4) are used depending on the tile width

template <int PROD, int EPI>
__global__ void __launch_bounds__(P2_THREADS, 1) gemm_nt_tc2_kernel(NTParams p, const uint8_t *__restrict__ w_split, TCGeom g, int stages) {
    constexpr int EPC = 4;                       // TF32x3 only
    extern __shared__ __align__(128) uint8_t smem[];
    const size_t stage_bytes = tc_stage_bytes(g.n_tile);
    uint8_t *epi_base = smem + stages * stage_bytes;
    float *vt = reinterpret_cast<float *>(epi_base);                              // [128][33]   (max/min aggregation)
    float *tw_all = vt + 128 * 33;                                                 // 4 x [32][33] (per-warp transposition)
    const float **rowp = reinterpret_cast<const float **>(tw_all + 4 * 32 * 33);   // [128] aux row pointers
    const float **rowq = rowp + 128;
    float *red = reinterpret_cast<float *>(rowq + 128);                            // [2][256] column statistics
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 512);                      // [4]
    uint64_t *empty = full + P2_STAGES;                                            // [4]
    uint64_t *tmem_full = empty + P2_STAGES;                                       // [2]
    uint64_t *tmem_empty = tmem_full + 2;                                          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_n = blockIdx.y;
    const int col0 = tile_n * g.n_tile;
    const int64_t n_row_tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
    const int my_tiles = (int)((n_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles blockIdx.x, +grid, ...

    for (int i = tid; i < 512; i += P2_THREADS) red[i] = 0.f;
    if (warp == P2_MMA_WARP && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], P2_PRODUCER_WARPS * 32 + 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        mbar_fence_init();
    }
    if (warp == P2_MMA_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_it = my_tiles * g.num_kb;

    if (warp < P2_PRODUCER_WARPS) {
        // =========================== producers ===========================
        const int r = tid & 127, half = tid >> 7;
        const uint8_t *wsrc = w_split + (size_t)tile_n * g.num_kb * ((size_t)g.n_tile * 128);
        const bool vec = (PROD == NT_PROD_PLAIN) ? (((p.lda & 3) == 0) && aligned16(p.a))
                                                 : (((p.e.ldpq & 3) == 0) && ((p.e.qoff & 3) == 0) && aligned16(p.e.pq));
        struct Regs { float p[2][EPC]; float q[PROD == NT_PROD_EDGE ? 2 : 1][EPC]; };
        Regs v0, v1, v2;
        int cur_tile = -1;                       // tile index (0..my_tiles) whose row pointers are cached
        const float *ap = nullptr, *aq = nullptr;
        bool r_ok = false;

        auto fetch = [&](int it, Regs &v) {
            const bool live_it = it < total_it;
            const int ti = live_it ? it / g.num_kb : 0, kb = live_it ? it - ti * g.num_kb : 0;
            if (live_it && ti != cur_tile) {
                cur_tile = ti;
                const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * p.rows_per_tile;
                const int rows_here = (int)min((int64_t)p.rows_per_tile, p.rows - row0);
                r_ok = r < rows_here;
                ap = aq = nullptr;
                if (r_ok) {
                    if (PROD == NT_PROD_PLAIN) ap = p.a + (row0 + r) * (int64_t)p.lda;
                    else edge_row_ptrs(p.e, row0 + r, ap, aq);
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k = (kb * 4 + 2 * half + jj) * EPC;
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = 0.f;
                const bool live = live_it && r_ok && k < p.K;
                if (live) load_chunk<EPC>(ap, k, p.K, vec, t);
#pragma unroll
                for (int e = 0; e < EPC; ++e) v.p[jj][e] = t[e];
                if (PROD == NT_PROD_EDGE) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) t[e] = 0.f;
                    if (live && aq) load_chunk<EPC>(aq, k, p.K, vec, t);
#pragma unroll
                    for (int e = 0; e < EPC; ++e) v.q[jj][e] = t[e];
                }
            }
        };
        auto consume = [&](int it, Regs &v) {
            const int s = it % stages, use = it / stages;
            const int kb = it % g.num_kb;
            mbar_wait(&empty[s], (use & 1) ^ 1);
            uint8_t *a_hi = smem + s * stage_bytes, *a_lo = a_hi + TC_A_BYTES, *b_all = a_lo + TC_A_BYTES;
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)g.n_tile * 128u;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(b_all, wsrc + (size_t)kb * bytes, bytes, &full[s]);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * half + jj;
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = 0.f;
#pragma unroll
                for (int e = 0; e < EPC; ++e)
                    t[e] = (PROD == NT_PROD_EDGE) ? fmaxf(v.p[jj][e] + v.q[jj][e], 0.f) : v.p[jj][e];
                uint4 h, l;
                pack_chunk(t, true, h, l);
                *reinterpret_cast<uint4 *>(a_hi + j * (TC_M * 16) + r * 16) = h;
                *reinterpret_cast<uint4 *>(a_lo + j * (TC_M * 16) + r * 16) = l;
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
        };
        fetch(0, v0); fetch(1, v1); fetch(2, v2);
        for (int it = 0; it < total_it; it += 3) {
            consume(it, v0); fetch(it + 3, v0);
            if (it + 1 < total_it) { consume(it + 1, v1); fetch(it + 4, v1); }
            if (it + 2 < total_it) { consume(it + 2, v2); fetch(it + 5, v2); }
        }
    } else if (warp == P2_MMA_WARP) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, (uint32_t)g.n_tile, 0, 0);
            const uint32_t lbo_a = TC_M * 16, lbo_b = (uint32_t)g.n_tile * 16, sbo = 128;
            int it = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int as = ti & 1, ause = ti >> 1;
                mbar_wait(&tmem_empty[as], (ause & 1) ^ 1);         // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(as * 256);
                for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
                    const int s = it % stages, use = it / stages;
                    mbar_wait(&full[s], use & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * stage_bytes), a_lo = a_hi + TC_A_BYTES;
                    const uint32_t b_hi = a_lo + TC_A_BYTES, b_lo = b_hi + 4 * lbo_b;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t dah = make_smem_desc(a_hi + kk * 2 * lbo_a, lbo_a, sbo);
                        const uint64_t dal = make_smem_desc(a_lo + kk * 2 * lbo_a, lbo_a, sbo);
                        const uint64_t dbh = make_smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, sbo);
                        const uint64_t dbl = make_smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, sbo);
                        umma_tf32(d, dah, dbh, idesc, (kb | kk) ? 1u : 0u);
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dal, dbh, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[as]);
            }
        }
    } else {
        // =========================== epilogue warps (thread = row of the tile, TMEM lane = row) ===========================
        const int quad = warp & 3;
        const int et = tid - P2_EPI_WARP0 * 32;                   // 0..127
        const int r = quad * 32 + lane;
        float *tw = tw_all + quad * (32 * 33);
        const int n_chunks = (g.n_tile + 31) / 32;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int as = ti & 1, ause = ti >> 1;
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * p.rows_per_tile;
            const int rows_here = (int)min((int64_t)p.rows_per_tile, p.rows - row0);
            const bool valid = r < rows_here;
            const int64_t grow = row0 + r;
            const int64_t wrow0 = row0 + quad * 32;
            const int wrows = max(0, min(32, rows_here - quad * 32));
            if (EPI == NT_EPI_BNRELU_BWD) {
                const float *aux_p = nullptr, *aux_q = nullptr;
                if (valid) {
                    if (p.aux_edge) edge_row_ptrs(p.ae, grow, aux_p, aux_q);
                    else aux_p = p.aux + grow * (int64_t)p.ldaux;
                }
                rowp[r] = aux_p; rowq[r] = aux_q;
                __syncwarp();
            }
            mbar_wait(&tmem_full[as], ause & 1);
            tc_fence_after();
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int c0 = ch * 32;
                const int cl = col0 + c0 + lane;
                const bool cl_ok = (c0 + lane) < g.n_tile && cl < p.n_out;
                float auxv[32];
                if (EPI == NT_EPI_BNRELU_BWD) {
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        float a = 0.f;
                        if (rr < wrows && cl_ok) {
                            const float *pp = rowp[quad * 32 + rr];
                            a = pp[cl];
                            if (p.aux_edge) {
                                const float *qq = rowq[quad * 32 + rr];
                                if (qq) a += __ldg(qq + cl);
                                a = fmaxf(a, 0.f);
                            }
                        }
                        auxv[rr] = a;
                    }
                }
                float acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + c0), acc);
                if (EPI == NT_EPI_BNRELU_BWD) {
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) tw[rr * 33 + lane] = auxv[rr];
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 32; ++i) auxv[i] = tw[lane * 33 + i];
                    __syncwarp();
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int c = col0 + c0 + i;
                    const bool c_ok = (c0 + i) < g.n_tile && c < p.n_out;
                    float o = 0.f;
                    if (EPI == NT_EPI_BIAS) {
                        o = acc[i] + ((c_ok && p.bias) ? __ldg(p.bias + c) : 0.f);
                    } else if (EPI == NT_EPI_RELU_STATS || EPI == NT_EPI_RELU_MAXMIN) {
                        o = fmaxf(acc[i] + ((c_ok && p.bias) ? __ldg(p.bias + c) : 0.f), 0.f);
                    } else if (valid && c_ok) {
                        const float a = auxv[i];
                        o = (a > 0.f) ? (acc[i] - __ldg(p.k0 + c) - (a - __ldg(p.mu + c)) * __ldg(p.k1 + c)) : 0.f;
                    }
                    acc[i] = o;
                }
                if (p.out) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) tw[lane * 33 + i] = acc[i];
                    __syncwarp();
                    if (cl_ok) {
                        float *dst = p.out + wrow0 * (int64_t)p.ldo + cl;
                        for (int rr = 0; rr < wrows; ++rr) dst[(int64_t)rr * p.ldo] = tw[rr * 33 + lane];
                    }
                    __syncwarp();
                }
                if (EPI != NT_EPI_BIAS) {
                    const bool want = (EPI == NT_EPI_BNRELU_BWD) ? (p.colsum != nullptr) : (p.stats != nullptr);
                    if (want) {
                        float sv[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const bool c_ok = (c0 + i) < g.n_tile && (col0 + c0 + i) < p.n_out;
                            sv[i] = (valid && c_ok) ? acc[i] : 0.f;
                        }
                        const float t1 = warp_column_sums(sv, lane);
                        atomicAdd(&red[c0 + lane], t1);
                        if (EPI != NT_EPI_BNRELU_BWD) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const bool c_ok = (c0 + i) < g.n_tile && (col0 + c0 + i) < p.n_out;
                                sv[i] = (valid && c_ok) ? acc[i] * acc[i] : 0.f;
                            }
                            const float t2 = warp_column_sums(sv, lane);
                            atomicAdd(&red[256 + c0 + lane], t2);
                        }
                    }
                }
                if (EPI == NT_EPI_RELU_MAXMIN) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) vt[r * 33 + i] = acc[i];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int kk = p.k_agg;
                    const int nodes_here = rows_here / kk;
                    const int64_t node0 = row0 / kk;
                    for (int t = et; t < nodes_here * 32; t += 128) {
                        const int nd = t >> 5, cc = t & 31;
                        const int c = col0 + c0 + cc;
                        if ((c0 + cc) >= g.n_tile || c >= p.n_out) continue;
                        float mx = vt[(nd * kk) * 33 + cc], mn = mx;
                        int ix = 0, in = 0;
                        for (int sl = 1; sl < kk; ++sl) {
                            const float x = vt[(nd * kk + sl) * 33 + cc];
                            if (x > mx) { mx = x; ix = sl; }
                            if (x < mn) { mn = x; in = sl; }
                        }
                        const int64_t o = (node0 + nd) * (int64_t)p.n_out + c;
                        p.vmax[o] = mx; p.vmin[o] = mn; p.imax[o] = (uint8_t)ix; p.imin[o] = (uint8_t)in;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
            // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[as]);
        }
        // flush the per-CTA column statistics once (they accumulate over all of this CTA's tiles)
        if (EPI != NT_EPI_BIAS) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int c = et; c < g.n_tile; c += 128) {
                const int col = col0 + c;
                if (col >= p.n_out) continue;
                if (EPI == NT_EPI_BNRELU_BWD) {
                    if (p.colsum) atomicAdd(p.colsum + col, (double)red[c]);
                } else if (p.stats) {
                    atomicAdd(p.stats + col, (double)red[c]);
                    atomicAdd(p.stats + p.n_out + col, (double)red[256 + c]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == P2_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

static size_t tc2_smem_bytes(int n_tile, int stages) {
    return stages * tc_stage_bytes(n_tile) + (size_t)(128 * 33 + 4 * 32 * 33) * 4 + 256 * 8 + 512 * 4 + 16 * 8 + 16;
}

template <int PROD, int EPI>
static int launch_tc2(const NTParams &p, const void *w_split, cudaStream_t st) {
    const TCGeom g = tc_geometry(p.n_out, p.K, NT_PREC_TF32X3);
    const int stages = tc2_smem_bytes(g.n_tile, 4) <= 227 * 1024 ? 4 : 3;
    const size_t smem = tc2_smem_bytes(g.n_tile, stages);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nt_tc2_kernel<PROD, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return fail("nt_gemm_nt(tc2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int64_t n_row_tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
    int ctas = (int)(n_row_tiles < 148 / g.n_tiles ? n_row_tiles : 148 / g.n_tiles);
    if (ctas < 1) ctas = 1;
    dim3 grid(ctas, g.n_tiles);
    gemm_nt_tc2_kernel<PROD, EPI><<<grid, P2_THREADS, smem, st>>>(p, reinterpret_cast<const uint8_t *>(w_split), g, stages);
    return check_launch("nt_gemm_nt(tc2)");
}

int launch_nt_tc2(const NTParams &p, int producer, int epilogue, const void *w_split, cudaStream_t st) {
    const bool edge = producer == NT_PROD_EDGE;
    switch (epilogue) {
        case NT_EPI_BIAS: return launch_tc2<NT_PROD_PLAIN, NT_EPI_BIAS>(p, w_split, st);
        case NT_EPI_RELU_STATS:
            return edge ? launch_tc2<NT_PROD_EDGE, NT_EPI_RELU_STATS>(p, w_split, st)
                        : launch_tc2<NT_PROD_PLAIN, NT_EPI_RELU_STATS>(p, w_split, st);
        case NT_EPI_RELU_MAXMIN:
            return edge ? launch_tc2<NT_PROD_EDGE, NT_EPI_RELU_MAXMIN>(p, w_split, st)
                        : launch_tc2<NT_PROD_PLAIN, NT_EPI_RELU_MAXMIN>(p, w_split, st);
        default: return launch_tc2<NT_PROD_PLAIN, NT_EPI_BNRELU_BWD>(p, w_split, st);
    }
}

}  // namespace nt
