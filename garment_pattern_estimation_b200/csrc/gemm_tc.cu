// Tensor-core engine for the fused row GEMMs (sm_100a): tcgen05.mma with the accumulator in TMEM.
//
//   out = epilogue( producer(A) . W^T )       A: [rows, K] produced on the fly, W: [n_out, K]  (same contract as gemm.cu)
//
// Precision: the reference computes these Linear layers in fp32 (cuBLAS SGEMM, TF32 off).  To stay inside the 1e-3
// activation tolerance AND keep the second EdgeConv layer's kNN graph stable, every fp32 operand is split into two bf16
// terms (x = hi + lo, |x - hi - lo| <= 2^-17 |x|) and each K-step issues three MMAs  hi.hi + hi.lo + lo.hi  into the
// same fp32 TMEM accumulator (error-compensated "bf16x3", ~1e-5 relative).
//
// CTA = one tile of up to 128 rows x N_tile (<= 256) columns; 2 CTAs per SM (85 KB smem, 256 TMEM columns each) so one
// CTA's epilogue overlaps the other's main loop.
//   warps 0-3 : A-operand producers during the main loop (thread = row: gather / ReLU / split / st.shared in the UMMA
//               no-swizzle K-major core-matrix layout), then the epilogue (tcgen05.ld, thread = row).
//   warp 4    : barrier init + single-thread MMA issue (tcgen05.mma / tcgen05.commit).
//   warp 5    : TMEM alloc / dealloc.
//   W tiles   : pre-split bf16 in the same core-matrix layout (nt_gemm_prepare_weights), one cp.async.bulk per stage.
#include "gemm_tc_shared.cuh"

namespace nt {

// ------------------------------------------------------------------------------------------------------------------
// weight pre-split: W [n_out, K] fp32 -> per (column tile, K block): [hi|lo][chunk 0..3][n 0..n_tile)[8 bf16]
// ------------------------------------------------------------------------------------------------------------------
__global__ void tc_prepare_weights_kernel(const float *__restrict__ w, int ldw, int n_out, int K, TCGeom g,
                                          uint4 *__restrict__ out) {
    // one thread per 16-byte chunk of the hi plane (and the matching lo chunk)
    const int64_t per_block = (int64_t)4 * g.n_tile;                     // chunks per plane per (tile, kb)
    const int64_t total = (int64_t)g.n_tiles * g.num_kb * per_block;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t blk = i / per_block;
    const int within = (int)(i - blk * per_block);
    const int j = within / g.n_tile, n = within - j * g.n_tile;
    const int tile = (int)(blk / g.num_kb), kb = (int)(blk - (int64_t)tile * g.num_kb);
    const int col = tile * g.n_tile + n;
    const int k0 = (kb * 4 + j) * g.epc;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (e < g.epc && col < n_out && k0 + e < K) ? w[(int64_t)col * ldw + k0 + e] : 0.f;
    uint4 *base = out + blk * (2 * per_block);
    uint4 h, l;
    pack_chunk(v, g.epc == 4, h, l);
    base[within] = h;
    base[per_block + within] = l;
}

template <int PROD, int EPI, bool TF32>
__global__ void __launch_bounds__(TC_THREADS, 2) gemm_nt_tc_kernel(NTParams p, const uint8_t *__restrict__ w_split, TCGeom g) {
    extern __shared__ __align__(128) uint8_t smem[];
    const size_t stage_bytes = tc_stage_bytes(g.n_tile);
    uint8_t *stage_base[TC_STAGES] = {smem, smem + stage_bytes};
    uint8_t *tail = smem + TC_STAGES * stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);              // [2]
    uint64_t *empty = full + TC_STAGES;                               // [2]
    uint64_t *tmem_full = empty + TC_STAGES;                          // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    float *red = reinterpret_cast<float *>(tail + 64);                // [2][256]
    float *colv = red + 512;                                          // [4][256]: bias | k0 | k1 | mu of this column tile

    constexpr int EPC = TF32 ? 4 : 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * p.rows_per_tile;
    const int rows_here = (int)min((int64_t)p.rows_per_tile, p.rows - row0);
    const int tile_n = blockIdx.y;
    const int col0 = tile_n * g.n_tile;

    if (tid < 128) {
        for (int i = tid; i < 512; i += 128) red[i] = 0.f;
        for (int i = tid; i < 256; i += 128) {
            const int c = tile_n * g.n_tile + i;
            const bool ok = i < g.n_tile && c < p.n_out;
            colv[i] = (ok && p.bias) ? __ldg(p.bias + c) : 0.f;
            colv[256 + i] = (ok && EPI == NT_EPI_BNRELU_BWD) ? __ldg(p.k0 + c) : 0.f;
            colv[512 + i] = (ok && EPI == NT_EPI_BNRELU_BWD) ? __ldg(p.k1 + c) : 0.f;
            colv[768 + i] = (ok && EPI == NT_EPI_BNRELU_BWD) ? __ldg(p.mu + c) : 0.f;
        }
    }
    if (warp == 4 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 128 + 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== A-operand producers ===========================
        // Lane mapping: chunk jj = lane >> 3 (16 bytes of K) of the rows warp*32 + 8*i + (lane & 7), i = 0..3.  One warp load
        // then touches 8 rows x 64 contiguous bytes (8 L1 lines) instead of 32 rows x 16 bytes (32 lines) -- ncu showed the
        // L1TEX tag stage as the busiest unit of the thread-per-row mapping -- and the 16-byte shared-memory stores of a
        // quarter-warp land in 128 contiguous bytes (conflict-free).
        const int r = tid;
        const bool r_ok = r < rows_here;
        const int jj = lane >> 3;
        int prow[4];
        bool p_ok[4];
        const float *ap[4], *aq[4];
        bool vec = false;
        if (PROD == NT_PROD_PLAIN) vec = ((p.lda & 3) == 0) && aligned16(p.a);
        else vec = ((p.e.ldpq & 3) == 0) && ((p.e.qoff & 3) == 0) && aligned16(p.e.pq);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            prow[i] = warp * 32 + 8 * i + (lane & 7);
            p_ok[i] = prow[i] < rows_here;
            ap[i] = aq[i] = nullptr;
            if (p_ok[i]) {
                if (PROD == NT_PROD_PLAIN) ap[i] = p.a + (row0 + prow[i]) * (int64_t)p.lda;
                else edge_row_ptrs(p.e, row0 + prow[i], ap[i], aq[i]);
            }
        }
        const uint8_t *wsrc = w_split + (size_t)tile_n * g.num_kb * ((size_t)g.n_tile * 128);
        // Register prefetch ring, depth 3: the gather is latency-bound (random rows of PQ miss L2 while the activation
        // stream evicts it), so the loads of K blocks kb+1 .. kb+3 are in flight while block kb is packed and handed to
        // the tensor core.  (Loop unrolled by 3 so every buffer is addressed statically.)
        // NOTE: fetch() only LOADS (raw centre and neighbour values stay in separate registers); the add / ReLU that
        // consumes them is deferred to consume(), otherwise the first use would stall on the loads right away.
        struct Regs { float p[4][EPC]; float q[PROD == NT_PROD_EDGE ? 4 : 1][EPC]; };
        Regs v0, v1, v2;
        auto fetch = [&](int kb, Regs &v) {
            const int k = (kb * 4 + jj) * EPC;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = 0.f;
                const bool live = p_ok[i] && kb < g.num_kb && k < p.K;
                if (live) load_chunk<EPC>(ap[i], k, p.K, vec, t);
#pragma unroll
                for (int e = 0; e < EPC; ++e) v.p[i][e] = t[e];
                if (PROD == NT_PROD_EDGE) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) t[e] = 0.f;
                    if (live && aq[i]) load_chunk<EPC>(aq[i], k, p.K, vec, t);
#pragma unroll
                    for (int e = 0; e < EPC; ++e) v.q[i][e] = t[e];
                }
            }
        };
        auto consume = [&](int kb, Regs &v) {
            const int s = kb & 1, use = kb >> 1;
            mbar_wait(&empty[s], (use & 1) ^ 1);
            uint8_t *a_hi = stage_base[s], *a_lo = a_hi + TC_A_BYTES, *b_all = a_lo + TC_A_BYTES;
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)g.n_tile * 128u;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(b_all, wsrc + (size_t)kb * bytes, bytes, &full[s]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = 0.f;
#pragma unroll
                for (int e = 0; e < EPC; ++e)
                    t[e] = (PROD == NT_PROD_EDGE) ? fmaxf(v.p[i][e] + v.q[i][e], 0.f) : v.p[i][e];
                uint4 h, l;
                pack_chunk(t, TF32, h, l);
                *reinterpret_cast<uint4 *>(a_hi + jj * (TC_M * 16) + prow[i] * 16) = h;
                *reinterpret_cast<uint4 *>(a_lo + jj * (TC_M * 16) + prow[i] * 16) = l;
            }
            fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&full[s]);
        };
        fetch(0, v0); fetch(1, v1); fetch(2, v2);
        for (int kb = 0; kb < g.num_kb; kb += 3) {
            consume(kb, v0); fetch(kb + 3, v0);
            if (kb + 1 < g.num_kb) { consume(kb + 1, v1); fetch(kb + 4, v1); }
            if (kb + 2 < g.num_kb) { consume(kb + 2, v2); fetch(kb + 5, v2); }
        }

        // =========================== epilogue (thread = row, TMEM lane = row) ===========================
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float *vt = reinterpret_cast<float *>(smem);                    // [128][33] staging (aliases stage 0)
        // per-warp transposition tile: 32 x 36 floats (16-byte accesses, fast path) or 32 x 33 (scalar path) in the same region;
        // 4 x 4608 B = 18432 B = the smallest possible stage (n_tile = 16)
        float *tw4 = reinterpret_cast<float *>(stage_base[1]) + warp * (32 * 36);
        float *tw = tw4;
        const bool valid = r_ok;
        const int64_t grow = row0 + r;
        const int64_t wrow0 = row0 + warp * 32;                          // first row of this warp
        const int wrows = max(0, min(32, rows_here - warp * 32));        // valid rows of this warp
        // aux row pointers of all 128 rows, shared through smem (stage 0 is free; the aggregation tile is not used here)
        const float **rowp = reinterpret_cast<const float **>(smem);
        const float **rowq = rowp + 128;
        if (EPI == NT_EPI_BNRELU_BWD) {
            const float *aux_p = nullptr, *aux_q = nullptr;
            if (valid) {
                if (p.aux_edge) edge_row_ptrs(p.ae, grow, aux_p, aux_q);
                else aux_p = p.aux + grow * (int64_t)p.ldaux;
            }
            rowp[r] = aux_p; rowq[r] = aux_q;
            __syncwarp();                        // every warp only reads the 32 entries it wrote itself
        }
        const int n_chunks = (g.n_tile + 31) / 32;
        // vectorised row-wise access (thread = row, 32 consecutive columns = 128 contiguous bytes) needs 16-byte aligned rows
        const bool vec_io = (!p.out || (((p.ldo & 3) == 0) && aligned16(p.out))) &&
                            (EPI != NT_EPI_BNRELU_BWD ||
                             (p.aux_edge ? (((p.ae.ldpq & 3) == 0) && ((p.ae.qoff & 3) == 0) && aligned16(p.ae.pq))
                                         : (((p.ldaux & 3) == 0) && aligned16(p.aux))));
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int c0 = ch * 32;
            if (vec_io && c0 + 32 <= g.n_tile && col0 + c0 + 32 <= p.n_out) {
                // ================= fast path: all 32 columns of the chunk exist =================
                // Global traffic is COALESCED (a warp request = 4 rows x 128 contiguous bytes: lane -> row 4m + (lane >> 3),
                // 16-byte column group lane & 7) and exchanged with the thread-per-row register layout of the TMEM load through a
                // per-warp 32 x 36 transposition tile with conflict-free 16-byte accesses.
                const int cg = col0 + c0;                                // first global column of the chunk
                const int sub = lane >> 3, q4 = (lane & 7) * 4;
                float auxv[32];
                if (EPI == NT_EPI_BNRELU_BWD) {
                    float4 ld[8];
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int rr = 4 * m + sub;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rr < wrows) {
                            a = __ldg(reinterpret_cast<const float4 *>(rowp[warp * 32 + rr] + cg + q4));
                            if (p.aux_edge) {
                                const float *qq = rowq[warp * 32 + rr];
                                if (qq) {
                                    const float4 b = __ldg(reinterpret_cast<const float4 *>(qq + cg + q4));
                                    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                                }
                                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                            }
                        }
                        ld[m] = a;
                    }
#pragma unroll
                    for (int m = 0; m < 8; ++m) *reinterpret_cast<float4 *>(tw4 + (4 * m + sub) * 36 + q4) = ld[m];
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 a = *reinterpret_cast<const float4 *>(tw4 + lane * 36 + 4 * i);
                        auxv[4 * i] = a.x; auxv[4 * i + 1] = a.y; auxv[4 * i + 2] = a.z; auxv[4 * i + 3] = a.w;
                    }
                    __syncwarp();
                }
                float acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(colv + c0 + 4 * i);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                    if (EPI == NT_EPI_BNRELU_BWD) {
                        const float4 k04 = *reinterpret_cast<const float4 *>(colv + 256 + c0 + 4 * i);
                        const float4 k14 = *reinterpret_cast<const float4 *>(colv + 512 + c0 + 4 * i);
                        const float4 mu4 = *reinterpret_cast<const float4 *>(colv + 768 + c0 + 4 * i);
                        const float k0v[4] = {k04.x, k04.y, k04.z, k04.w}, k1v[4] = {k14.x, k14.y, k14.z, k14.w};
                        const float muv[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float a = auxv[4 * i + e];
                            acc[4 * i + e] = (valid && a > 0.f) ? (acc[4 * i + e] - k0v[e] - (a - muv[e]) * k1v[e]) : 0.f;
                        }
                    } else if (EPI == NT_EPI_BIAS) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[4 * i + e] += bb[e];
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[4 * i + e] = valid ? fmaxf(acc[4 * i + e] + bb[e], 0.f) : 0.f;
                    }
                }
                const bool want = (EPI == NT_EPI_BIAS || EPI == NT_EPI_RELU_MAXMIN)
                                      ? false
                                      : ((EPI == NT_EPI_BNRELU_BWD) ? (p.colsum != nullptr) : (p.stats != nullptr));
                if (p.out || want) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4 *>(tw4 + lane * 36 + 4 * i) =
                            make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                    __syncwarp();
                    if (p.out) {
                        float *dst = p.out + wrow0 * (int64_t)p.ldo + cg + q4;
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const int rr = 4 * m + sub;
                            if (rr < wrows)
                                *reinterpret_cast<float4 *>(dst + (int64_t)rr * p.ldo) =
                                    *reinterpret_cast<const float4 *>(tw4 + rr * 36 + q4);
                        }
                    }
                    if (want) {
                        // column sums of this warp's 32 rows (lane = column; invalid rows hold zeros)
                        float t1 = 0.f, t2 = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) {
                            const float x = tw4[rr * 36 + lane];
                            t1 += x;
                            if (EPI != NT_EPI_BNRELU_BWD) t2 = fmaf(x, x, t2);
                        }
                        atomicAdd(&red[c0 + lane], t1);
                        if (EPI != NT_EPI_BNRELU_BWD) atomicAdd(&red[256 + c0 + lane], t2);
                    }
                    __syncwarp();
                }
                if (EPI == NT_EPI_RELU_MAXMIN) {
                    // max / min over the k edge rows of every centre point; the same pass yields the column statistics
                    // (a thread always lands on the same column: 128 % 32 == 0)
#pragma unroll
                    for (int i = 0; i < 32; ++i) vt[r * 33 + i] = acc[i];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int kk = p.k_agg;
                    const int nodes_here = rows_here / kk;
                    const int64_t node0 = row0 / kk;
                    float t1 = 0.f, t2 = 0.f;
                    for (int t = tid; t < nodes_here * 32; t += 128) {
                        const int nd = t >> 5, cc = t & 31;
                        float x = vt[(nd * kk) * 33 + cc];
                        float mx = x, mn = x;
                        int ix = 0, in = 0;
                        t1 += x; t2 = fmaf(x, x, t2);
                        for (int sl = 1; sl < kk; ++sl) {
                            x = vt[(nd * kk + sl) * 33 + cc];
                            if (x > mx) { mx = x; ix = sl; }
                            if (x < mn) { mn = x; in = sl; }
                            t1 += x; t2 = fmaf(x, x, t2);
                        }
                        const int64_t o = (node0 + nd) * (int64_t)p.n_out + cg + cc;
                        p.vmax[o] = mx; p.vmin[o] = mn; p.imax[o] = (uint8_t)ix; p.imin[o] = (uint8_t)in;
                    }
                    if (p.stats) { atomicAdd(&red[c0 + lane], t1); atomicAdd(&red[256 + c0 + lane], t2); }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                continue;
            }
            const int cl = col0 + c0 + lane;                             // the column this lane owns in row-wise passes
            const bool cl_ok = (c0 + lane) < g.n_tile && cl < p.n_out;
            float auxv[32];
            if (EPI == NT_EPI_BNRELU_BWD) {
                // coalesced read of the aux rows (lane = column): all 32 (64 with the gathered operand) loads are issued
                // back to back, BEFORE the TMEM load, then transposed through smem so each thread gets its own row
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    float a = 0.f;
                    if (rr < wrows && cl_ok) {
                        const float *pp = rowp[warp * 32 + rr];
                        a = pp[cl];
                        if (p.aux_edge) {
                            const float *qq = rowq[warp * 32 + rr];
                            if (qq) a += __ldg(qq + cl);
                            a = fmaxf(a, 0.f);
                        }
                    }
                    auxv[rr] = a;
                }
            }
            float acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
            if (EPI == NT_EPI_BNRELU_BWD) {
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) tw[rr * 33 + lane] = auxv[rr];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; ++i) auxv[i] = tw[lane * 33 + i];
                __syncwarp();
            }
            float s1[32], s2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int c = col0 + c0 + i;
                const bool c_ok = (c0 + i) < g.n_tile && c < p.n_out;
                float o = 0.f, q1 = 0.f, q2 = 0.f;
                if (EPI == NT_EPI_BIAS) {
                    o = acc[i] + ((c_ok && p.bias) ? __ldg(p.bias + c) : 0.f);
                } else if (EPI == NT_EPI_RELU_STATS || EPI == NT_EPI_RELU_MAXMIN) {
                    o = fmaxf(acc[i] + ((c_ok && p.bias) ? __ldg(p.bias + c) : 0.f), 0.f);
                    if (valid && c_ok) { q1 = o; q2 = o * o; }
                } else {   // NT_EPI_BNRELU_BWD
                    if (valid && c_ok) {
                        const float a = auxv[i];
                        o = (a > 0.f) ? (acc[i] - __ldg(p.k0 + c) - (a - __ldg(p.mu + c)) * __ldg(p.k1 + c)) : 0.f;
                        q1 = o;
                    }
                }
                acc[i] = o; s1[i] = q1; s2[i] = q2;
            }
            // ---- stores: transpose through smem so every warp store writes 128 contiguous bytes of one row
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
            // ---- per-column statistics of this warp's 32 rows -> smem accumulators
            if (EPI != NT_EPI_BIAS) {
                const bool want = (EPI == NT_EPI_BNRELU_BWD) ? (p.colsum != nullptr) : (p.stats != nullptr);
                if (want) {
                    float t1 = warp_column_sums(s1, lane);
                    atomicAdd(&red[c0 + lane], t1);
                    if (EPI != NT_EPI_BNRELU_BWD) {
                        float t2 = warp_column_sums(s2, lane);
                        atomicAdd(&red[256 + c0 + lane], t2);
                    }
                }
            }
            // ---- max / min over the k edge rows of every centre point
            if (EPI == NT_EPI_RELU_MAXMIN) {
#pragma unroll
                for (int i = 0; i < 32; ++i) vt[r * 33 + i] = acc[i];
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int kk = p.k_agg;
                const int nodes_here = rows_here / kk;
                const int64_t node0 = row0 / kk;
                for (int t = tid; t < nodes_here * 32; t += 128) {
                    const int nd = t >> 5, cc = t & 31;
                    const int c = col0 + c0 + cc;
                    if ((c0 + cc) >= g.n_tile || c >= p.n_out) continue;
                    float mx = vt[(nd * kk) * 33 + cc], mn = mx;
                    int ix = 0, in = 0;
                    for (int sl = 1; sl < kk; ++sl) {
                        float x = vt[(nd * kk + sl) * 33 + cc];
                        if (x > mx) { mx = x; ix = sl; }
                        if (x < mn) { mn = x; in = sl; }
                    }
                    const int64_t o = (node0 + nd) * (int64_t)p.n_out + c;
                    p.vmax[o] = mx; p.vmin[o] = mn; p.imax[o] = (uint8_t)ix; p.imin[o] = (uint8_t)in;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        if (EPI != NT_EPI_BIAS) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int c = tid; c < g.n_tile; c += 128) {
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
    } else if (warp == 4) {
        // =========================== MMA issuer (one thread) ===========================
        if (lane == 0) {
            const uint32_t idesc = TF32 ? make_idesc_tf32(TC_M, (uint32_t)g.n_tile, 0, 0)
                                        : make_idesc_bf16(TC_M, (uint32_t)g.n_tile, 0, 0);
            const uint32_t lbo_a = TC_M * 16, lbo_b = (uint32_t)g.n_tile * 16, sbo = 128;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                const int s = kb & 1, use = kb >> 1;
                mbar_wait(&full[s], use & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(stage_base[s]), a_lo = a_hi + TC_A_BYTES;
                const uint32_t b_hi = a_lo + TC_A_BYTES, b_lo = b_hi + 4 * lbo_b;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t dah = make_smem_desc(a_hi + kk * 2 * lbo_a, lbo_a, sbo);
                    const uint64_t dal = make_smem_desc(a_lo + kk * 2 * lbo_a, lbo_a, sbo);
                    const uint64_t dbh = make_smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, sbo);
                    const uint64_t dbl = make_smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, sbo);
                    if (TF32) {
                        umma_tf32(tmem_base, dah, dbh, idesc, (kb | kk) ? 1u : 0u);
                        umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_base, dal, dbh, idesc, 1u);
                    } else {
                        umma_bf16(tmem_base, dah, dbh, idesc, (kb | kk) ? 1u : 0u);
                        umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                        umma_bf16(tmem_base, dal, dbh, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);            // stage reusable once these MMAs have read it
            }
            umma_commit(tmem_full);                // accumulator complete -> epilogue
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

template <int PROD, int EPI, bool TF32>
static int launch_tc(const NTParams &p, const void *w_split, cudaStream_t st) {
    const TCGeom g = tc_geometry(p.n_out, p.K, TF32 ? NT_PREC_TF32X3 : NT_PREC_BF16X3);
    const size_t smem = TC_STAGES * tc_stage_bytes(g.n_tile) + 64 + 6 * 256 * sizeof(float);
    static bool configured = false;     // per instantiation; the attribute is idempotent, races are harmless
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nt_tc_kernel<PROD, EPI, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(TC_STAGES * tc_stage_bytes(256) + 64 + 6 * 256 * sizeof(float)));
        if (e != cudaSuccess) return fail("nt_gemm_nt(tc): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid((unsigned)((p.rows + p.rows_per_tile - 1) / p.rows_per_tile), g.n_tiles);
    gemm_nt_tc_kernel<PROD, EPI, TF32><<<grid, TC_THREADS, smem, st>>>(p, reinterpret_cast<const uint8_t *>(w_split), g);
    return check_launch("nt_gemm_nt(tc)");
}

template <bool TF32>
static int dispatch_tc(const NTParams &p, int producer, int epilogue, const void *w_split, cudaStream_t st) {
    const bool edge = producer == NT_PROD_EDGE;
    switch (epilogue) {
        case NT_EPI_BIAS: return launch_tc<NT_PROD_PLAIN, NT_EPI_BIAS, TF32>(p, w_split, st);
        case NT_EPI_RELU_STATS:
            return edge ? launch_tc<NT_PROD_EDGE, NT_EPI_RELU_STATS, TF32>(p, w_split, st)
                        : launch_tc<NT_PROD_PLAIN, NT_EPI_RELU_STATS, TF32>(p, w_split, st);
        case NT_EPI_RELU_MAXMIN:
            return edge ? launch_tc<NT_PROD_EDGE, NT_EPI_RELU_MAXMIN, TF32>(p, w_split, st)
                        : launch_tc<NT_PROD_PLAIN, NT_EPI_RELU_MAXMIN, TF32>(p, w_split, st);
        default: return launch_tc<NT_PROD_PLAIN, NT_EPI_BNRELU_BWD, TF32>(p, w_split, st);
    }
}

// Engine choice for the TF32x3 row GEMMs, PER CALL (nt_gemm_args.engine; the library keeps no mutable state): 0 = auto -- the
// streaming engine (gemm_tc3.cu) for eligible calls with at least TC3_MIN_ROWS rows, the one-tile-per-CTA engine (this file)
// otherwise; 1 = always this file; 3 / 4 / 5 = first-generation streaming engine whenever eligible, whatever the row count, with
// one / two row tiles per weight stage / the aux-row ring for BNRELU_BWD; 6 = second-generation streaming engine (gemm_tc4.cu),
// which is also what auto picks first (tests compare them; results are bit-identical across engines).
constexpr int64_t TC3_MIN_ROWS = 32768;

bool nt_tc_would_stream(const NTParams &p, int producer, int epilogue, int precision) {
    if (precision != NT_PREC_TF32X3) return false;
    if (!(p.engine >= 3 || (p.engine == 0 && p.rows >= TC3_MIN_ROWS))) return false;
    return tc3_eligible(p, producer, epilogue);
}

int launch_nt_tc(const NTParams &p, int producer, int epilogue, int precision, const void *w_split, cudaStream_t st) {
    if (precision == NT_PREC_TF32X3 && (p.engine >= 3 || (p.engine == 0 && p.rows >= TC3_MIN_ROWS))) {
        if (p.engine == 0 || p.engine == 6) {  // second-generation streaming engine (gemm_tc4.cu: A operand in TMEM, TMA epilogue)
            const int rc4 = launch_nt_tc4(p, producer, epilogue, w_split, st);
            if (rc4 >= 0) return rc4;
        }
        const int rc = launch_nt_tc3(p, producer, epilogue, w_split, st);
        if (rc >= 0) return rc;
    }
    if (p.scatter) return fail("nt_gemm_nt: the fused scatter epilogue needs the streaming engine (see nt_gemm_nt_scatter_supported)%s", "");
    return precision == NT_PREC_TF32X3 ? dispatch_tc<true>(p, producer, epilogue, w_split, st)
                                       : dispatch_tc<false>(p, producer, epilogue, w_split, st);
}

}  // namespace nt

using namespace nt;

extern "C" int64_t nt_gemm_weights_bytes(int n_out, int K, int precision) {
    if (n_out < 1 || K < 1) return 0;
    const TCGeom g = tc_geometry(n_out, K, precision);
    return (int64_t)g.n_tiles * g.num_kb * g.n_tile * 128;
}

extern "C" int nt_gemm_prepare_weights(const float *w, int ldw, int n_out, int K, int precision, void *w_split,
                                       void *stream) {
    NT_REQUIRE(w && w_split && n_out >= 1 && K >= 1 && ldw >= K, "nt_gemm_prepare_weights: bad arguments");
    NT_REQUIRE(precision == NT_PREC_BF16X3 || precision == NT_PREC_TF32X3, "nt_gemm_prepare_weights: bad precision");
    NT_REQUIRE((reinterpret_cast<uintptr_t>(w_split) & 15u) == 0, "nt_gemm_prepare_weights: w_split must be 16-byte aligned");
    const TCGeom g = tc_geometry(n_out, K, precision);
    const int64_t total = (int64_t)g.n_tiles * g.num_kb * 4 * g.n_tile;
    tc_prepare_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        w, ldw, n_out, K, g, reinterpret_cast<uint4 *>(w_split));
    return check_launch("nt_gemm_prepare_weights");
}
