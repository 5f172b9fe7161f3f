// Streaming engine for the big fused row GEMMs (same contract as gemm_tc.cu / nt_gemm_nt), sm_100a.
//
//   out = epilogue( A . W^T )      A: [rows, K] fp32 rows in HBM (plain producer), W: [n_out <= 224, K], TF32x3
//
// These GEMMs have K, n_out <= 200: they are streaming kernels (read A once, write the result once; BNRELU_BWD also reads
// the saved activation once), so the design goal is bytes in flight, not flops.  One persistent CTA per SM loops over
// 128-row tiles with four decoupled pipelines:
//
//   warps 0-3   A loaders/converters.  Every thread keeps P3_DEPTH k-blocks (16 columns each) of ITS OWN 4 x 16-byte row
//               chunks in flight with cp.async.cg (global -> shared, no registers, zero-fill for ragged rows / K tails):
//               6 x 8 KB per CTA are always outstanding, independent of what the epilogue is doing and across tile
//               boundaries.  A landed k-block is split hi/lo (TF32x3) from shared memory into the UMMA core-matrix stage.
//   warp 12     single-thread tcgen05.mma issue into one of TWO TMEM accumulators (2 x 256 columns).
//   warps 4-7   epilogue group 0 (tiles 0, 2, 4, ... of this CTA; TMEM accumulator 0)
//   warps 8-11  epilogue group 1 (tiles 1, 3, 5, ...;              TMEM accumulator 1)
//               each group has two tile periods to drain its accumulator; all global traffic of the epilogue is
//               coalesced 16-byte accesses exchanged with the thread-per-row TMEM layout through a 32 x 36 transposition
//               tile per warp, and the BNRELU_BWD aux rows of chunk c+1 are loaded while chunk c is processed.
//
// The rows of a tile are visited in 64..128-byte pieces over a whole tile period, so every 32-row slab of A (and of the aux
// operand) is also prefetched into L2 with one cp.async.bulk.prefetch when the tile starts.  Build with -DNT_TC3_TRACE for a
// per-role cycle trace (tools/tc3_trace.py; findings in DESIGN.md section 4: the forward GEMMs are bound by the converter warps,
// BNRELU_BWD by the epilogue).
//
// The one-tile-per-CTA engine (gemm_tc.cu) remains the general path (gathered operands, unaligned rows, n_out > 224, small
// row counts); launch_nt_tc3 returns -1 when a call is not eligible.
#include "gemm_tc_shared.cuh"
#include <stdlib.h>

namespace nt {

constexpr int P3_EPI0_WARP = 4;
constexpr int P3_MMA_WARP = 12;
constexpr int P3_THREADS = 13 * 32;
// TILES = row tiles that share one weight k-block stage.  1: two TMEM accumulators are double-buffered between the MMA warp and
// the epilogue groups (tile t+1 is multiplied while tile t drains).  2: both accumulators belong to the same pass, the two
// tiles are multiplied against the SAME weight stage (half the L2 weight traffic per row -- the binding resource, DESIGN.md
// section 4) and drained together by the two epilogue groups while the loaders already stage the next pass.
// CFG 1: one tile per stage, 3 stages, 6 raw k-blocks in flight (48 KB of A per SM)
// CFG 2: two tiles per stage, 2 stages, 3 raw k-blocks (x2 tiles) in flight (48 KB)
// CFG 3: one tile per stage, 2 stages, 12 raw k-blocks in flight (96 KB)
// CFG 4: (BNRELU_BWD only; EXPERIMENTAL, written after the GPU budget of round 1 was spent -- never run yet) one tile per
//        stage, 2 stages, 3 raw k-blocks, and the aux rows of the epilogue staged through a 3-slot cp.async shared-memory ring
//        per warp (two 32-column chunks ahead, no registers) instead of the one-chunk-ahead register prefetch: the cycle trace
//        shows the epilogue of this GEMM waiting on exactly those loads.  The slot of chunk c doubles as its transposition tile.
template <int CFG> struct P3Cfg {
    static constexpr int TILES = CFG == 2 ? 2 : 1;
    static constexpr int STAGES = CFG == 1 ? 3 : 2;
    static constexpr int DEPTH = CFG == 1 ? 6 : ((CFG == 2 || CFG == 4) ? 3 : 12);   // k-blocks in flight per producer thread
    static constexpr int TW_SLOTS = CFG == 4 ? 3 : 1;                                 // transposition tiles per epilogue warp
};
constexpr int P3_SLAB = 4 * TC_M * 16;             // raw k-block of one tile: [4 chunks][128 rows][16 B]
constexpr int P3_TW = 32 * 36;                     // floats of one per-warp transposition tile

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Bulk prefetch of a contiguous global range into L2.  The kernel visits every operand row in 64..128-byte pieces spread over a
// whole tile period (one piece per k-block / per 32-column chunk), which DRAM serves with poor row-buffer locality; the rows of a
// tile are contiguous in memory, so one sequential prefetch per 32-row slab turns the piecewise visits into L2 hits.
__device__ __forceinline__ void l2_prefetch(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- optional cycle trace (built with -DNT_TC3_TRACE; developer tool, see tools/tc3_trace.py): per CTA, cycles spent by one
// representative thread of every role in each of its phases, accumulated over the kernel
#ifdef NT_TC3_TRACE
__device__ unsigned long long g_tc3_trace[256][16];
#define TRACE_T0() long long _t0 = clock64()
#define TRACE_ADD(slot) do { long long _t1 = clock64(); _acc[slot] += (unsigned long long)(_t1 - _t0); _t0 = _t1; } while (0)
#define TRACE_DECL() unsigned long long _acc[16] = {0}
#define TRACE_FLUSH(lo, hi) do { for (int _i = lo; _i < hi; ++_i) g_tc3_trace[blockIdx.x][_i] = _acc[_i]; } while (0)
#else
#define TRACE_T0()
#define TRACE_ADD(slot)
#define TRACE_DECL()
#define TRACE_FLUSH(lo, hi)
#endif

__host__ __device__ inline size_t tc3_stage_bytes(int n_tile, int tiles) { return (size_t)tiles * 2 * TC_A_BYTES + (size_t)n_tile * 128; }
__host__ __device__ inline size_t tc3_smem_bytes(int n_tile, int cfg = 1) {
    const int tiles = cfg == 2 ? 2 : 1, stages = cfg == 1 ? 3 : 2, depth = cfg == 1 ? 6 : ((cfg == 2 || cfg == 4) ? 3 : 12);
    const int tw_slots = cfg == 4 ? 3 : 1;
    return stages * tc3_stage_bytes(n_tile, tiles) + (size_t)depth * tiles * P3_SLAB + (size_t)8 * tw_slots * P3_TW * 4 +
           4 * 256 * 4 + 512 * 4 + 128;
}

template <int EPI, bool SCAT, int CFG>
__global__ void __launch_bounds__(P3_THREADS, 1) gemm_nt_tc3_kernel(NTParams p, const uint8_t *__restrict__ w_split, TCGeom g) {
    constexpr int TILES = P3Cfg<CFG>::TILES, P3_STAGES = P3Cfg<CFG>::STAGES, P3_DEPTH = P3Cfg<CFG>::DEPTH;
    constexpr int TW_SLOTS = P3Cfg<CFG>::TW_SLOTS;
    constexpr bool AUXRING = (CFG == 4) && (EPI == NT_EPI_BNRELU_BWD);
    constexpr int A_STAGE = TILES * 2 * TC_A_BYTES;                                // A part of a stage: [tile][hi|lo]
    extern __shared__ __align__(128) uint8_t smem[];
    const size_t stage_bytes = tc3_stage_bytes(g.n_tile, TILES);
    uint8_t *raw = smem + P3_STAGES * stage_bytes;
    float *tw_all = reinterpret_cast<float *>(raw + P3_DEPTH * TILES * P3_SLAB);  // 8 x [32][36]
    float *colv = tw_all + 8 * TW_SLOTS * P3_TW;                                   // [4][256]: bias | k0 | k1 | mu
    float *red = colv + 4 * 256;                                                   // [2][256] column statistics
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 512);                      // [3]
    uint64_t *empty = full + P3_STAGES;                                            // [3]
    uint64_t *tmem_full = empty + P3_STAGES;                                       // [2]
    uint64_t *tmem_empty = tmem_full + 2;                                          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_row_tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
    const int64_t n_passes = (n_row_tiles + TILES - 1) / TILES;                              // a pass = TILES consecutive row tiles
    const int my_tiles = (int)((n_passes - blockIdx.x + gridDim.x - 1) / gridDim.x);         // passes blockIdx.x, +grid, ...

    for (int i = tid; i < 512; i += P3_THREADS) red[i] = 0.f;
    for (int i = tid; i < 256; i += P3_THREADS) {
        const bool ok = i < p.n_out;
        colv[i] = (ok && p.bias) ? __ldg(p.bias + i) : 0.f;
        colv[256 + i] = (ok && EPI == NT_EPI_BNRELU_BWD) ? __ldg(p.k0 + i) : 0.f;
        colv[512 + i] = (ok && EPI == NT_EPI_BNRELU_BWD) ? __ldg(p.k1 + i) : 0.f;
        colv[768 + i] = (ok && EPI == NT_EPI_BNRELU_BWD) ? __ldg(p.mu + i) : 0.f;
    }
    if (warp == P3_MMA_WARP && lane == 0) {
        for (int s = 0; s < P3_STAGES; ++s) { mbar_init(&full[s], 128 + 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], TILES == 1 ? 128 : 256); }
        mbar_fence_init();
    }
    if (warp == P3_MMA_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_it = my_tiles * g.num_kb;

    if (warp < 4) {
        // =========================== A loaders / converters ===========================
        // lane -> 16-byte chunk jj = lane >> 3 of the rows warp*32 + 8*i + (lane & 7): a warp request touches 8 rows x 64
        // contiguous bytes, and the 16-byte shared-memory accesses of a quarter-warp cover 128 contiguous bytes.
        const int jj = lane >> 3;
        int prow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) prow[i] = warp * 32 + 8 * i + (lane & 7);
        const uint32_t raw_u32 = smem_u32(raw);
        const uint32_t slot_off = (uint32_t)(jj * (TC_M * 16));

        // fetch stream position (pass, k-block) -- advanced incrementally, it runs P3_DEPTH iterations ahead of consume.  The row
        // pointers of a pass are computed once (k-block 0); per k-block only the column offset and the K-tail size change
        // (ncu: the converter warps, not the tensor pipe or L2, set the pace of this kernel -- keep their instruction count low).
        int f_it = 0, f_kb = 0, f_slab = 0;
        int64_t f_row0 = (int64_t)blockIdx.x * TILES * p.rows_per_tile;           // first row of the pass being fetched
        const float *rowp[TILES][4];
        auto issue = [&]() {
            if (f_it < total_it) {
                if (f_kb == 0) {
#pragma unroll
                    for (int tt = 0; tt < TILES; ++tt) {
                        const int64_t t_row0 = f_row0 + (int64_t)tt * p.rows_per_tile;
                        const int rows_here = (int)max((int64_t)0, min((int64_t)p.rows_per_tile, p.rows - t_row0));
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            rowp[tt][i] = prow[i] < rows_here ? p.a + (t_row0 + prow[i]) * (int64_t)p.lda : nullptr;
                        const int slab_rows = min(32, rows_here - warp * 32);
                        if (lane == 0 && slab_rows > 0) {    // up to the last USED column of the last row (never past the operand)
                            const uint32_t pf = (uint32_t)(((slab_rows - 1) * p.lda + p.K) * 4) & ~15u;
                            if (pf) l2_prefetch(p.a + (t_row0 + warp * 32) * (int64_t)p.lda, pf);
                        }
                    }
                }
                const int k = (f_kb * 4 + jj) * 4;
                const uint32_t kbytes = k < p.K ? (uint32_t)min(16, (p.K - k) * 4) : 0u;
#pragma unroll
                for (int tt = 0; tt < TILES; ++tt) {
                    const uint32_t dst0 = raw_u32 + (uint32_t)((f_slab * TILES + tt) * P3_SLAB) + slot_off;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float *rp = rowp[tt][i];
                        cp_async16(dst0 + (uint32_t)(prow[i] * 16), rp ? rp + k : p.a, rp ? kbytes : 0u);
                    }
                }
                ++f_it;
                if (++f_slab == P3_DEPTH) f_slab = 0;
                if (++f_kb == g.num_kb) { f_kb = 0; f_row0 += (int64_t)gridDim.x * TILES * p.rows_per_tile; }
            }
            cp_async_commit();                       // exactly one group per call, possibly empty
        };
#pragma unroll 1
        for (int d = 0; d < P3_DEPTH; ++d) issue();

        int s = 0, use = 0, kb = 0, slab = 0;
        TRACE_DECL();
        TRACE_T0();
#pragma unroll 1
        for (int it = 0; it < total_it; ++it) {
            cp_async_wait<P3_DEPTH - 1>();           // this thread's chunks of k-block `it` have landed
            TRACE_ADD(0);
            mbar_wait(&empty[s], (use & 1) ^ 1);
            TRACE_ADD(1);
            uint8_t *stage = smem + s * stage_bytes, *b_all = stage + A_STAGE;
            {
                // weight k-block: one bulk copy per producer warp (4 concurrent requests of n_tile*32 bytes; n_tile % 16 == 0 keeps
                // them 16-byte granular).  tid 0 registers the whole transaction; the others only add their complete_tx.
                const uint32_t bytes = (uint32_t)g.n_tile * 128u, part = bytes >> 2;
                if (tid == 0) mbar_arrive_expect_tx(&full[s], bytes);
                if (lane == 0) bulk_g2s(b_all + warp * part, w_split + (size_t)kb * bytes + warp * part, part, &full[s]);
            }
#pragma unroll
            for (int tt = 0; tt < TILES; ++tt) {
                uint8_t *a_hi = stage + tt * 2 * TC_A_BYTES, *a_lo = a_hi + TC_A_BYTES;
                const uint8_t *rs = raw + (slab * TILES + tt) * P3_SLAB + slot_off;
                // all four raw chunks are read BEFORE the first store: the compiler cannot prove that the stage stores do not alias
                // the raw slab (same shared array) and would otherwise serialise load -> split -> store four times (SASS of the first
                // version: one LDS.128 latency exposed per chunk)
                float4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4 *>(rs + prow[i] * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                    split_tf32(x[i].x, h0, l0); split_tf32(x[i].y, h1, l1); split_tf32(x[i].z, h2, l2); split_tf32(x[i].w, h3, l3);
                    *reinterpret_cast<uint4 *>(a_hi + slot_off + prow[i] * 16) = make_uint4(h0, h1, h2, h3);
                    *reinterpret_cast<uint4 *>(a_lo + slot_off + prow[i] * 16) = make_uint4(l0, l1, l2, l3);
                }
            }
            fence_proxy_async();                     // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&full[s]);
            TRACE_ADD(2);
            issue();                                 // refill the raw slab this thread has just read
            TRACE_ADD(3);
            if (++s == P3_STAGES) { s = 0; ++use; }
            if (++kb == g.num_kb) kb = 0;
            if (++slab == P3_DEPTH) slab = 0;
        }
        cp_async_wait<0>();
        if (tid == 0) TRACE_FLUSH(0, 4);
    } else if (warp == P3_MMA_WARP) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, (uint32_t)g.n_tile, 0, 0);
            const uint32_t lbo_a = TC_M * 16, lbo_b = (uint32_t)g.n_tile * 16, sbo = 128;
            int s = 0, use = 0;
            TRACE_DECL();
            TRACE_T0();
            for (int ti = 0; ti < my_tiles; ++ti) {
                // TILES == 1: accumulator ti & 1, handed over per tile; TILES == 2: both accumulators, handed over per pass
                const int bar = TILES == 1 ? (ti & 1) : 0, ause = TILES == 1 ? (ti >> 1) : ti;
                mbar_wait(&tmem_empty[bar], (ause & 1) ^ 1);        // the epilogue group(s) have drained the accumulator(s)
                tc_fence_after();
                TRACE_ADD(4);
                for (int kb = 0; kb < g.num_kb; ++kb) {
                    mbar_wait(&full[s], use & 1);
                    tc_fence_after();
                    TRACE_ADD(5);
                    const uint32_t stage = smem_u32(smem + s * stage_bytes);
                    const uint32_t b_hi = stage + A_STAGE, b_lo = b_hi + 4 * lbo_b;
#pragma unroll
                    for (int tt = 0; tt < TILES; ++tt) {
                        const uint32_t d = tmem_base + (uint32_t)((TILES == 1 ? (ti & 1) : tt) * 256);
                        const uint32_t a_hi = stage + tt * 2 * TC_A_BYTES, a_lo = a_hi + TC_A_BYTES;
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
                    }
                    umma_commit(&empty[s]);            // stage reusable once these MMAs have read it
                    if (++s == P3_STAGES) { s = 0; ++use; }
                    TRACE_ADD(6);
                }
                umma_commit(&tmem_full[bar]);          // accumulator(s) complete -> epilogue
            }
            TRACE_FLUSH(4, 7);
        }
    } else {
        // =========================== epilogue groups (thread = row of the tile, TMEM lane = row) ===========================
        const int grp = (warp - P3_EPI0_WARP) >> 2;                // 0: warps 4-7, 1: warps 8-11
        const int quad = warp & 3;                                  // TMEM lane quadrant this warp may read
        const int et = quad * 32 + lane;                            // 0..127 inside the group = row of the tile
        // transposition tiles: [group][slot][quad][32 x 36]; the four tiles of a (group, slot) are contiguous: row r at r*36
        float *twg = tw_all + (grp * TW_SLOTS) * (4 * P3_TW);
        float *tw4 = twg + quad * P3_TW;
        const int sub = lane >> 3, q4 = (lane & 7) * 4;
        const int n_chunks = (g.n_tile + 31) / 32;
        TRACE_DECL();
        TRACE_T0();
        for (int ti = (TILES == 1 ? grp : 0); ti < my_tiles; ti += (TILES == 1 ? 2 : 1)) {
            // group `grp` always drains accumulator `grp`: TILES == 1 -> every other tile; TILES == 2 -> tile `grp` of every pass
            const int as = grp, bar = TILES == 1 ? grp : 0, ause = TILES == 1 ? (ti >> 1) : ti;
            const int64_t pass = (int64_t)blockIdx.x + (int64_t)ti * gridDim.x;
            const int64_t row0 = (pass * TILES + (TILES == 1 ? 0 : grp)) * p.rows_per_tile;
            const int rows_here = (int)max((int64_t)0, min((int64_t)p.rows_per_tile, p.rows - row0));   // 0: phantom tile
            const bool valid = et < rows_here;
            const int64_t wrow0 = row0 + quad * 32;                  // first row of this warp
            const int wrows = max(0, min(32, rows_here - quad * 32));
            const float *auxw = (EPI == NT_EPI_BNRELU_BWD) ? p.aux + wrow0 * (int64_t)p.ldaux : nullptr;
            float4 ld[8];
            // coalesced aux read of one 32-column chunk: lane -> row 4m + sub, columns q4 .. q4+3
            auto load_aux = [&](int ch) {
                const int cg = ch * 32, nv = min(32, p.n_out - cg);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int rr = 4 * m + sub;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rr < wrows) {
                        const float *src = auxw + (int64_t)rr * p.ldaux + cg + q4;
                        if (q4 + 3 < nv) {
                            a = __ldg(reinterpret_cast<const float4 *>(src));
                        } else if (q4 < nv) {
                            a.x = __ldg(src);
                            if (q4 + 1 < nv) a.y = __ldg(src + 1);
                            if (q4 + 2 < nv) a.z = __ldg(src + 2);
                        }
                    }
                    ld[m] = a;
                }
            };
            // AUXRING: chunk `ch` of the aux rows -> slot ch % 3 of this warp's ring (layout of the transposition tile: row 4m+sub,
            // columns q4..q4+3), zero-filled outside the tile / the valid columns; exactly one cp.async group per call
            auto issue_aux = [&](int ch) {
                if (ch < n_chunks) {
                    const int cg = ch * 32, nv = min(32, p.n_out - cg);
                    const uint32_t dst0 = smem_u32(tw_all + ((grp * TW_SLOTS + ch % TW_SLOTS) * 4 + quad) * P3_TW);
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int rr = 4 * m + sub;
                        const bool live = rr < wrows && q4 < nv;
                        const uint32_t nbytes = live ? (uint32_t)min(16, (nv - q4) * 4) : 0u;
                        const float *src = live ? auxw + (int64_t)rr * p.ldaux + cg + q4 : p.aux;
                        cp_async16(dst0 + (uint32_t)((rr * 36 + q4) * 4), src, nbytes);
                    }
                }
                cp_async_commit();
            };
            if (EPI == NT_EPI_BNRELU_BWD) {
                // L2 prefetch of this warp's aux slab of the CURRENT tile (prefetching the group's next tile instead -- a whole tile
                // period ahead -- measured slower: 1.43 vs 1.32 ms per step for the group; the lines do not survive in L2)
                if (lane == 0 && wrows > 0) {
                    const uint32_t pf = (uint32_t)(((wrows - 1) * p.ldaux + p.n_out) * 4) & ~15u;
                    if (pf) l2_prefetch(auxw, pf);
                }
                if (AUXRING) { issue_aux(0); issue_aux(1); }
                else load_aux(0);                                    // does not depend on the accumulator
            }
            // fused edge scatter: global neighbour row of each of this lane's 8 rows (4m + sub) of the warp slab
            int jrow[SCAT ? 8 : 1];
            if (SCAT) {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int rr = 4 * m + sub;
                    jrow[m] = 0;
                    if (rr < wrows) {
                        const int64_t e = wrow0 + rr, centre = e / p.e.k;
                        jrow[m] = (int)((centre / p.e.n_per_cloud) * (int64_t)p.e.n_per_cloud + __ldg(p.e.idx + e));
                    }
                }
            }
            TRACE_ADD(9 + grp * 3);                                  // tile prologue (aux prefetch, index loads)
            mbar_wait(&tmem_full[bar], ause & 1);
            tc_fence_after();
            TRACE_ADD(7 + grp * 3);                                  // waiting for the accumulator
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int c0 = ch * 32;
                const int nv = min(32, p.n_out - c0);                // valid columns of this chunk (>= 1)
                float auxv[32];
                if (AUXRING) {                                       // this chunk's ring slot is also its transposition tile
                    twg = tw_all + (grp * TW_SLOTS + ch % TW_SLOTS) * (4 * P3_TW);
                    tw4 = twg + quad * P3_TW;
                    cp_async_wait<1>();                              // chunk ch has landed (chunk ch+1 may still be in flight)
                    __syncwarp();
                } else if (EPI == NT_EPI_BNRELU_BWD) {
#pragma unroll
                    for (int m = 0; m < 8; ++m) *reinterpret_cast<float4 *>(tw4 + (4 * m + sub) * 36 + q4) = ld[m];
                    __syncwarp();
                }
                if (EPI == NT_EPI_BNRELU_BWD) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 a = *reinterpret_cast<const float4 *>(tw4 + lane * 36 + 4 * i);
                        auxv[4 * i] = a.x; auxv[4 * i + 1] = a.y; auxv[4 * i + 2] = a.z; auxv[4 * i + 3] = a.w;
                    }
                    __syncwarp();
                    if (AUXRING) issue_aux(ch + 2);                  // slot (ch+2) % 3 was last read in chunk ch-1
                    else if (ch + 1 < n_chunks) load_aux(ch + 1);    // in flight while this chunk is processed
                }
                float acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + c0), acc);
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
                const bool want = (EPI == NT_EPI_BIAS) ? false
                                                       : ((EPI == NT_EPI_BNRELU_BWD) ? (p.colsum != nullptr) : (p.stats != nullptr));
                constexpr bool scat = SCAT;
                if (p.out || want || scat || EPI == NT_EPI_RELU_MAXMIN) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4 *>(tw4 + lane * 36 + 4 * i) =
                            make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                    __syncwarp();
                    if (p.out) {
                        float *dst = p.out + wrow0 * (int64_t)p.ldo + c0 + q4;
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const int rr = 4 * m + sub;
                            if (rr < wrows) {
                                const float4 v = *reinterpret_cast<const float4 *>(tw4 + rr * 36 + q4);
                                float *d = dst + (int64_t)rr * p.ldo;
                                if (q4 + 3 < nv) {
                                    *reinterpret_cast<float4 *>(d) = v;
                                } else if (q4 < nv) {
                                    d[0] = v.x;
                                    if (q4 + 1 < nv) d[1] = v.y;
                                    if (q4 + 2 < nv) d[2] = v.z;
                                }
                            }
                        }
                    }
                    if (scat) {
                        // neighbour half: dpq[j, n_out + col] += dz[e, col] (16-byte reductions; zero vectors -- ReLU-masked
                        // rows -- are skipped); n_out % 4 == 0 is an eligibility condition, so every live quad is whole
                        float *qdst = p.scatter + p.n_out + c0 + q4;
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            const int rr = 4 * m + sub;
                            if (rr < wrows && q4 < nv) {
                                const float4 v = *reinterpret_cast<const float4 *>(tw4 + rr * 36 + q4);
                                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                                    atomicAdd(reinterpret_cast<float4 *>(qdst + (int64_t)jrow[SCAT ? m : 0] * p.ldscatter), v);
                            }
                        }
                        // centre half: dpq[c, col] = sum over the k edge rows of centre c (tiles hold whole centres: plain store)
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                        const int kk = p.e.k;
                        const int nodes_here = rows_here / kk;
                        const int64_t node0 = row0 / kk;
                        if (lane < nv) {
                            for (int t = et; t < nodes_here * 32; t += 128) {
                                const int nd = t >> 5;
                                float sum = 0.f;
                                for (int sl = 0; sl < kk; ++sl) sum += twg[(nd * kk + sl) * 36 + lane];
                                p.scatter[(node0 + nd) * (int64_t)p.ldscatter + c0 + lane] = sum;
                            }
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                    }
                    if (EPI == NT_EPI_RELU_MAXMIN) {
                        // max / min over the k edge rows of every centre point (nodes straddle warps: group barrier); the same
                        // pass yields the column statistics (a thread always lands on the same column: 128 % 32 == 0)
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                        const int kk = p.k_agg;
                        const int nodes_here = rows_here / kk;
                        const int64_t node0 = row0 / kk;
                        float t1 = 0.f, t2 = 0.f;
                        if (lane < nv) {
                            for (int t = et; t < nodes_here * 32; t += 128) {
                                const int nd = t >> 5;
                                float x = twg[(nd * kk) * 36 + lane];
                                float mx = x, mn = x;
                                int ix = 0, in = 0;
                                t1 += x; t2 = fmaf(x, x, t2);
                                for (int sl = 1; sl < kk; ++sl) {
                                    x = twg[(nd * kk + sl) * 36 + lane];
                                    if (x > mx) { mx = x; ix = sl; }
                                    if (x < mn) { mn = x; in = sl; }
                                    t1 += x; t2 = fmaf(x, x, t2);
                                }
                                const int64_t o = (node0 + nd) * (int64_t)p.n_out + c0 + lane;
                                p.vmax[o] = mx; p.vmin[o] = mn; p.imax[o] = (uint8_t)ix; p.imin[o] = (uint8_t)in;
                            }
                            if (p.stats) { atomicAdd(&red[c0 + lane], t1); atomicAdd(&red[256 + c0 + lane], t2); }
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                    } else {
                        if (want && lane < nv) {
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
                }
            }
            // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[bar]);
            TRACE_ADD(8 + grp * 3);                                  // draining
        }
        if (et == 0) TRACE_FLUSH(7 + grp * 3, 10 + grp * 3);
        // flush the per-CTA column statistics once (they accumulate over all of this CTA's tiles, both groups)
        if (EPI != NT_EPI_BIAS) {
            asm volatile("bar.sync 3, 256;" ::: "memory");
            for (int c = (warp - P3_EPI0_WARP) * 32 + lane; c < p.n_out; c += 256) {
                if (EPI == NT_EPI_BNRELU_BWD) {
                    if (p.colsum) atomicAdd(p.colsum + c, (double)red[c]);
                } else if (p.stats) {
                    atomicAdd(p.stats + c, (double)red[c]);
                    atomicAdd(p.stats + p.n_out + c, (double)red[256 + c]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == P3_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

template <int EPI, bool SCAT, int CFG>
static int launch_tc3_t(const NTParams &p, const void *w_split, const TCGeom &g, int sms, cudaStream_t st) {
    constexpr int TILES = P3Cfg<CFG>::TILES;
    const size_t smem = tc3_smem_bytes(g.n_tile, CFG);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nt_tc3_kernel<EPI, SCAT, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail("nt_gemm_nt(tc3): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int64_t n_row_tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
    const int64_t n_passes = (n_row_tiles + TILES - 1) / TILES;
    const int ctas = (int)(n_passes < sms ? n_passes : sms);
    gemm_nt_tc3_kernel<EPI, SCAT, CFG><<<ctas, P3_THREADS, smem, st>>>(p, reinterpret_cast<const uint8_t *>(w_split), g);
    return check_launch("nt_gemm_nt(tc3)");
}

// configuration (P3Cfg) from the call's engine: 3 -> one row tile per weight stage, 4 -> two, 5 or auto -> the aux-row ring for
// BNRELU_BWD (cfg 4; measured r2: 1.33 -> 1.09 ms per C2 step) and one tile per stage for the other epilogues
static int tc3_cfg(int engine) { return engine == 3 ? 1 : (engine == 4 ? 2 : 4); }

template <int EPI, bool SCAT = false>
static int launch_tc3(const NTParams &p, const void *w_split, const TCGeom &g, int sms, cudaStream_t st) {
    const int cfg = tc3_cfg(p.engine);       // configuration index (see P3Cfg)
    if (cfg == 2 && tc3_smem_bytes(g.n_tile, 2) <= 227 * 1024) return launch_tc3_t<EPI, SCAT, 2>(p, w_split, g, sms, st);
    if (cfg == 3 && tc3_smem_bytes(g.n_tile, 3) <= 227 * 1024) return launch_tc3_t<EPI, SCAT, 3>(p, w_split, g, sms, st);
    if constexpr (EPI == NT_EPI_BNRELU_BWD) {
        if (cfg == 4 && tc3_smem_bytes(g.n_tile, 4) <= 227 * 1024) return launch_tc3_t<EPI, SCAT, 4>(p, w_split, g, sms, st);
    }
    return launch_tc3_t<EPI, SCAT, 1>(p, w_split, g, sms, st);
}

bool tc3_eligible(const NTParams &p, int producer, int epilogue) {
    if (producer != NT_PROD_PLAIN || p.n_out > 224) return false;
    const TCGeom g = tc_geometry(p.n_out, p.K, NT_PREC_TF32X3);
    if (g.n_tiles != 1 || tc3_smem_bytes(g.n_tile) > 227 * 1024) return false;
    if ((p.lda & 3) != 0 || !aligned16(p.a)) return false;
    if (p.out && ((p.ldo & 3) != 0 || !aligned16(p.out))) return false;
    if (epilogue == NT_EPI_BNRELU_BWD && (p.aux_edge || (p.ldaux & 3) != 0 || !aligned16(p.aux))) return false;
    if (p.scatter) {
        if (epilogue != NT_EPI_BNRELU_BWD || !p.e.idx || p.e.k < 1 || p.e.k > TC_M || p.e.n_per_cloud < 1) return false;
        if (p.rows >= (int64_t)1 << 31) return false;
        if ((p.n_out & 3) != 0 || (p.ldscatter & 3) != 0 || p.ldscatter < 2 * p.n_out || !aligned16(p.scatter)) return false;
        if (p.rows % p.e.k != 0 || p.rows_per_tile % p.e.k != 0) return false;
    }
    return true;
}

// -1: not eligible (the caller falls back to the one-tile-per-CTA engine)
int launch_nt_tc3(const NTParams &p, int producer, int epilogue, const void *w_split, cudaStream_t st) {
    if (!tc3_eligible(p, producer, epilogue)) return -1;
    const TCGeom g = tc_geometry(p.n_out, p.K, NT_PREC_TF32X3);
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1)
            return fail("nt_gemm_nt(tc3): cannot query the SM count%s", "");
        sms = n;
    }
    switch (epilogue) {
        case NT_EPI_BIAS: return launch_tc3<NT_EPI_BIAS>(p, w_split, g, sms, st);
        case NT_EPI_RELU_STATS: return launch_tc3<NT_EPI_RELU_STATS>(p, w_split, g, sms, st);
        case NT_EPI_RELU_MAXMIN: return launch_tc3<NT_EPI_RELU_MAXMIN>(p, w_split, g, sms, st);
        default:
            return p.scatter ? launch_tc3<NT_EPI_BNRELU_BWD, true>(p, w_split, g, sms, st)
                             : launch_tc3<NT_EPI_BNRELU_BWD, false>(p, w_split, g, sms, st);
    }
}

}  // namespace nt

#ifdef NT_TC3_TRACE
extern "C" int nt_debug_tc3_trace(unsigned long long *host_out) {      // [256][16]
    return cudaMemcpyFromSymbol(host_out, nt::g_tc3_trace, sizeof(nt::g_tc3_trace)) == cudaSuccess ? 0 : 1;
}
#endif
