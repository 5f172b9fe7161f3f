// Streaming engine, second generation (nt_gemm_args.engine = 6, and what engine 0 = auto picks for eligible calls): the persistent tcgen05 row
// GEMM of gemm_tc3.cu with the two measured bottlenecks of that kernel taken out (cycle trace):
//
//   * the forward GEMMs were bound by their four converter warps (one warp per scheduler: LDS -> TF32 split -> 2 x STS ->
//     fence.proxy.async -> arrive is a serial latency chain).  Here EIGHT converter warps work on alternating k-blocks (two
//     independent chains per scheduler), and they no longer fetch: a loader thread brings every raw k-block in with ONE TMA
//     tensor-map load ([128 rows x 16 columns] box, 64B-swizzled, K tail and rows past the end zero-filled by the hardware) --
//     the 4 x 32-lane cp.async per k-block and thread cost the converters a third of their time (LSU back-pressure).
//   * the epilogues spent their time on per-thread global traffic: 8 cp.async (aux rows) + 8 LDS.128 + 8 STG.128 per thread and
//     32-column chunk, with the address / tail predicates around them.  Here every global access of the epilogue is a TMA
//     tensor-map copy issued by one lane per warp: `cp.async.bulk.tensor.2d` loads the [32 rows x 32 columns] aux box into a
//     128B-swizzled shared-memory tile (completion on an mbarrier, out-of-bounds rows / columns zero-filled by the hardware) and
//     the result box leaves through a TMA store (`cp.async.bulk.tensor.2d.global.shared::cta`, clipped at the tensor bounds --
//     no tail code).  A thread touches only ITS row of the tile (TMEM lane = row): it reads the aux values and writes the result
//     in place, 16-byte accesses whose chunk index is XOR-ed with (row & 7) -- the SWIZZLE_128B pattern -- so they are
//     conflict-free without padding.  Column statistics go to per-warp accumulators (no shared-memory atomics).
//   * shared-memory bandwidth was the next wall (per 128-row tile of the 200 -> 200 GEMM: 1.86 MB through the 128 B/cycle port =
//     14.5 k of the measured 19 k cycles): the split A operand was written to shared memory by the converters (205 KB) and read
//     back three times by the tensor core (307 KB).  Here the A operand lives in TENSOR MEMORY: a converter thread owns one row
//     (= one TMEM lane), splits its 16 fp32 values of the k-block and writes hi | lo with two tcgen05.st (one 32-bit column per
//     K element -- tools/microbench/ts_mode_test.cu: bit-identical to the shared-memory operand, second K = 8 step at +8
//     columns), and the MMAs read it from there (`tcgen05.mma [d], [a], b-desc`).  No generic-proxy store, no
//     fence.proxy.async on the operand path; the 32 KB of A stages become a third weight stage.  TMEM: two accumulators of
//     n_tile columns + three A stages of 32 columns (n_tile <= 208).
//   (Tried and dropped: independent rings for the A stages and FOUR weight stages filled by their own loader thread -- the MMA
//   thread waits 35-45 % of its time for full stages either way, 0.151 vs 0.142 ms: what it waits for is the weight stream itself,
//   n_tile*128 bytes per k-block and CTA from L2 = 3.2x the A bytes)
//
// Results are bit-identical to the other engines (same operand split, same MMA order, same epilogue arithmetic); the tests
// compare them.  Calls this file does not take (fused scatter, operands that are not 16-byte aligned) stay on gemm_tc3.cu.
#include "gemm_tc_shared.cuh"
#include <cuda.h>

namespace nt {

template <int EPI, int SETS_> struct P4Cfg {
    static constexpr bool BWD = EPI == NT_EPI_BNRELU_BWD;
    static constexpr int SETS = SETS_;                     // converter warp sets (4 warps each)
    static constexpr int STAGES = 3;                       // weight k-blocks in shared memory + A k-blocks (hi | lo) in tensor memory
    static constexpr int DEPTH = BWD ? 4 : 6;              // raw k-blocks in flight per CTA (multiple of SETS)
    static constexpr int SLOTS = BWD ? 3 : 2;              // swizzled [32][32] tiles per epilogue warp
    static constexpr int EPI0_WARP = 4 * SETS;
    static constexpr int MMA_WARP = EPI0_WARP + 8;
    static constexpr int LOAD_WARP = MMA_WARP + 1;         // one thread: TMA loads of the raw A k-blocks
    static constexpr int THREADS = (LOAD_WARP + 1) * 32;
};
constexpr int P4_SLAB = TC_M * 64;                         // raw k-block: [128 rows][16 fp32], 64B-swizzled TMA box
constexpr int P4_TILE = 32 * 32 * 4;                       // one epilogue tile (4 KB, 1024-byte aligned)

__device__ __forceinline__ void l2_prefetch_4(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// ---- TMA tensor-map copies (2-D boxes; coordinates = {column, row}) -------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int col, int row, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(col), "r"(row)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src_smem, int col, int row) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(src_smem), "r"(col), "r"(row)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from tensor memory (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: 16 consecutive columns of this thread's lane (warp w writes lanes 32 (w % 4) .. +31); no wait inside
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#ifdef NT_TC3_TRACE
__device__ unsigned long long g_tc4_trace[256][16];
#define TRACE_T0() long long _t0 = clock64()
#define TRACE_ADD(slot) do { long long _t1 = clock64(); _acc[slot] += (unsigned long long)(_t1 - _t0); _t0 = _t1; } while (0)
#define TRACE_DECL() unsigned long long _acc[16] = {0}
#define TRACE_FLUSH(lo, hi) do { for (int _i = lo; _i < hi; ++_i) g_tc4_trace[blockIdx.x][_i] = _acc[_i]; } while (0)
#else
#define TRACE_T0()
#define TRACE_ADD(slot)
#define TRACE_DECL()
#define TRACE_FLUSH(lo, hi)
#endif

constexpr int P4_MAX_NTILE = 208;                          // 2 accumulators x n_tile + 3 A stages x 32 columns <= 512 TMEM columns
__host__ __device__ inline size_t tc4_stage_bytes(int n_tile) { return (size_t)n_tile * 128; }      // weight k-block: hi | lo planes
// 1 KB alignment slack | epilogue tiles | operand stages | raw slabs | column constants [4][256] | per-warp statistics [8][512] |
// barriers
__host__ __device__ inline size_t tc4_smem_bytes(int n_tile, bool bwd) {
    const int stages = 3, depth = bwd ? 4 : 6, slots = bwd ? 3 : 2;
    return 1024 + (size_t)8 * slots * P4_TILE + stages * tc4_stage_bytes(n_tile) + (size_t)depth * P4_SLAB + 4 * 256 * 4 +
           (size_t)8 * (bwd ? 256 : 512) * 4 + 512;
}

// RELU_MAXMIN: thread (row slot et, column) scans the k edge rows of the centre points et / 32, et / 32 + 4, ... of one 128-row
// group tile (element (R, c) at float index R*32 + (((c >> 2) ^ (R & 7)) << 2) + (c & 3)): max / min / their slots + column sums.
// KC > 0: compile-time k (all loads of a point issued before the compare chain).
template <int KC>
__device__ __forceinline__ void maxmin_scan(const float *gt, int kk_rt, int et, int cq, int cr, int nodes_here, int64_t node0,
                                            const NTParams &p, int col, float &t1, float &t2) {
    const int kk = KC ? KC : kk_rt;
    for (int t = et; t < nodes_here * 32; t += 128) {
        const int nd = t >> 5;
        const int R0 = nd * kk;
        float mx, mn;
        int ix = 0, in = 0;
        if (KC) {
            float xs[KC ? KC : 1];
#pragma unroll
            for (int sl = 0; sl < (KC ? KC : 1); ++sl) { const int R = R0 + sl; xs[sl] = gt[R * 32 + ((cq ^ (R & 7)) << 2) + cr]; }
            mx = mn = xs[0];
            t1 += xs[0]; t2 = fmaf(xs[0], xs[0], t2);
#pragma unroll
            for (int sl = 1; sl < (KC ? KC : 1); ++sl) {
                const float x = xs[sl];
                if (x > mx) { mx = x; ix = sl; }
                if (x < mn) { mn = x; in = sl; }
                t1 += x; t2 = fmaf(x, x, t2);
            }
        } else {
            float x = gt[R0 * 32 + ((cq ^ (R0 & 7)) << 2) + cr];
            mx = mn = x;
            t1 += x; t2 = fmaf(x, x, t2);
            for (int sl = 1; sl < kk; ++sl) {
                const int R = R0 + sl;
                x = gt[R * 32 + ((cq ^ (R & 7)) << 2) + cr];
                if (x > mx) { mx = x; ix = sl; }
                if (x < mn) { mn = x; in = sl; }
                t1 += x; t2 = fmaf(x, x, t2);
            }
        }
        const int64_t o = (node0 + nd) * (int64_t)p.n_out + col;
        p.vmax[o] = mx; p.vmin[o] = mn; p.imax[o] = (uint8_t)ix; p.imin[o] = (uint8_t)in;
    }
}

template <int EPI, int SETS_>
__global__ void __launch_bounds__(P4Cfg<EPI, SETS_>::THREADS, 1)
gemm_nt_tc4_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_out,
                   const __grid_constant__ CUtensorMap tm_out_tail, const __grid_constant__ CUtensorMap tm_aux, NTParams p, const uint8_t *__restrict__ w_split, TCGeom g) {
    using C = P4Cfg<EPI, SETS_>;
    constexpr bool BWD = C::BWD;
    constexpr int SETS = C::SETS, STAGES = C::STAGES, DEPTH = C::DEPTH, SLOTS = C::SLOTS;
    constexpr int STATW = BWD ? 256 : 512;                                       // floats of statistics per epilogue warp
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B tiles need 1 KB alignment
    uint8_t *tiles = smem;                                                       // [group][slot][quad][32 x 128 B]
    uint8_t *raw = tiles + 8 * SLOTS * P4_TILE;                                  // [DEPTH][128 rows x 64 B] (8 KB aligned)
    uint8_t *stages = raw + DEPTH * P4_SLAB;
    const size_t stage_bytes = tc4_stage_bytes(g.n_tile);
    float *colv = reinterpret_cast<float *>(stages + STAGES * stage_bytes);      // [4][256]: bias | k0 | k1 | mu
    float *wstat = colv + 4 * 256;                                               // [8 epilogue warps][STATW]
    uint64_t *full = reinterpret_cast<uint64_t *>(wstat + 8 * STATW);            // [STAGES]
    uint64_t *empty = full + STAGES;                                             // [STAGES]
    uint64_t *tmem_full = empty + STAGES;                                        // [2]
    uint64_t *tmem_empty = tmem_full + 2;                                        // [2]
    uint64_t *aux_bar = tmem_empty + 2;                                          // [8 epilogue warps][SLOTS]
    uint64_t *raw_full = aux_bar + 8 * 3;                                        // [DEPTH]
    uint64_t *raw_empty = raw_full + DEPTH;                                      // [DEPTH]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(raw_empty + DEPTH);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_row_tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
    const int my_tiles = (int)((n_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles blockIdx.x, +grid, ...

    for (int i = tid; i < 8 * STATW; i += C::THREADS) wstat[i] = 0.f;
    for (int i = tid; i < 256; i += C::THREADS) {
        const bool ok = i < p.n_out;
        colv[i] = (ok && p.bias) ? __ldg(p.bias + i) : 0.f;
        colv[256 + i] = (ok && BWD) ? __ldg(p.k0 + i) : 0.f;
        colv[512 + i] = (ok && BWD) ? __ldg(p.k1 + i) : 0.f;
        colv[768 + i] = (ok && BWD) ? __ldg(p.mu + i) : 0.f;
    }
    if (warp == C::MMA_WARP && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 128 + 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
        for (int i = 0; i < 8 * 3; ++i) mbar_init(&aux_bar[i], 1);
        for (int i = 0; i < DEPTH; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 128); }
        mbar_fence_init();
    }
    if (warp == C::MMA_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_it = my_tiles * g.num_kb;

    if (warp < C::EPI0_WARP) {
        // =========================== A loaders / converters ===========================
        // set q = warp >> 2 handles the k-block stream positions q, q + SETS, ...; inside a set, warp cw owns rows 32 cw .. 32 cw + 31
        // and lane -> row cw*32 + lane (= the TMEM lane this warp may write).  The raw k-block arrives by TMA (loader warp) as
        // [128 rows][64 B] with the 16-byte chunks of a row XOR-ed with (row >> 1) & 3 (SWIZZLE_64B): the four LDS.128 of a
        // thread are conflict-free.  Columns >= K and rows >= rows arrive as zeros (tensor-map bounds).
        const int set = warp >> 2, cw = warp & 3;
        const bool set_leader = cw == 0 && lane == 0;
        const int myrow = cw * 32 + lane;
        const uint32_t my_a = tmem_base + ((uint32_t)(cw * 32) << 16) + (uint32_t)(2 * g.n_tile);     // A stages of this warp's lanes
        const int sw = (myrow >> 1) & 3;

        int s = set % STAGES, use = set / STAGES, kb = set, slab = set % DEPTH, suse = set / DEPTH;
        TRACE_DECL();
        TRACE_T0();
#pragma unroll 1
        for (int it = set; it < total_it; it += SETS) {
            while (kb >= g.num_kb) kb -= g.num_kb;
            mbar_wait(&raw_full[slab], suse & 1);    // the raw k-block has landed (async proxy -> mbarrier -> visible)
            TRACE_ADD(0);
            mbar_wait(&empty[s], (use & 1) ^ 1);     // the MMAs that read this stage (weights in smem, A in TMEM) are complete
            tc_fence_after();
            TRACE_ADD(1);
            {
                // weight k-block: one bulk copy per converter warp of the set (4 concurrent requests of n_tile*32 bytes)
                uint8_t *b_all = stages + s * stage_bytes;
                const uint32_t bytes = (uint32_t)g.n_tile * 128u, part = bytes >> 2;
                if (set_leader) mbar_arrive_expect_tx(&full[s], bytes);
                if (lane == 0) bulk_g2s(b_all + cw * part, w_split + (size_t)kb * bytes + cw * part, part, &full[s]);
            }
            {
                const uint8_t *rs = raw + slab * P4_SLAB + myrow * 64;
                float4 x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4 *>(rs + ((j ^ sw) << 4));
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    split_tf32(x[j].x, hi[4 * j], lo[4 * j]); split_tf32(x[j].y, hi[4 * j + 1], lo[4 * j + 1]);
                    split_tf32(x[j].z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(x[j].w, hi[4 * j + 3], lo[4 * j + 3]);
                }
                const uint32_t ta = my_a + (uint32_t)(s * 32);
                tmem_st16_nowait(ta, hi);
                tmem_st16_nowait(ta + 16, lo);
                tmem_st_wait();
            }
            tc_fence_before();                       // the TMEM writes are ordered before the arrive the MMA thread waits for
            mbar_arrive(&full[s]);
            mbar_arrive(&raw_empty[slab]);           // this row of the raw k-block is in registers / TMEM: the slab may be refilled
            TRACE_ADD(2);
            s += SETS;
            while (s >= STAGES) { s -= STAGES; ++use; }
            kb += SETS;
            slab += SETS;
            while (slab >= DEPTH) { slab -= DEPTH; ++suse; }
        }
        if (tid == 0) TRACE_FLUSH(0, 4);
    } else if (warp == C::LOAD_WARP) {
        // =========================== raw A loader: one TMA box [128 rows x 16 columns] per k-block, DEPTH k-blocks ahead ===========
        if (lane == 0) {
            int slab = 0, suse = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * p.rows_per_tile;
                const int rows_here = (int)max((int64_t)0, min((int64_t)p.rows_per_tile, p.rows - row0));
                // the k-blocks visit every row of the tile in 64-byte pieces over a tile period: pull the tile's rows (contiguous in
                // memory) into L2 with four sequential bulk prefetches first (DRAM row-buffer locality; measured in round 1;
                // prefetching one tile AHEAD instead measured equal or worse: 0.142 / 0.171 / 0.159 / 0.185 vs 0.142 / 0.169 / 0.159 / 0.171 ms)
                for (int q = 0; q < 4; ++q) {
                    const int slab_rows = min(32, rows_here - q * 32);
                    if (slab_rows > 0) {
                        const uint32_t pf = (uint32_t)(((slab_rows - 1) * p.lda + p.K) * 4) & ~15u;
                        if (pf) l2_prefetch_4(p.a + (row0 + q * 32) * (int64_t)p.lda, pf);
                    }
                }
                for (int kb = 0; kb < g.num_kb; ++kb) {
                    mbar_wait(&raw_empty[slab], (suse & 1) ^ 1);     // the converters have read the k-block that was here
                    mbar_arrive_expect_tx(&raw_full[slab], P4_SLAB);
                    tma_load_2d(smem_u32(raw + slab * P4_SLAB), &tm_a, kb * 16, (int)row0, &raw_full[slab]);
                    if (++slab == DEPTH) { slab = 0; ++suse; }
                }
            }
        }
    } else if (warp == C::MMA_WARP) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_M, (uint32_t)g.n_tile, 0, 0);
            const uint32_t lbo_b = (uint32_t)g.n_tile * 16, sbo = 128;
            int s = 0, use = 0;
            TRACE_DECL();
            TRACE_T0();
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int bar = ti & 1, ause = ti >> 1;
                mbar_wait(&tmem_empty[bar], (ause & 1) ^ 1);        // the epilogue group has drained this accumulator
                tc_fence_after();
                TRACE_ADD(4);
                for (int kb = 0; kb < g.num_kb; ++kb) {
                    mbar_wait(&full[s], use & 1);
                    tc_fence_after();
                    TRACE_ADD(5);
                    const uint32_t b_hi = smem_u32(stages + s * stage_bytes), b_lo = b_hi + 4 * lbo_b;
                    const uint32_t a_hi = tmem_base + (uint32_t)(2 * g.n_tile + s * 32), a_lo = a_hi + 16;
                    const uint32_t d = tmem_base + (uint32_t)(bar * g.n_tile);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t dbh = make_smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, sbo);
                        const uint64_t dbl = make_smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, sbo);
                        umma_tf32_ts(d, a_hi + kk * 8, dbh, idesc, (kb | kk) ? 1u : 0u);
                        umma_tf32_ts(d, a_hi + kk * 8, dbl, idesc, 1u);
                        umma_tf32_ts(d, a_lo + kk * 8, dbh, idesc, 1u);
                    }
                    umma_commit(&empty[s]);            // stage reusable once these MMAs have read it
                    if (++s == STAGES) { s = 0; ++use; }
                    TRACE_ADD(6);
                }
                umma_commit(&tmem_full[bar]);          // accumulator complete -> epilogue
            }
            TRACE_FLUSH(4, 7);
        }
    } else {
        // =========================== epilogue groups (thread = row of the tile, TMEM lane = row) ===========================
        const int ew = warp - C::EPI0_WARP;                          // 0..7
        const int grp = ew >> 2;                                     // group 0: even tiles / accumulator 0, group 1: odd tiles
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read (EPI0_WARP % 4 == 0)
        const int et = quad * 32 + lane;                             // row of the tile
        const int r7 = lane & 7;                                     // swizzle key of this thread's row
        uint8_t *gtiles = tiles + (size_t)grp * SLOTS * 4 * P4_TILE; // [slot][quad][4 KB] of this group
        float *mystat = wstat + ew * STATW;
        uint64_t *mybar = aux_bar + ew * 3;
        const int n_chunks = (g.n_tile + 31) / 32;
        // box height of this warp's result stores: 32 rows, or the remainder of a tile whose height is not a multiple of 32
        // (tiles of whole centre points, RELU_MAXMIN): that quadrant stores through the second tensor map
        const int nominal = max(0, min(32, p.rows_per_tile - quad * 32));
        const CUtensorMap *my_out = nominal == 32 ? &tm_out : &tm_out_tail;
        uint32_t n = 0;                                              // chunks processed by this warp (slot = n % SLOTS)
        TRACE_DECL();
        TRACE_T0();
        for (int ti = grp; ti < my_tiles; ti += 2) {
            const int ause = ti >> 1;
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)ti * gridDim.x) * p.rows_per_tile;
            const int rows_here = (int)max((int64_t)0, min((int64_t)p.rows_per_tile, p.rows - row0));
            const bool valid = et < rows_here;
            const int wrow0 = (int)(row0 + quad * 32);               // first row of this warp (rows < 2^31: eligibility)
            const int wrows = max(0, min(32, rows_here - quad * 32));
            auto issue_aux = [&](int ch, uint32_t nn) {              // lane 0: aux box of chunk ch -> slot nn % SLOTS
                const uint32_t sl = nn % SLOTS;
                mbar_arrive_expect_tx(&mybar[sl], P4_TILE);
                tma_load_2d(smem_u32(gtiles + (sl * 4 + quad) * P4_TILE), &tm_aux, ch * 32, wrow0, &mybar[sl]);
            };
            if (BWD) {
                if (lane == 0 && wrows > 0) {
                    const uint32_t pf = (uint32_t)(((wrows - 1) * p.ldaux + p.n_out) * 4) & ~15u;
                    if (pf) l2_prefetch_4(p.aux + (int64_t)wrow0 * p.ldaux, pf);
                    issue_aux(0, n);                                 // slots n % 3 and (n + 1) % 3 are free: the stores of chunks
                    if (n_chunks > 1) issue_aux(1, n + 1);           // <= n - 2 were waited for in the previous chunk loop
                }
            }
            TRACE_ADD(9 + grp * 3);                                  // tile prologue
            mbar_wait(&tmem_full[grp], ause & 1);
            tc_fence_after();
            TRACE_ADD(7 + grp * 3);                                  // waiting for the accumulator
            for (int ch = 0; ch < n_chunks; ++ch, ++n) {
                const int c0 = ch * 32;
                const int nv = min(32, p.n_out - c0);                // valid columns of this chunk (>= 1)
                const uint32_t sl = n % SLOTS;
                uint8_t *gslot = gtiles + sl * 4 * P4_TILE;          // the group's four tiles of this slot (rows 0..127)
                uint8_t *myrow = gslot + quad * P4_TILE + lane * 128;
                if (BWD) {
                    if (wrows > 0) mbar_wait(&mybar[sl], (n / SLOTS) & 1);     // aux box landed (async proxy -> mbarrier -> visible)
                } else {
                    if (lane == 0) bulk_wait_read<1>();              // the store that last read this slot (chunk n - 2) is done
                    __syncwarp();
                }
                float acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * g.n_tile + c0), acc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 *cell = reinterpret_cast<float4 *>(myrow + ((i ^ r7) << 4));      // logical chunk i of this row
                    if (BWD) {
                        const float4 a4 = *cell;
                        const float4 k04 = *reinterpret_cast<const float4 *>(colv + 256 + c0 + 4 * i);
                        const float4 k14 = *reinterpret_cast<const float4 *>(colv + 512 + c0 + 4 * i);
                        const float4 mu4 = *reinterpret_cast<const float4 *>(colv + 768 + c0 + 4 * i);
                        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                        const float k0v[4] = {k04.x, k04.y, k04.z, k04.w}, k1v[4] = {k14.x, k14.y, k14.z, k14.w};
                        const float muv[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float a = av[e];
                            acc[4 * i + e] = (valid && a > 0.f) ? (acc[4 * i + e] - k0v[e] - (a - muv[e]) * k1v[e]) : 0.f;
                        }
                    } else {
                        const float4 b4 = *reinterpret_cast<const float4 *>(colv + c0 + 4 * i);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                        if (EPI == NT_EPI_BIAS) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[4 * i + e] += bb[e];
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[4 * i + e] = valid ? fmaxf(acc[4 * i + e] + bb[e], 0.f) : 0.f;
                        }
                    }
                    *cell = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                }
                fence_proxy_async();                                 // the tile is read by the TMA store (async proxy)
                if (EPI == NT_EPI_RELU_MAXMIN) asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                else __syncwarp();
                if (lane == 0) {
                    if (p.out && nominal > 0 && wrows > 0) tma_store_2d(my_out, smem_u32(gslot + quad * P4_TILE), c0, wrow0);
                    bulk_commit();                                   // one group per chunk (possibly empty)
                    if (BWD) {
                        bulk_wait_read<1>();                         // store of chunk n - 1 has read its slot = (n + 2) % 3
                        if (ch + 2 < n_chunks && wrows > 0) issue_aux(ch + 2, n + 2);
                    }
                }
                // element (row r, column c) of a 4 KB tile: float index r*32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3)
                if (EPI == NT_EPI_RELU_MAXMIN) {
                    // max / min over the k edge rows of every centre point (nodes straddle warps: the four tiles of the group are
                    // contiguous, row R of the tile at R*128 bytes) and the column statistics in the same pass
                    const float *gt = reinterpret_cast<const float *>(gslot);
                    const int kk = p.k_agg;
                    const int nodes_here = rows_here / kk;
                    const int64_t node0 = row0 / kk;
                    const int cq = lane >> 2, cr = lane & 3;
                    float t1 = 0.f, t2 = 0.f;
                    if (lane < nv) {
                        if (kk == 5) maxmin_scan<5>(gt, kk, et, cq, cr, nodes_here, node0, p, c0 + lane, t1, t2);
                        else maxmin_scan<0>(gt, kk, et, cq, cr, nodes_here, node0, p, c0 + lane, t1, t2);
                        if (p.stats) { mystat[c0 + lane] += t1; mystat[256 + c0 + lane] += t2; }
                    }
                    // no trailing barrier: the next chunk writes the OTHER slot, and a thread reaches the barrier of chunk n + 1 only
                    // after this scan, so the writes of chunk n + 2 (this slot again) are ordered behind it
                } else {
                    const bool want = (EPI == NT_EPI_BIAS) ? false : (BWD ? (p.colsum != nullptr) : (p.stats != nullptr));
                    if (want && lane < nv) {
                        // column sums of this warp's 32 rows (lane = column; invalid rows hold zeros)
                        const float *t4 = reinterpret_cast<const float *>(gslot + quad * P4_TILE);
                        const int cq = lane >> 2, cr = lane & 3;
                        float t1 = 0.f, t2 = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) {
                            const float x = t4[rr * 32 + ((cq ^ (rr & 7)) << 2) + cr];
                            t1 += x;
                            if (!BWD) t2 = fmaf(x, x, t2);
                        }
                        mystat[c0 + lane] += t1;
                        if (!BWD) mystat[256 + c0 + lane] += t2;
                    }
                    if (BWD) __syncwarp();                           // the column reads are done before lane 0 may refill the slot
                }
            }
            // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[grp]);
            TRACE_ADD(8 + grp * 3);                                  // draining
        }
        if (lane == 0) bulk_wait_all();                              // result boxes are in global memory before the CTA exits
        if (et == 0) TRACE_FLUSH(7 + grp * 3, 10 + grp * 3);
        // flush the per-CTA column statistics once (sum over the eight warps' private accumulators)
        if (EPI != NT_EPI_BIAS) {
            asm volatile("bar.sync 3, 256;" ::: "memory");
            for (int c = ew * 32 + lane; c < p.n_out; c += 256) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) {
                    s1 += wstat[w8 * STATW + c];
                    if (!BWD) s2 += wstat[w8 * STATW + 256 + c];
                }
                if (BWD) {
                    if (p.colsum) atomicAdd(p.colsum + c, (double)s1);
                } else if (p.stats) {
                    atomicAdd(p.stats + c, (double)s1);
                    atomicAdd(p.stats + p.n_out + c, (double)s2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::MMA_WARP) tmem_dealloc(tmem_base, 512);
}

// ---- host side: tensor maps --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// fp32 matrix [rows, cols] with row stride ld (floats), boxes of box_rows x box_cols (32: 128-byte swizzle, 16: 64-byte swizzle)
static bool make_map(CUtensorMap *m, const float *base, int64_t rows, int cols, int ld, int box_rows, int box_cols = 32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc4_eligible(const NTParams &p, int producer, int epilogue) {
    if (!tc3_eligible(p, producer, epilogue) || p.scatter) return false;
    const bool bwd = epilogue == NT_EPI_BNRELU_BWD;
    if (bwd && !p.out) return false;
    if (p.rows >= ((int64_t)1 << 31) - 256) return false;
    const TCGeom g = tc_geometry(p.n_out, p.K, NT_PREC_TF32X3);
    return g.n_tile <= P4_MAX_NTILE && tc4_smem_bytes(g.n_tile, bwd) <= 227 * 1024;
}

template <int EPI, int SETS>
static int launch_tc4_t(const NTParams &p, const void *w_split, const TCGeom &g, int sms, cudaStream_t st) {
    constexpr bool BWD = EPI == NT_EPI_BNRELU_BWD;
    const size_t smem = tc4_smem_bytes(g.n_tile, BWD);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_nt_tc4_kernel<EPI, SETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail("nt_gemm_nt(tc4): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    // tensor maps: the result (32-row boxes, and the box of a tile's last quadrant when tiles are not a multiple of 32 rows high)
    // and the aux operand of BNRELU_BWD.  Unused maps alias a valid one (the kernel never dereferences them).
    CUtensorMap tm_a, tm_out, tm_tail, tm_aux;
    if (!make_map(&tm_a, p.a, p.rows, p.K, p.lda, TC_M, 16)) return fail("nt_gemm_nt(tc4): cuTensorMapEncodeTiled failed%s", "");
    const float *some = p.out ? p.out : p.a;
    const int some_cols = p.out ? p.n_out : p.K, some_ld = p.out ? p.ldo : p.lda;
    if (!make_map(&tm_out, some, p.rows, some_cols, some_ld, 32)) return fail("nt_gemm_nt(tc4): cuTensorMapEncodeTiled failed%s", "");
    const int tail = p.rows_per_tile % 32;
    tm_tail = tm_out;
    if (tail && !make_map(&tm_tail, some, p.rows, some_cols, some_ld, tail)) return fail("nt_gemm_nt(tc4): cuTensorMapEncodeTiled failed%s", "");
    tm_aux = tm_out;
    if (BWD && !make_map(&tm_aux, p.aux, p.rows, p.n_out, p.ldaux, 32)) return fail("nt_gemm_nt(tc4): cuTensorMapEncodeTiled failed%s", "");
    const int64_t n_row_tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
    const int ctas = (int)(n_row_tiles < sms ? n_row_tiles : sms);
    gemm_nt_tc4_kernel<EPI, SETS><<<ctas, P4Cfg<EPI, SETS>::THREADS, smem, st>>>(tm_a, tm_out, tm_tail, tm_aux, p, reinterpret_cast<const uint8_t *>(w_split), g);
    return check_launch("nt_gemm_nt(tc4)");
}

// -1: not eligible (the caller tries the first-generation streaming engine, then the one-tile-per-CTA engine)
int launch_nt_tc4(const NTParams &p, int producer, int epilogue, const void *w_split, cudaStream_t st) {
    if (encode_fn() == nullptr) return -1;
    const TCGeom g = tc_geometry(p.n_out, p.K, NT_PREC_TF32X3);
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1)
            return fail("nt_gemm_nt(tc4): cannot query the SM count%s", "");
        sms = n;
    }
    if (g.n_tiles > 1) {
        // NT_EPI_BIAS with several column tiles (the split first EdgeConv Linear: n_out = 2 H1 = 400): one launch per column tile of the
        // prepared weights -- the rows are read once per tile (K is small for these calls), every tile streams
        if (epilogue != NT_EPI_BIAS || g.n_tile > P4_MAX_NTILE || g.n_tiles > 4 || tc4_smem_bytes(g.n_tile, false) > 227 * 1024) return -1;
        TCGeom g1 = g;
        g1.n_tiles = 1;
        NTParams q[4];
        for (int t = 0; t < g.n_tiles; ++t) {
            q[t] = p;
            q[t].n_out = p.n_out - t * g.n_tile < g.n_tile ? p.n_out - t * g.n_tile : g.n_tile;
            q[t].out = p.out + t * g.n_tile;
            q[t].bias = p.bias ? p.bias + t * g.n_tile : nullptr;
            if (!tc4_eligible(q[t], producer, epilogue)) return -1;
        }
        for (int t = 0; t < g.n_tiles; ++t) {
            const uint8_t *ws = reinterpret_cast<const uint8_t *>(w_split) + (size_t)t * g.num_kb * g.n_tile * 128;
            const int rc = launch_tc4_t<NT_EPI_BIAS, 2>(q[t], ws, g1, sms, st);
            if (rc) return rc;
        }
        return 0;
    }
    if (!tc4_eligible(p, producer, epilogue)) return -1;
    switch (epilogue) {
        case NT_EPI_BIAS: return launch_tc4_t<NT_EPI_BIAS, 2>(p, w_split, g, sms, st);
        case NT_EPI_RELU_STATS: return launch_tc4_t<NT_EPI_RELU_STATS, 2>(p, w_split, g, sms, st);
        case NT_EPI_RELU_MAXMIN: return launch_tc4_t<NT_EPI_RELU_MAXMIN, 2>(p, w_split, g, sms, st);
        default: return launch_tc4_t<NT_EPI_BNRELU_BWD, 2>(p, w_split, g, sms, st);      // (one converter set: 0.212 vs 0.206 ms)
    }
}

}  // namespace nt

#ifdef NT_TC3_TRACE
extern "C" int nt_debug_tc4_trace(unsigned long long *host_out) {      // [256][16]
    return cudaMemcpyFromSymbol(host_out, nt::g_tc4_trace, sizeof(nt::g_tc4_trace)) == cudaSuccess ? 0 : 1;
}
#endif
