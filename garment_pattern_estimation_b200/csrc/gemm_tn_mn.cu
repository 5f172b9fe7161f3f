// Weight-gradient GEMM without transposing producers (sm_100a, tcgen05 + TMEM):   out[m, n] (+)= sum_r A[r, m] * (B[r, n] - mu[n])
//
// The contraction runs over ROWS, and a row of A / B is contiguous along m / n: for the tensor core that is an MN-major operand.
// gemm_tn_tc.cu transposes in its producers (four scalar loads from four rows -> one K-major 16-byte chunk, TF32x3) because
// kind::tf32 rejected MN-major no-swizzle descriptors; kind::f16 accepts them (tools/microbench/mn_major_test.cu, round 2: LBO = 128 B
// between 8-row groups along K, SBO = distance between 8-element groups along M/N).  So the rows are consumed as they lie:
//   loader warp   : one cp.async.bulk per operand row (16 rows per stage) into a raw fp32 ring that takes all the shared memory left
//                   over -- ~128 KB in flight per SM, which is what hides the HBM latency (a register-staged first version of this
//                   kernel had one 51 KB stage in flight and ran at 3.7 TB/s);
//   8 converters  : thread = (column quad, row subset): LDS.128 -> (centre) -> bf16 hi / lo split (x = hi + lo + O(2^-17 |x|)) ->
//                   two STS.64.  A lane pair writes one 16-byte core-matrix row; the layout is [8-column group][k-row][16 B] with the
//                   group stride padded to 17 rows (conflict-free);
//   MMA issuer    : three MMAs per 16-row K-step and accumulator (hi.hi + hi.lo + lo.hi, ~1e-5 relative) at the bf16 rate.
// Each CTA owns a slice of rows and a full [<=256 x <=256] fp32 partial product in TMEM; partials are reduced in double by a second
// kernel (deterministic, no atomics) -- the same contract as gemm_tn_tc.cu, which stays for gathered operands.
#include <cstdio>
#include <cstdlib>
#include "gemm_params.cuh"
#include "tc_common.cuh"

namespace nt {
using namespace tc;

constexpr int MN_CONVERTER_WARPS = 16;                       // 4 per scheduler: the convert chain is latency-bound with 2
constexpr int MN_RSUBS = MN_CONVERTER_WARPS / 2;            // 64 column quads x MN_RSUBS row subsets
constexpr int MN_THREADS = (MN_CONVERTER_WARPS + 2) * 32;    // then the loader warp (bulk copies) and the TMEM + MMA issuer warp
constexpr int MN_RB = 16;                                    // rows (K) per stage = one bf16 K-step
constexpr int MN_RPT = MN_RB / MN_RSUBS;                     // rows per converter thread and stage
constexpr int MN_OP_STAGES = 3;                              // split-operand ring
constexpr int MN_MAX_RAW_STAGES = 8;                         // raw fp32 ring: as many stages as fit (bytes in flight hide the HBM latency)
constexpr int MN_GROUP_STRIDE = MN_RB * 16 + 16;             // 272 B: 16 rows x 16 B + one row of padding (conflict-free stores)
// How the loader warp brings 16 rows of an operand into the raw ring.  Measured (B200, round 2): one bulk copy per ROW (<= 1 KB) runs at
// ~0.2 us per copy and 4-byte cp.async from a single warp at ~50 ns per instruction -- both an order of magnitude too slow -- so
// the 16 rows are always fetched WHOLE (all ld columns, any ld: 16 rows x ld floats are one contiguous, 64-byte aligned piece of memory)
// with ONE bulk copy, and the converters pick the tile's columns out of the raw rows.
constexpr int MN_LOAD_ELEMENTS = 0;                          // fallback (base pointer not 16-byte aligned / rows too long for the ring)
constexpr int MN_LOAD_BLOCK = 1;
constexpr int MN_HEAD_BYTES = 256;                           // barriers + TMEM slot
constexpr int MN_EPI_BYTES = MN_CONVERTER_WARPS * 32 * 33 * 4;                // epilogue transposition tiles (re-use the rings)

struct MNParams {
    const float *a; int lda; int m;
    const float *b; int ldb; int n;
    const float *mu;
    int64_t rows, rows_per_split;
    int m_pad, n_pad;                    // m_pad in {128, 256}; n_pad multiple of 16, <= 256
    int m0, n0;
    int m_cnt, n_cnt;                    // valid columns of this tile (<= 256 each)
    int a_mode, b_mode;                  // MN_LOAD_*: how the loader warp brings a 16-row block of the operand into the raw ring
    int a_rb, b_rb;                      // bytes between rows of the operand's raw block
    int a_col0, b_col0;                  // byte offset of the tile's first column inside a raw row
    int a_vec, b_vec;                    // 1: a thread's 4 columns are 16-byte aligned in the raw block (LDS.128)
    int raw_stages;
    float *partial;                      // [splits][m_pad][n_pad]
};

__device__ __forceinline__ void cp_async4(uint32_t dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {             // counts against the barrier's initial count (.noinc)
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(MN_THREADS, 1) gemm_tn_mn_kernel(MNParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int ga = p.m_pad / 8, gb = p.n_pad / 8;                         // 8-column groups per operand
    const size_t a_plane = (size_t)ga * MN_GROUP_STRIDE, b_plane = (size_t)gb * MN_GROUP_STRIDE;
    const size_t op_stage = 2 * a_plane + 2 * b_plane;
    const int a_quads = (p.m_cnt + 3) / 4, b_quads = (p.n_cnt + 3) / 4;
    const size_t a_rb = (size_t)p.a_rb, b_rb = (size_t)p.b_rb, raw_b_off = a_rb * MN_RB, raw_stage = (a_rb + b_rb) * MN_RB;
    const bool any_elements = p.a_mode == MN_LOAD_ELEMENTS || p.b_mode == MN_LOAD_ELEMENTS;
    uint8_t *ops = smem + MN_HEAD_BYTES;                                   // [barriers | operand ring | raw ring]
    uint8_t *raw = ops + MN_OP_STAGES * op_stage;
    uint64_t *raw_full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *raw_empty = raw_full + MN_MAX_RAW_STAGES;
    uint64_t *op_full = raw_empty + MN_MAX_RAW_STAGES;
    uint64_t *op_empty = op_full + MN_OP_STAGES;
    uint64_t *tmem_full = op_empty + MN_OP_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_tiles = p.m_pad / 128;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < m_tiles * p.n_pad) tmem_cols <<= 1;

    if (warp == MN_CONVERTER_WARPS && lane == 0) {
        for (int s = 0; s < p.raw_stages; ++s) { mbar_init(&raw_full[s], any_elements ? 33 : 1); mbar_init(&raw_empty[s], MN_CONVERTER_WARPS); }
        for (int s = 0; s < MN_OP_STAGES; ++s) { mbar_init(&op_full[s], MN_CONVERTER_WARPS); mbar_init(&op_empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == MN_CONVERTER_WARPS + 1) tmem_alloc(tmem_slot, tmem_cols);
    // padded columns of the operand ring are never written again: zero them (and everything else) once
    for (size_t i = (size_t)tid * 16; i < MN_OP_STAGES * op_stage; i += (size_t)MN_THREADS * 16)
        *reinterpret_cast<uint4 *>(ops + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_split;
    const int64_t r_end = min(p.rows, r_begin + p.rows_per_split);
    const int n_stages = (int)((max((int64_t)0, r_end - r_begin) + MN_RB - 1) / MN_RB);

    if (warp < MN_CONVERTER_WARPS) {
        // ===== converters: raw fp32 rows -> bf16 hi / lo MN-major core matrices.  thread = (column quad, row subset) =====
        const int quad = tid & 63, rsub = tid >> 6;                       // rows rsub + MN_RSUBS * i
        const bool a_on = quad < a_quads, b_on = quad < b_quads;
        const int a_left = p.m_cnt - 4 * quad, b_left = p.n_cnt - 4 * quad;
        const bool a_tail = a_on && a_left < 4, b_tail = b_on && b_left < 4;      // partial last quad: its padding may hold anything
        float4 muv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.mu && b_on) {
            const float *mu = p.mu + p.n0 + 4 * quad;
            muv.x = __ldg(mu);
            if (b_left > 1) muv.y = __ldg(mu + 1);
            if (b_left > 2) muv.z = __ldg(mu + 2);
            if (b_left > 3) muv.w = __ldg(mu + 3);
        }
        const uint32_t src_a = (uint32_t)(rsub * a_rb + p.a_col0 + quad * 16), src_b = (uint32_t)(raw_b_off + rsub * b_rb + p.b_col0 + quad * 16);
        auto load4 = [](const uint8_t *q, bool vec) {
            if (vec) return *reinterpret_cast<const float4 *>(q);
            const float *f = reinterpret_cast<const float *>(q);
            return make_float4(f[0], f[1], f[2], f[3]);
        };
        const uint32_t dst_q = (uint32_t)((quad >> 1) * MN_GROUP_STRIDE + (quad & 1) * 8 + rsub * 16);   // [8-column group][k-row][16 B]
        auto split_store = [&](uint8_t *hi_plane, uint8_t *lo_plane, uint32_t off, float4 v) {
            uint32_t h01, l01, h23, l23;
            split_bf16x2(v.x, v.y, h01, l01);
            split_bf16x2(v.z, v.w, h23, l23);
            *reinterpret_cast<uint2 *>(hi_plane + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2 *>(lo_plane + off) = make_uint2(l01, l23);
        };
        const bool all_vec = p.a_vec && p.b_vec;
        int rs = 0, os = 0;
        uint32_t raw_phase = 0, op_phase = 1;
        int rows_left = (int)(r_end - r_begin);
        for (int st = 0; st < n_stages; ++st, rows_left -= MN_RB) {
            mbar_wait(&raw_full[rs], raw_phase);
            mbar_wait(&op_empty[os], op_phase);
            const uint8_t *src = raw + rs * raw_stage;
            uint8_t *a_hi = ops + os * op_stage, *a_lo = a_hi + a_plane, *b_hi = a_lo + a_plane, *b_lo = b_hi + b_plane;
            float4 va[MN_RPT], vb[MN_RPT];
            if (all_vec) {                                                // the common case, kept free of per-load branches
#pragma unroll
                for (int i = 0; i < MN_RPT; ++i) {
                    va[i] = a_on ? *reinterpret_cast<const float4 *>(src + src_a + MN_RSUBS * i * a_rb) : make_float4(0.f, 0.f, 0.f, 0.f);
                    vb[i] = b_on ? *reinterpret_cast<const float4 *>(src + src_b + MN_RSUBS * i * b_rb) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int i = 0; i < MN_RPT; ++i) {
                    va[i] = a_on ? load4(src + src_a + MN_RSUBS * i * a_rb, p.a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                    vb[i] = b_on ? load4(src + src_b + MN_RSUBS * i * b_rb, p.b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int i = 0; i < MN_RPT; ++i) { vb[i].x -= muv.x; vb[i].y -= muv.y; vb[i].z -= muv.z; vb[i].w -= muv.w; }
            if (a_tail || b_tail || rows_left < MN_RB) {                 // rare: ragged columns / the last rows of the slice
#pragma unroll
                for (int i = 0; i < MN_RPT; ++i) {
                    const bool row_ok = rsub + MN_RSUBS * i < rows_left;
                    if (!row_ok || a_left < 2) va[i].y = 0.f;
                    if (!row_ok || a_left < 3) va[i].z = 0.f;
                    if (!row_ok || a_left < 4) va[i].w = 0.f;
                    if (!row_ok) va[i].x = 0.f;
                    if (!row_ok || b_left < 2) vb[i].y = 0.f;
                    if (!row_ok || b_left < 3) vb[i].z = 0.f;
                    if (!row_ok || b_left < 4) vb[i].w = 0.f;
                    if (!row_ok) vb[i].x = 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < MN_RPT; ++i) {
                if (a_on) split_store(a_hi, a_lo, dst_q + 16 * MN_RSUBS * i, va[i]);
                if (b_on) split_store(b_hi, b_lo, dst_q + 16 * MN_RSUBS * i, vb[i]);
            }
            // one arrival per warp: mbarrier arrivals on one word serialise per thread
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&op_full[os]);
                mbar_arrive(&raw_empty[rs]);
            }
            if (++rs == p.raw_stages) { rs = 0; raw_phase ^= 1; }
            if (++os == MN_OP_STAGES) { os = 0; op_phase ^= 1; }
        }

        // =========================== epilogue: TMEM -> coalesced partial tile ===========================
        const int part = warp >> 2, quadrant = warp & 3;                  // a warp reads the TMEM lanes of its quadrant (warp % 4)
        float *tw = reinterpret_cast<float *>(ops) + warp * (32 * 33);           // ring memory is free once tmem_full fired
        float *dst_tile = p.partial + (size_t)blockIdx.x * p.m_pad * p.n_pad;
        if (n_stages > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        for (int mt = 0; mt < m_tiles; ++mt) {
            for (int c0 = part * 32; c0 < p.n_pad; c0 += 32 * (MN_CONVERTER_WARPS / 4)) {      // the warps of a quadrant interleave the column chunks
                float acc[32];
                if (n_stages > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(quadrant * 32) << 16) + (uint32_t)(mt * p.n_pad + c0), acc);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) tw[lane * 33 + i] = acc[i];
                __syncwarp();
                if (c0 + lane < p.n_pad) {
                    float *dst = dst_tile + (size_t)(mt * 128 + quadrant * 32) * p.n_pad + c0 + lane;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) dst[(size_t)rr * p.n_pad] = tw[rr * 33 + lane];
                }
                __syncwarp();
            }
        }
    } else if (warp == MN_CONVERTER_WARPS) {
        // ===== loader: whole raw fp32 rows -> raw ring (asynchronously: the ring depth is what hides the HBM latency) =====
        const float *ga_base = p.a + p.m0, *gb_base = p.b + p.n0;                 // ELEMENTS mode copies the tile's columns only
        // bytes of the LAST row of a block that are copied: up to the tile's last column (what lies behind the last row of a tensor may
        // not exist), rounded up to the 16 bytes a bulk copy moves when that stays inside the row stride
        auto last_row_bytes = [](int col0, int cnt, int rb) {
            const uint32_t end = (uint32_t)(col0 + cnt * 4), up = (end + 15u) & ~15u;
            return up <= (uint32_t)rb ? up : end;
        };
        const uint32_t a_last = last_row_bytes(p.a_col0, p.m_cnt, p.a_rb), b_last = last_row_bytes(p.b_col0, p.n_cnt, p.b_rb);
        int rs = 0;
        uint32_t phase = 1;
        for (int st = 0; st < n_stages; ++st) {
            mbar_wait(&raw_empty[rs], phase);
            uint8_t *dst = raw + rs * raw_stage;
            uint64_t *bar = &raw_full[rs];
            const int64_t r0 = r_begin + (int64_t)st * MN_RB;
            const int valid = (int)min((int64_t)MN_RB, r_end - r0);
            if (lane == 0) {
                // a bulk copy moves multiples of 16 bytes: up to 3 floats left over (unpadded rows only) go through registers
                const uint32_t a_bytes = p.a_mode == MN_LOAD_BLOCK ? (uint32_t)(valid - 1) * p.a_rb + a_last : 0u;
                const uint32_t b_bytes = p.b_mode == MN_LOAD_BLOCK ? (uint32_t)(valid - 1) * p.b_rb + b_last : 0u;
                const float *a_src = p.a + r0 * p.lda, *b_src = p.b + r0 * p.ldb;
                for (uint32_t o = a_bytes & ~15u; o < a_bytes; o += 4) *reinterpret_cast<float *>(dst + o) = __ldg(a_src + o / 4);
                for (uint32_t o = b_bytes & ~15u; o < b_bytes; o += 4) *reinterpret_cast<float *>(dst + raw_b_off + o) = __ldg(b_src + o / 4);
                const uint32_t tx = (a_bytes & ~15u) + (b_bytes & ~15u);
                if (tx) mbar_arrive_expect_tx(bar, tx);
                else mbar_arrive(bar);
                if (a_bytes & ~15u) bulk_g2s(dst, a_src, a_bytes & ~15u, bar);
                if (b_bytes & ~15u) bulk_g2s(dst + raw_b_off, b_src, b_bytes & ~15u, bar);
            }
            if (any_elements) {
                if (p.a_mode == MN_LOAD_ELEMENTS)
                    for (int r = 0; r < valid; ++r)
                        for (int c = lane; c < p.m_cnt; c += 32) cp_async4(smem_u32(dst + r * a_rb + c * 4), ga_base + (r0 + r) * p.lda + c);
                if (p.b_mode == MN_LOAD_ELEMENTS)
                    for (int r = 0; r < valid; ++r)
                        for (int c = lane; c < p.n_cnt; c += 32)
                            cp_async4(smem_u32(dst + raw_b_off + r * b_rb + c * 4), gb_base + (r0 + r) * p.ldb + c);
                cp_async_arrive(bar);                                     // fires when this lane's copies have landed
            }
            if (++rs == p.raw_stages) { rs = 0; phase ^= 1; }
        }
    } else {
        // =========================== MMA issuer: MN-major A and B, one 16-row K-step per stage ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.n_pad, 1, 1);
            for (int st = 0; st < n_stages; ++st) {
                const int os = st % MN_OP_STAGES;
                mbar_wait(&op_full[os], (st / MN_OP_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(ops + os * op_stage), a_lo = a_hi + (uint32_t)a_plane;
                const uint32_t b_hi = a_lo + (uint32_t)a_plane, b_lo = b_hi + (uint32_t)b_plane;
                const uint64_t dbh = make_smem_desc(b_hi, 128, MN_GROUP_STRIDE), dbl = make_smem_desc(b_lo, 128, MN_GROUP_STRIDE);
                for (int mt = 0; mt < m_tiles; ++mt) {
                    const uint32_t moff = (uint32_t)mt * 16u * MN_GROUP_STRIDE;          // 16 groups = 128 output rows further
                    const uint64_t dah = make_smem_desc(a_hi + moff, 128, MN_GROUP_STRIDE);
                    const uint64_t dal = make_smem_desc(a_lo + moff, 128, MN_GROUP_STRIDE);
                    const uint32_t d = tmem_base + (uint32_t)(mt * p.n_pad);
                    umma_bf16(d, dah, dbh, idesc, st ? 1u : 0u);
                    umma_bf16(d, dah, dbl, idesc, 1u);
                    umma_bf16(d, dal, dbh, idesc, 1u);
                }
                umma_commit(&op_empty[os]);
            }
            if (n_stages > 0) umma_commit(tmem_full);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MN_CONVERTER_WARPS + 1) tmem_dealloc(tmem_base, tmem_cols);
}

// out[m, n] += sum_s partial[s, m - m0, n - n0]     (double accumulation, fixed order: deterministic; OutT = float or double)
// block = 8 warps x 32 consecutive elements: warp j sums the splits j, j + 8, ... (coalesced 128-byte reads, 8 independent chains per
// element instead of one 148-long chain -- the one-thread-per-element version took 14 us per call, 0.2 ms per C2 step), then the eight
// partial sums are added in warp order through shared memory
template <typename OutT>
__global__ void __launch_bounds__(256) mn_reduce_kernel(const float *__restrict__ partial, int splits, int m_pad, int n_pad, int m0, int n0,
                                                        int m, int n, OutT *__restrict__ out, int ldo) {
    __shared__ double part[8][32];
    const int j = threadIdx.x >> 5, e = threadIdx.x & 31;
    const int i = blockIdx.x * 32 + e;
    const int total = m_pad * n_pad;
    double acc = 0.0;
    if (i < total) {
        const float *src = partial + i;
        int s = j;
        for (; s + 8 < splits; s += 16) {                      // two loads in flight per iteration
            const float v0 = src[(size_t)s * total], v1 = src[(size_t)(s + 8) * total];
            acc += (double)v0;
            acc += (double)v1;
        }
        for (; s < splits; s += 8) acc += (double)src[(size_t)s * total];
    }
    part[j][e] = acc;
    __syncthreads();
    if (j == 0 && i < total) {
        const int mm = i / n_pad, nn = i - mm * n_pad;
        if (m0 + mm < m && n0 + nn < n) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += part[w][e];
            OutT *dst = out + (size_t)(m0 + mm) * ldo + n0 + nn;
            *dst = (OutT)((double)*dst + t);
        }
    }
}

int gemm_tn_mn(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows, const float *mu, void *out, int out_double,
               int ldo, float *workspace, cudaStream_t st) {
    int64_t splits = (rows + 8 * MN_RB - 1) / (8 * MN_RB);              // at least 8 stages per CTA
    if (splits > 148) splits = 148;
    if (splits < 1) splits = 1;
    int64_t rps = (rows + splits - 1) / splits;
    rps = ((rps + MN_RB - 1) / MN_RB) * MN_RB;
    splits = (rows + rps - 1) / rps;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tn_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail("nt_gemm_tn(mn): cudaFuncSetAttribute failed%s", "");
        configured = true;
    }
    for (int m0 = 0; m0 < m; m0 += 256) {
        for (int n0 = 0; n0 < n; n0 += 256) {
            MNParams p{};
            p.a = a; p.lda = lda; p.m = m; p.b = b; p.ldb = ldb; p.n = n; p.mu = mu;
            p.rows = rows; p.rows_per_split = rps; p.m0 = m0; p.n0 = n0;
            p.m_cnt = min(m - m0, 256); p.n_cnt = min(n - n0, 256);
            p.m_pad = p.m_cnt > 128 ? 256 : 128;
            p.n_pad = ((p.n_cnt + 15) / 16) * 16;
            p.partial = workspace;
            const size_t op_stage = 2 * (size_t)(p.m_pad / 8 + p.n_pad / 8) * MN_GROUP_STRIDE;
            const size_t fixed = MN_OP_STAGES * op_stage + MN_HEAD_BYTES;
            const size_t budget = 227 * 1024 - 16;                  // 16 bytes of slack: a partial last quad is read whole
            auto elements = [](int cnt, int &mode, int &rb, int &col0, int &vec) {
                mode = MN_LOAD_ELEMENTS; rb = (cnt + 3) / 4 * 16; col0 = 0; vec = 1;
            };
            auto block = [](int ld, int first, int &mode, int &rb, int &col0, int &vec) {
                mode = MN_LOAD_BLOCK; rb = ld * 4; col0 = first * 4; vec = (ld % 4 == 0) && (first % 4 == 0);
            };
            if (aligned16(a)) block(lda, m0, p.a_mode, p.a_rb, p.a_col0, p.a_vec); else elements(p.m_cnt, p.a_mode, p.a_rb, p.a_col0, p.a_vec);
            if (aligned16(b)) block(ldb, n0, p.b_mode, p.b_rb, p.b_col0, p.b_vec); else elements(p.n_cnt, p.b_mode, p.b_rb, p.b_col0, p.b_vec);
            // rows too long for two ring stages (row strides beyond ~1000 floats): the wider operand falls back to its tile's columns
            if (fixed + 2 * (size_t)(p.a_rb + p.b_rb) * MN_RB > budget) {
                if (p.a_rb >= p.b_rb) elements(p.m_cnt, p.a_mode, p.a_rb, p.a_col0, p.a_vec);
                else elements(p.n_cnt, p.b_mode, p.b_rb, p.b_col0, p.b_vec);
            }
            if (fixed + 2 * (size_t)(p.a_rb + p.b_rb) * MN_RB > budget) {
                elements(p.m_cnt, p.a_mode, p.a_rb, p.a_col0, p.a_vec);
                elements(p.n_cnt, p.b_mode, p.b_rb, p.b_col0, p.b_vec);
            }
            const size_t raw_stage = (size_t)(p.a_rb + p.b_rb) * MN_RB;
            int raw_stages = (int)((budget - fixed) / raw_stage);
            if (raw_stages > MN_MAX_RAW_STAGES) raw_stages = MN_MAX_RAW_STAGES;
            if (raw_stages < 2) return fail("nt_gemm_tn(mn): tile does not fit shared memory%s", "");
            p.raw_stages = raw_stages;
            size_t smem = fixed + raw_stages * raw_stage + 16;
            if (smem < MN_HEAD_BYTES + MN_EPI_BYTES) smem = MN_HEAD_BYTES + MN_EPI_BYTES;
            static const bool debug = getenv("NT_TN_DEBUG") != nullptr;
            if (debug)
                fprintf(stderr, "gemm_tn_mn rows %ld m %d (ld %d) n %d (ld %d) tile (%d,%d) cnt (%d,%d) modes (%d,%d) splits %ld raw_stages %d mu %d\n",
                        (long)rows, m, lda, n, ldb, m0, n0, p.m_cnt, p.n_cnt, p.a_mode, p.b_mode, (long)splits, raw_stages, mu != nullptr);
            gemm_tn_mn_kernel<<<(unsigned)splits, MN_THREADS, smem, st>>>(p);
            int rc = check_launch("nt_gemm_tn(mn)");
            if (rc) return rc;
            const int total = p.m_pad * p.n_pad;
            if (out_double)
                mn_reduce_kernel<double><<<(total + 31) / 32, 256, 0, st>>>(workspace, (int)splits, p.m_pad, p.n_pad, m0, n0, m, n,
                                                                            reinterpret_cast<double *>(out), ldo);
            else
                mn_reduce_kernel<float><<<(total + 31) / 32, 256, 0, st>>>(workspace, (int)splits, p.m_pad, p.n_pad, m0, n0, m, n,
                                                                           reinterpret_cast<float *>(out), ldo);
            rc = check_launch("nt_gemm_tn(reduce)");
            if (rc) return rc;
        }
    }
    return 0;
}

}  // namespace nt
