// Persistent tcgen05 LSTM decoder, forward and backward (sm_100a) -- replaces the cuDNN call behind LSTMDecoderModule
// (nn/net_blocks.py:363-402: nn.LSTM(batch_first=True), 3 layers, H = 250, the encoding repeated T = 14 times).
//
// The recurrence is 3 x 14 dependent cells; every cell is a [rows, 512] x [512, 1000] product followed by element-wise gate
// math.  One cooperative kernel runs the whole recurrence:
//
//   * CTA (l, rt, c) = layer l, row tile rt (128 or 256 rows), unit slice c (16 of the 250 hidden units = 64 gate columns).
//     Its 64 x 512 slice of [W_ih | W_hh] (bf16 hi + lo planes, 128 KB) is loaded ONCE and stays in shared memory for all T
//     steps; the cell state of its (rows x 16) block stays in registers.
//   * Cells run as a dataflow wavefront over (layer, step): slice c of h_t is written to L2 by its CTA ALREADY in the byte layout
//     the tensor core reads from shared memory (UMMA K-major core matrices, bf16 hi / lo planes), one 16-element K-step per
//     slice, followed by a release flag; the 32 consumer CTAs (same layer, next step; next layer, same step) acquire the flags
//     slice by slice and pull the K-steps with cp.async.bulk straight into their operand ring -- no conversion, no generic-proxy
//     staging, no grid-wide barrier.
//   * warp roles: warps 0-7 gate math (thread = row, TMEM lane = row), warp 8 lane 0 = flag polling + bulk copies, warp 9
//     lane 0 = tcgen05.mma issue (three bf16 products hi.hi + hi.lo + lo.hi per K-step, fp32 accumulation in TMEM, accumulator
//     double-buffered so the input-part MMAs of step t+1 overlap the gate math of step t).
//
// Backward: same decomposition.  Phase 1 (per cell): gate gradients of the slice from the saved gates / cell states, written as
// split K-steps (K = the 4H gate columns).  Phase 2: dG_t [rows, 4H] x [W_ih | W_hh][:, slice] -> d(input) and d(h_{t-1}) for
// the slice's own 16 + 16 units (weights 1024 x 32, stationary); d(h_{t-1}) never leaves the CTA, d(input) goes to the one CTA
// of the layer below that owns those units.  Weight gradients: a third kernel contracts dG and [x | h_prev] over all (step, row)
// pairs, reading the SAME split blocks sideways as UMMA MN-major operands (no transposed copies, no fp32 copies); the bias
// gradient falls out of it through a ones column (lstm_layout.cuh).
//
// All addresses and the cell arithmetic live in lstm_layout.cuh and are exercised on the host by tools/lstm_emulate.cu.
#include "gemm_params.cuh"
#include "tc_common.cuh"
#include "lstm_layout.cuh"

namespace nt {
namespace lstm {
using namespace tc;

constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int NST = 6;                          // operand ring depth (stages of one K-step)
constexpr int TAIL_BYTES = 1024;                // barriers, TMEM slot, bias
constexpr size_t SMEM_BYTES = (size_t)FWD_W_BYTES + (size_t)NST * 2 * SUB_BYTES + TAIL_BYTES;     // 230400 <= 227 KB
static_assert(FWD_W_BYTES == BWD_W_BYTES, "both kernels share the shared-memory carve-up");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct WeightPtrs {
    const float *w_ih[MAX_LAYERS], *w_hh[MAX_LAYERS], *b_ih[MAX_LAYERS], *b_hh[MAX_LAYERS];
};

// ---- inter-CTA flags (global memory) ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// orders this thread's generic-proxy accesses (global AND shared) with the async proxy (cp.async.bulk of any CTA)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Bounded spin (a protocol bug must surface as a trap, never as a hung GPU): ~2^22 polls x >= 64 ns.
__device__ __forceinline__ void wait_flag(const uint32_t *p) {
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        if (ld_acquire(p) != 0u) return;
        __nanosleep(64);
    }
    __trap();
}
// Warp-wide polling: lane i watches flag i of the current step (nullptr = nothing to wait for).  Returns the number of stages,
// counted from `issued`, whose flags are all set (a prefix: stages are issued in a fixed order the MMA thread relies on).
// One round trip to L2 serves all 32 flags -- polling them one after the other from a single thread cost ~0.4 us per stage.
__device__ __forceinline__ int ready_prefix(const uint32_t *my_flag, bool &ready, int issued, int total) {
    if (!ready) {
        ready = (my_flag == nullptr) || (ld_acquire(my_flag) != 0u);
        if (ready) fence_proxy_async_all();       // this lane's bulk copy (async proxy) must observe what the flag publishes
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, ready);
    const uint32_t pending = ~(mask >> issued);                 // bit k set = stage issued + k not ready
    const int run = pending ? (__ffs((int)pending) - 1) : 32;
    return run < total - issued ? run : total - issued;
}

// All gate-math threads have written their part of a block: make it visible to the other CTAs (generic AND async proxy) and
// raise the flag(s).
__device__ __forceinline__ void publish(uint32_t *flag_a, uint32_t *flag_b, int tid) {
    fence_proxy_async_all();
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    if (tid == 0) {
        st_release(flag_a, 1u);
        if (flag_b) st_release(flag_b, 1u);
    }
}

// mbarrier wait with a short bound (tc::mbar_wait spins 2^28 times): while these kernels are young a protocol bug must end in
// a trap within seconds.  try_wait suspends up to the hinted 2 us per call.
__device__ __forceinline__ void mbar_wait_b(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 21); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(2000u)
            : "memory");
        if (done) return;
    }
    __trap();
}

// Development aid: per-CTA, per-step clock64 stamps of the three roles (nt_debug_lstm_trace sets the buffer; nullptr = off).
constexpr int TRACE_SLOTS = 16;
static unsigned long long *g_trace = nullptr;
__device__ __forceinline__ void stamp(unsigned long long *trace, int step, int T, int slot) {
    if (trace) trace[((size_t)blockIdx.x * T + step) * TRACE_SLOTS + slot] = (unsigned long long)clock64();
}

struct Smem {
    uint8_t *w, *ring;
    uint64_t *full, *empty, *tmem_full, *tmem_empty, *wbar;
    uint32_t *tmem_slot;
    float *bias;
};
__device__ __forceinline__ Smem carve(uint8_t *smem) {
    Smem s;
    s.w = smem;
    s.ring = smem + FWD_W_BYTES;
    uint8_t *tail = s.ring + (size_t)NST * 2 * SUB_BYTES;
    s.full = reinterpret_cast<uint64_t *>(tail);
    s.empty = s.full + NST;
    s.tmem_full = s.empty + NST;
    s.tmem_empty = s.tmem_full + 2;
    s.wbar = s.tmem_empty + 2;
    s.tmem_slot = reinterpret_cast<uint32_t *>(s.wbar + 1);
    s.bias = reinterpret_cast<float *>(tail + 512);
    return s;
}

// ==================================================================================================================================
// weight preparation (once per parameter version): hi / lo split in the two B-operand layouts + permuted bias sums
// ==================================================================================================================================
__global__ void lstm_prepare_weights_kernel(WeightPtrs w, int L, int H, int E, uint8_t *fwd, uint8_t *bwd, float *bias) {
    // one thread per 16-byte chunk (8 K-elements) of a hi plane
    const int64_t per_lc_f = (int64_t)FWD_KSTEPS * 2 * FWD_N, per_lc_b = (int64_t)BWD_KSTEPS * 2 * BWD_N;
    const int64_t n_f = (int64_t)L * SLICES * per_lc_f, n_b = (int64_t)L * SLICES * per_lc_b, n_bias = (int64_t)L * SLICES * FWD_N;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_f || i < n_f + n_b) {
        const bool is_f = i < n_f;
        if (!is_f) i -= n_f;
        const int64_t per = is_f ? per_lc_f : per_lc_b;
        const int N = is_f ? FWD_N : BWD_N;
        const int lc = (int)(i / per), within = (int)(i - (int64_t)lc * per);
        const int l = lc / SLICES, c = lc % SLICES;
        const int ks = within / (2 * N), rem = within % (2 * N), chunk = rem / N, n = rem % N;
        const int in_dim = l == 0 ? E : H;
        uint16_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float v = is_f ? fwd_w_value(w.w_ih[l], w.w_hh[l], in_dim, H, c, ks, n, 8 * chunk + e)
                                 : bwd_w_value(w.w_ih[l], w.w_hh[l], in_dim, H, c, ks, n, 8 * chunk + e);
            split_hi_lo(v, hi[e], lo[e]);
        }
        uint8_t *base = is_f ? fwd : bwd;
        const int64_t oh = is_f ? fwd_w_offset(l, c, ks, 0, n, 8 * chunk) : bwd_w_offset(l, c, ks, 0, n, 8 * chunk);
        const int64_t ol = is_f ? fwd_w_offset(l, c, ks, 1, n, 8 * chunk) : bwd_w_offset(l, c, ks, 1, n, 8 * chunk);
        *reinterpret_cast<uint4 *>(base + oh) = make_uint4(hi[0] | (uint32_t)hi[1] << 16, hi[2] | (uint32_t)hi[3] << 16,
                                                           hi[4] | (uint32_t)hi[5] << 16, hi[6] | (uint32_t)hi[7] << 16);
        *reinterpret_cast<uint4 *>(base + ol) = make_uint4(lo[0] | (uint32_t)lo[1] << 16, lo[2] | (uint32_t)lo[3] << 16,
                                                           lo[4] | (uint32_t)lo[5] << 16, lo[6] | (uint32_t)lo[7] << 16);
        return;
    }
    i -= n_f + n_b;
    if (i < n_bias) {
        const int lc = (int)(i / FWD_N), n = (int)(i % FWD_N);
        const int l = lc / SLICES, c = lc % SLICES, g = n >> 4, unit = UNITS * c + (n & 15);
        bias[i] = unit < H ? w.b_ih[l][g * H + unit] + w.b_hh[l][g * H + unit] : 0.f;
    }
}

// ==================================================================================================================================
// forward
// ==================================================================================================================================
struct FwdParams {
    Dims d;
    int rt0, nrt;                       // row tiles handled by this launch
    const float *x; int64_t ldx;        // [R, E] layer-0 input (the same vector at every step)
    const float *h0, *c0;               // [L, R, H]
    const uint8_t *w;                   // forward B operands, [L][16][128 KB]
    const float *bias;                  // [L][16][64]
    FwdOut out;
    uint32_t *flags;
    unsigned long long *trace;
};

__global__ void __launch_bounds__(THREADS, 1) lstm_fwd_kernel(FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const Smem s = carve(smem_raw);
    const Dims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x % SLICES;
    const int rt = p.rt0 + (blockIdx.x / SLICES) % p.nrt;
    const int l = blockIdx.x / (SLICES * p.nrt);
    const int stage_bytes = d.stage_bytes();
    const uint32_t tmem_cols = d.nsub == 2 ? 256u : 128u;       // 2 accumulator buffers x nsub x 64 columns

    if (warp == EPI_WARPS + 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s.tmem_full[i], 1); mbar_init(&s.tmem_empty[i], EPI_THREADS); }
        mbar_init(s.wbar, 1);
        mbar_fence_init();
    }
    if (warp == EPI_WARPS) tmem_alloc(s.tmem_slot, tmem_cols);
    if (tid < FWD_N) s.bias[tid] = p.bias[((int64_t)l * SLICES + c) * FWD_N + tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s.tmem_slot;

    if (warp < EPI_WARPS) {
        // =========================== gate math: thread = row ===========================
        const int sub = warp >> 2, quad = warp & 3;
        const bool active = sub < d.nsub;
        const int row_in_tile = sub * TILE_M + quad * 32 + lane;
        const int64_t grow = (int64_t)rt * d.tile_rows() + row_in_tile;
        const bool valid = active && grow < d.R;
        float cst[UNITS], h[UNITS];
        // ---- "step -1": publish the initial state (and, for layer 0, the input) in the operand layout
#pragma unroll
        for (int u = 0; u < UNITS; ++u) {
            const int unit = UNITS * c + u;
            const bool live = valid && unit < d.H;
            cst[u] = live ? __ldg(p.c0 + ((int64_t)l * d.R + grow) * d.H + unit) : 0.f;
            h[u] = live ? __ldg(p.h0 + ((int64_t)l * d.R + grow) * d.H + unit) : (unit == d.H ? 1.f : 0.f);     // ones column
        }
        if (active) {
            store_split16(p.out.act + act_block_index(d, l + 1, 0, rt) * d.act_block_bytes(), d, c, row_in_tile, h);
            if (l == 0) {
                float xv[UNITS];
#pragma unroll
                for (int u = 0; u < UNITS; ++u) xv[u] = (valid && UNITS * c + u < d.E) ? __ldg(p.x + grow * p.ldx + UNITS * c + u) : 0.f;
                store_split16(p.out.act + act_block_index(d, 0, 0, rt) * d.act_block_bytes(), d, c, row_in_tile, xv);
            }
            if (p.out.cs) store16_rowinner(p.out.cs + cs_offset(d, l, 0, rt, c), d.tile_rows(), row_in_tile, cst);
        }
        publish(p.flags + act_flag_index(d, l + 1, 0, rt, c), l == 0 ? p.flags + act_flag_index(d, 0, 0, rt, c) : nullptr, tid);

        for (int t = 0; t < d.T; ++t) {
            const int buf = t & 1;
            mbar_wait_b(&s.tmem_full[buf], (uint32_t)(t >> 1) & 1u);
            tc_fence_after();
            if (tid == 0) stamp(p.trace, t, d.T, 0);
            float acc[FWD_N];
            if (active) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * d.nsub * FWD_N + sub * FWD_N);
                float a0[32], a1[32];
                tmem_ld32(taddr, a0);
                tmem_ld32(taddr + 32, a1);
#pragma unroll
                for (int i = 0; i < 32; ++i) { acc[i] = a0[i]; acc[32 + i] = a1[i]; }
            }
            tc_fence_before();
            mbar_arrive(&s.tmem_empty[buf]);              // the accumulator buffer may be overwritten by step t + 2
            if (tid == 0) stamp(p.trace, t, d.T, 1);
            float gsave[4][UNITS];
            if (active) fwd_cell_row(d, p.out, l, t, rt, c, row_in_tile, acc, s.bias, cst, h, gsave);
            if (tid == 0) stamp(p.trace, t, d.T, 2);
            publish(p.flags + act_flag_index(d, l + 1, t + 1, rt, c), nullptr, tid);
            if (tid == 0) stamp(p.trace, t, d.T, 3);
            if (active) {
                if (l == d.L - 1) fwd_store_y(d, p.out.y, t, rt, c, row_in_tile, h);
                fwd_store_saved(d, p.out, l, t, rt, c, row_in_tile, gsave, cst);
            }
        }
    } else if (warp == EPI_WARPS) {
        // =========================== operand loader: flags -> cp.async.bulk into the ring ===========================
        // lane = stage of the step: lanes 0-15 the input part (h of the layer below at this step), lanes 16-31 the recurrent part
        // (own h of the previous step), slices in the rotated order c, c+1, ...
        if (lane == 0) {
            mbar_arrive_expect_tx(s.wbar, (uint32_t)FWD_W_BYTES);
            const uint8_t *wsrc = p.w + ((int64_t)l * SLICES + c) * FWD_W_BYTES;
            for (int i = 0; i < FWD_W_BYTES / 16384; ++i) bulk_g2s(s.w + i * 16384, wsrc + i * 16384, 16384u, s.wbar);
        }
        const int part = lane >> 4, cc = (c + (lane & 15)) & (SLICES - 1);
        const int src = part == 0 ? l : l + 1;
        uint32_t it = 0;
        for (int t = 0; t < d.T; ++t) {
            const int slot = part == 0 ? (l == 0 ? 0 : t + 1) : t;
            const uint32_t *my_flag = p.flags + act_flag_index(d, src, slot, rt, cc);
            const uint8_t *my_src = p.out.act + act_block_index(d, src, slot, rt) * d.act_block_bytes() + (size_t)cc * stage_bytes;
            bool ready = false;
            int issued = 0;
            if (lane == 0) stamp(p.trace, t, d.T, 4);
            for (uint32_t spin = 0; issued < 2 * SLICES; ++spin) {
                const int run = ready_prefix(my_flag, ready, issued, 2 * SLICES);
                if (run == 0) {
                    if (spin > (1u << 22)) __trap();
                    continue;
                }
                if (lane == 0 && issued == 0) stamp(p.trace, t, d.T, 5);
                if (lane == 0 && issued < SLICES && issued + run >= SLICES) stamp(p.trace, t, d.T, 6);
                for (int q = issued; q < issued + run; ++q, ++it) {
                    // the lane that owns stage q issues it (its own acquire + proxy fence order the copy after the producer's
                    // writes); the ring slot protocol is sequential, so lanes take turns
                    if (lane == q) {
                        const uint32_t st = it % NST;
                        mbar_wait_b(&s.empty[st], ((it / NST) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&s.full[st], (uint32_t)stage_bytes);
                        bulk_g2s(s.ring + (size_t)st * stage_bytes, my_src, (uint32_t)stage_bytes, &s.full[st]);
                    }
                    __syncwarp();
                }
                issued += run;
            }
            if (lane == 0) stamp(p.trace, t, d.T, 7);
        }
    } else {
        // =========================== MMA issuer (one thread) ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(TILE_M, FWD_N, 0, 0);
            mbar_wait_b(s.wbar, 0);
            uint32_t it = 0;
            for (int t = 0; t < d.T; ++t) {
                const int buf = t & 1;
                mbar_wait_b(&s.tmem_empty[buf], ((uint32_t)(t >> 1) & 1u) ^ 1u);
                tc_fence_after();
                stamp(p.trace, t, d.T, 8);
                for (int part = 0; part < 2; ++part) {
                    for (int j = 0; j < SLICES; ++j, ++it) {
                        const int cc = (c + j) & (SLICES - 1);
                        const uint32_t st = it % NST;
                        mbar_wait_b(&s.full[st], (it / NST) & 1u);
                        tc_fence_after();
                        if (j == 0) stamp(p.trace, t, d.T, 9 + part);
                        const uint32_t b = smem_u32(s.w + (size_t)(part * SLICES + cc) * FWD_W_KSTEP_BYTES);
                        const uint64_t dbh = make_smem_desc(b, 1024, 128), dbl = make_smem_desc(b + 2048, 1024, 128);
                        for (int sub = 0; sub < d.nsub; ++sub) {
                            const uint32_t a = smem_u32(s.ring + (size_t)st * stage_bytes + (size_t)sub * SUB_BYTES);
                            const uint64_t dah = make_smem_desc(a, CHUNK_BYTES, 128), dal = make_smem_desc(a + PLANE_BYTES, CHUNK_BYTES, 128);
                            const uint32_t dst = tmem_base + (uint32_t)(buf * d.nsub * FWD_N + sub * FWD_N);
                            umma_bf16(dst, dah, dbh, idesc, (part | j) ? 1u : 0u);
                            umma_bf16(dst, dah, dbl, idesc, 1u);
                            umma_bf16(dst, dal, dbh, idesc, 1u);
                        }
                        umma_commit(&s.empty[st]);
                    }
                }
                umma_commit(&s.tmem_full[buf]);
                stamp(p.trace, t, d.T, 11);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) tmem_dealloc(tmem_base, tmem_cols);
}

// ==================================================================================================================================
// backward
// ==================================================================================================================================
struct BwdParams {
    Dims d;
    int rt0, nrt;
    const uint8_t *w;                   // backward B operands, [L][16][128 KB]
    BwdIo io;
    uint32_t *flags_dg, *flags_dx;
    unsigned long long *trace;
};

__global__ void __launch_bounds__(THREADS, 1) lstm_bwd_kernel(BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const Smem s = carve(smem_raw);
    const Dims &d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x % SLICES;
    const int rt = p.rt0 + (blockIdx.x / SLICES) % p.nrt;
    const int l = blockIdx.x / (SLICES * p.nrt);
    const int stage_bytes = d.stage_bytes();
    const uint32_t tmem_cols = d.nsub == 2 ? 64u : 32u;         // nsub x 32 columns

    if (warp == EPI_WARPS + 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        mbar_init(&s.tmem_full[0], 1);
        mbar_init(&s.tmem_empty[0], EPI_THREADS);
        mbar_init(s.wbar, 1);
        mbar_fence_init();
    }
    if (warp == EPI_WARPS) tmem_alloc(s.tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s.tmem_slot;

    if (warp < EPI_WARPS) {
        const int sub = warp >> 2, quad = warp & 3;
        const bool active = sub < d.nsub;
        const int row_in_tile = sub * TILE_M + quad * 32 + lane;
        const int64_t grow = (int64_t)rt * d.tile_rows() + row_in_tile;
        const bool valid = active && grow < d.R;
        float dc[UNITS], dhrec[UNITS], dxsum[UNITS];
#pragma unroll
        for (int u = 0; u < UNITS; ++u) dc[u] = dhrec[u] = dxsum[u] = 0.f;
        for (int t = d.T - 1; t >= 0; --t) {
            const uint32_t step = (uint32_t)(d.T - 1 - t);
            // ---- phase 1: dh = upstream + recurrent part -> gate gradients of this slice
            float dh[UNITS];
            if (l == d.L - 1) {
#pragma unroll
                for (int u = 0; u < UNITS; ++u) {
                    const int unit = UNITS * c + u;
                    dh[u] = (valid && unit < d.H) ? __ldg(p.io.dy + ((int64_t)t * d.R + grow) * p.io.ld_dy + unit) : 0.f;
                }
            } else {
                if (lane == 0) wait_flag(p.flags_dx + dg_flag_index(d, l + 1, t, rt, c));
                __syncwarp();
                if (active) load16_rowinner(p.io.dxbuf + dx_block_offset(d, l + 1, t, rt, c), d.tile_rows(), row_in_tile, dh);
            }
            float dgo[4][UNITS];
            if (tid == 0) stamp(p.trace, (int)step, d.T, 0);
            if (active) {
#pragma unroll
                for (int u = 0; u < UNITS; ++u) dh[u] += dhrec[u];
                bwd_cell_row(d, p.io, l, t, rt, c, row_in_tile, dh, dc, dgo);
            }
            if (tid == 0) stamp(p.trace, (int)step, d.T, 1);
            publish(p.flags_dg + dg_flag_index(d, l, t, rt, c), nullptr, tid);
            if (tid == 0) stamp(p.trace, (int)step, d.T, 2);
            // ---- phase 2: [d input | d h_prev] of this slice's units = dG_t . [W_ih | W_hh][:, slice]
            mbar_wait_b(&s.tmem_full[0], step & 1u);
            tc_fence_after();
            if (tid == 0) stamp(p.trace, (int)step, d.T, 3);
            float v[32];
            if (active) tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * BWD_N), v);
            tc_fence_before();
            mbar_arrive(&s.tmem_empty[0]);
            if (active) {
#pragma unroll
                for (int u = 0; u < UNITS; ++u) dhrec[u] = v[UNITS + u];
                if (l > 0) {
                    float dx[UNITS];
#pragma unroll
                    for (int u = 0; u < UNITS; ++u) dx[u] = v[u];
                    store16_rowinner(p.io.dxbuf + dx_block_offset(d, l, t, rt, c), d.tile_rows(), row_in_tile, dx);
                } else {
#pragma unroll
                    for (int u = 0; u < UNITS; ++u) dxsum[u] += v[u];
                }
            }
            if (l > 0) publish(p.flags_dx + dg_flag_index(d, l, t, rt, c), nullptr, tid);
            if (tid == 0) stamp(p.trace, (int)step, d.T, 4);
        }
        if (l == 0 && valid && p.io.dx0) {
#pragma unroll
            for (int u = 0; u < UNITS; ++u)
                if (UNITS * c + u < d.E) p.io.dx0[grow * p.io.ld_dx0 + UNITS * c + u] = dxsum[u];
        }
    } else if (warp == EPI_WARPS) {
        // operand loader: lane i < 16 watches the dG flag of slice c + i; a ready slice releases its four K-steps (gates)
        if (lane == 0) {
            mbar_arrive_expect_tx(s.wbar, (uint32_t)BWD_W_BYTES);
            const uint8_t *wsrc = p.w + ((int64_t)l * SLICES + c) * BWD_W_BYTES;
            for (int i = 0; i < BWD_W_BYTES / 16384; ++i) bulk_g2s(s.w + i * 16384, wsrc + i * 16384, 16384u, s.wbar);
        }
        const int cc = (c + (lane & 15)) & (SLICES - 1);
        uint32_t it = 0;
        for (int t = d.T - 1; t >= 0; --t) {
            const uint32_t *my_flag = lane < SLICES ? p.flags_dg + dg_flag_index(d, l, t, rt, cc) : nullptr;
            const uint8_t *my_src = p.io.dgs + dg_block_index(d, l, t, rt) * d.dg_block_bytes() + (size_t)(4 * cc) * stage_bytes;
            bool ready = false;
            int issued = 0;
            for (uint32_t spin = 0; issued < SLICES; ++spin) {
                const int run = ready_prefix(my_flag, ready, issued, SLICES);
                if (run == 0) {
                    if (spin > (1u << 22)) __trap();
                    continue;
                }
                for (int q = issued; q < issued + run; ++q, it += 4) {
                    if (lane == q) {
                        for (int g = 0; g < 4; ++g) {
                            const uint32_t st = (it + g) % NST;
                            mbar_wait_b(&s.empty[st], (((it + g) / NST) & 1u) ^ 1u);
                            mbar_arrive_expect_tx(&s.full[st], (uint32_t)stage_bytes);
                            bulk_g2s(s.ring + (size_t)st * stage_bytes, my_src + (size_t)g * stage_bytes, (uint32_t)stage_bytes, &s.full[st]);
                        }
                    }
                    __syncwarp();
                }
                issued += run;
            }
        }
    } else {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(TILE_M, BWD_N, 0, 0);
            mbar_wait_b(s.wbar, 0);
            uint32_t it = 0;
            for (int t = d.T - 1; t >= 0; --t) {
                const uint32_t step = (uint32_t)(d.T - 1 - t);
                mbar_wait_b(&s.tmem_empty[0], (step & 1u) ^ 1u);
                tc_fence_after();
                stamp(p.trace, (int)step, d.T, 8);
                for (int j = 0; j < SLICES; ++j) {
                    const int cc = (c + j) & (SLICES - 1);
                    for (int g = 0; g < 4; ++g, ++it) {
                        const uint32_t st = it % NST;
                        mbar_wait_b(&s.full[st], (it / NST) & 1u);
                        tc_fence_after();
                        const uint32_t b = smem_u32(s.w + (size_t)(4 * cc + g) * BWD_W_KSTEP_BYTES);
                        const uint64_t dbh = make_smem_desc(b, 512, 128), dbl = make_smem_desc(b + 1024, 512, 128);
                        for (int sub = 0; sub < d.nsub; ++sub) {
                            const uint32_t a = smem_u32(s.ring + (size_t)st * stage_bytes + (size_t)sub * SUB_BYTES);
                            const uint64_t dah = make_smem_desc(a, CHUNK_BYTES, 128), dal = make_smem_desc(a + PLANE_BYTES, CHUNK_BYTES, 128);
                            const uint32_t dst = tmem_base + (uint32_t)(sub * BWD_N);
                            umma_bf16(dst, dah, dbh, idesc, (j | g) ? 1u : 0u);
                            umma_bf16(dst, dah, dbl, idesc, 1u);
                            umma_bf16(dst, dal, dbh, idesc, 1u);
                        }
                        umma_commit(&s.empty[st]);
                    }
                }
                umma_commit(&s.tmem_full[0]);
                stamp(p.trace, (int)step, d.T, 9);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) tmem_dealloc(tmem_base, tmem_cols);
}

// ==================================================================================================================================
// weight gradients:  raw[l][which][m][n] = sum over (t, row) of dG[(t,row), m] * B_which[(t,row), n]
// ==================================================================================================================================
constexpr int DW_EPI_WARPS = 4;
constexpr int DW_THREADS = (DW_EPI_WARPS + 2) * 32;
constexpr int DW_NST = 3;
constexpr size_t DW_SMEM_BYTES = (size_t)DW_NST * DW_STAGE_BYTES + 256;

struct DwParams {
    Dims d;
    const uint8_t *dgs, *act;
    float *partial;                     // [tile][split][256][256]
    int splits;
};

__global__ void __launch_bounds__(DW_THREADS, 1) lstm_dw_kernel(DwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const Dims &d = p.d;
    uint8_t *stages = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)DW_NST * DW_STAGE_BYTES);
    uint64_t *empty = full + DW_NST;
    uint64_t *tmem_full = empty + DW_NST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int l = tile / (2 * DW_M_TILES), which = (tile / DW_M_TILES) & 1, mt = tile % DW_M_TILES;
    const int n_kb_all = d.T * d.RT * d.nsub * 4;
    const int kb0 = (int)((int64_t)split * n_kb_all / p.splits), kb1 = (int)((int64_t)(split + 1) * n_kb_all / p.splits);

    if (warp == DW_EPI_WARPS + 1 && lane == 0) {
        for (int i = 0; i < DW_NST; ++i) { mbar_init(&full[i], 32); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == DW_EPI_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < DW_EPI_WARPS) {
        // =========================== epilogue: TMEM -> partial tile (thread = gate column) ===========================
        float *dst = p.partial + ((int64_t)tile * p.splits + split) * DW_TILE * DW_TILE;
        if (kb1 > kb0) {
            mbar_wait_b(tmem_full, 0);
            tc_fence_after();
        }
        for (int a = 0; a < 2; ++a)
            for (int c0 = 0; c0 < DW_TILE; c0 += 32) {
                float v[32];
                if (kb1 > kb0) {
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * DW_TILE + c0), v);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                float4 *o = reinterpret_cast<float4 *>(dst + (int64_t)(a * TILE_M + tid) * DW_TILE + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
    } else if (warp == DW_EPI_WARPS) {
        // =========================== loader: lane = 8-element group; 2 planes x (A, B) = four 512-byte copies per stage ===========
        for (int kb = kb0, i = 0; kb < kb1; ++kb, ++i) {
            const DwKBlock k = dw_kblock(d, kb);
            const uint32_t st = (uint32_t)i % DW_NST;
            mbar_wait_b(&empty[st], (((uint32_t)i / DW_NST) & 1u) ^ 1u);
            uint8_t *stage = stages + (size_t)st * DW_STAGE_BYTES;
            const uint8_t *a_blk = p.dgs + dg_block_index(d, l, k.t, k.rt) * d.dg_block_bytes();
            const uint8_t *b_blk = p.act + dw_b_block(d, l, which, k.t, k.rt) * d.act_block_bytes();
            mbar_arrive_expect_tx(&full[st], 4u * DW_GROUP_BYTES);
#pragma unroll
            for (int plane = 0; plane < 2; ++plane) {
                bulk_g2s(stage + plane * DW_PLANE_BYTES + lane * DW_GROUP_BYTES,
                         a_blk + group_offset(d, mt * DW_GROUPS + lane, k.sub, plane) + k.rowblk * DW_GROUP_BYTES, DW_GROUP_BYTES, &full[st]);
                bulk_g2s(stage + (2 + plane) * DW_PLANE_BYTES + lane * DW_GROUP_BYTES,
                         b_blk + group_offset(d, lane, k.sub, plane) + k.rowblk * DW_GROUP_BYTES, DW_GROUP_BYTES, &full[st]);
            }
        }
    } else if (lane == 0 && kb1 > kb0) {
        // =========================== MMA issuer: MN-major A and B, M128 N256 K16, two M halves ===========================
        const uint32_t idesc = make_idesc_bf16(TILE_M, DW_TILE, 1, 1);
        for (int kb = kb0, i = 0; kb < kb1; ++kb, ++i) {
            const uint32_t st = (uint32_t)i % DW_NST;
            mbar_wait_b(&full[st], ((uint32_t)i / DW_NST) & 1u);
            tc_fence_after();
            const uint32_t base = smem_u32(stages + (size_t)st * DW_STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < DW_KB_ROWS / KSTEP; ++j) {
                const uint64_t dbh = make_smem_desc(base + 2 * DW_PLANE_BYTES + j * 256, 128, DW_GROUP_BYTES);
                const uint64_t dbl = make_smem_desc(base + 3 * DW_PLANE_BYTES + j * 256, 128, DW_GROUP_BYTES);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const uint32_t aoff = base + a * (TILE_M / 8) * DW_GROUP_BYTES + j * 256;
                    const uint64_t dah = make_smem_desc(aoff, 128, DW_GROUP_BYTES), dal = make_smem_desc(aoff + DW_PLANE_BYTES, 128, DW_GROUP_BYTES);
                    const uint32_t dst = tmem_base + (uint32_t)(a * DW_TILE);
                    umma_bf16(dst, dah, dbh, idesc, (i | j) ? 1u : 0u);
                    umma_bf16(dst, dah, dbl, idesc, 1u);
                    umma_bf16(dst, dal, dbh, idesc, 1u);
                }
            }
            umma_commit(&empty[st]);
        }
        umma_commit(tmem_full);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == DW_EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

// reduce the split partials and scatter them into the PyTorch parameter layouts (gate rows un-permuted, padding dropped)
struct GradPtrs {
    float *dw_ih[MAX_LAYERS], *dw_hh[MAX_LAYERS], *db_ih[MAX_LAYERS], *db_hh[MAX_LAYERS];
};
__global__ void lstm_finish_grads_kernel(const float *__restrict__ partial, int splits, GradPtrs g, int L, int H, int E) {
    const int64_t per_layer = (int64_t)4 * H * 2 * HP;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_layer * L) return;
    const int l = (int)(i / per_layer);
    const int64_t r = i - (int64_t)l * per_layer;
    const int row = (int)(r / (2 * HP));                 // gate row g*H + unit
    const int col2 = (int)(r % (2 * HP));
    const int which = col2 / HP, col = col2 % HP;
    const int gate = row / H, unit = row % H;
    const int in_dim = l == 0 ? E : H;
    float *dst = nullptr;
    if (which == 0) {
        if (col < in_dim && g.dw_ih[l]) dst = g.dw_ih[l] + (int64_t)row * in_dim + col;
    } else if (col < H) {
        if (g.dw_hh[l]) dst = g.dw_hh[l] + (int64_t)row * H + col;
    } else if (col == H) {
        dst = g.db_ih[l] ? g.db_ih[l] + row : (g.db_hh[l] ? g.db_hh[l] + row : nullptr);
    }
    if (!dst) return;
    const int m = gate_perm_index(unit, gate);
    const int tile = (l * 2 + which) * DW_M_TILES + m / DW_TILE;
    const float *src = partial + ((int64_t)tile * splits * DW_TILE + (m % DW_TILE)) * DW_TILE + col;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += src[(int64_t)sp * DW_TILE * DW_TILE];
    *dst = acc;
    if (which == 1 && col == H && g.db_ih[l] && g.db_hh[l]) g.db_hh[l][row] = acc;
}

// ==================================================================================================================================
// host side
// ==================================================================================================================================
struct Plan {
    Dims d;
    int rtb;                 // row tiles per launch
    int sms;
    int dw_splits;
};

static int make_plan(int R, int T, int L, int H, int E, Plan &pl) {
    if (R < 1 || T < 1 || T > 255 || L < 1 || L > MAX_LAYERS) return fail("nt_lstm: need R >= 1, 1 <= T <= 255, 1 <= L <= %s4", "");
    if (H < 1 || H > HP - 1 || E < 1 || E > HP) return fail("nt_lstm: need 1 <= hidden <= 255 and 1 <= input <= 256%s", "");
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return fail("nt_lstm: cannot query the SM count%s", "");
    pl.sms = sms;
    pl.rtb = sms / (SLICES * L);
    if (pl.rtb < 1) return fail("nt_lstm: the device has too few SMs for %s layers x 16 slices of co-resident CTAs", "these");
    Dims &d = pl.d;
    d.R = R; d.T = T; d.L = L; d.H = H; d.E = E;
    d.nsub = (R > TILE_M * pl.rtb) ? 2 : 1;              // 128-row tiles while one launch covers all rows, else 256-row tiles
    d.RT = (R + d.tile_rows() - 1) / d.tile_rows();
    const int tiles = L * 2 * DW_M_TILES, n_kb = T * d.RT * d.nsub * 4;
    pl.dw_splits = sms / tiles < 1 ? 1 : sms / tiles;
    if (pl.dw_splits > n_kb) pl.dw_splits = n_kb;
    return 0;
}

struct Sizes {
    int64_t weights, act, flags_fwd, y, cs, gates, dgs, dxbuf, flags_bwd, partial;
};
static Sizes sizes_of(const Plan &pl) {
    const Dims &d = pl.d;
    Sizes z;
    z.weights = (int64_t)d.L * SLICES * (FWD_W_BYTES + BWD_W_BYTES) + (int64_t)d.L * SLICES * FWD_N * 4;
    z.act = (int64_t)(d.L + 1) * (d.T + 1) * d.RT * d.act_block_bytes();
    z.flags_fwd = (int64_t)(d.L + 1) * (d.T + 1) * d.RT * SLICES * 4;
    z.y = (int64_t)d.T * d.R * HP * 4;
    z.cs = (int64_t)d.L * (d.T + 1) * d.RT * SLICES * UNITS * d.tile_rows() * 4;
    z.gates = (int64_t)d.L * d.T * d.RT * SLICES * 4 * UNITS * d.tile_rows() * 4;
    z.dgs = (int64_t)d.L * d.T * d.RT * d.dg_block_bytes();
    z.dxbuf = (int64_t)d.L * d.T * d.RT * SLICES * d.tile_rows() * UNITS * 4;
    z.flags_bwd = (int64_t)2 * d.L * d.T * d.RT * SLICES * 4;
    z.partial = (int64_t)d.L * 2 * DW_M_TILES * pl.dw_splits * DW_TILE * DW_TILE * 4;
    return z;
}
static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

template <typename K, typename P>
static int launch_coop(K kernel, const P &params, int ctas, cudaStream_t st, const char *what) {
    static bool configured = false;          // one instance per (K, P) pair, i.e. per kernel
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return fail("nt_lstm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    void *args[] = {const_cast<P *>(&params)};
    // cooperative launch: the runtime guarantees that all CTAs are co-resident (they wait on each other's flags)
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)kernel, dim3((unsigned)ctas), dim3(THREADS), args, SMEM_BYTES, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

}  // namespace lstm
}  // namespace nt

using namespace nt;
using namespace nt::lstm;

// development only: device buffer of gridDim * T * 16 clock64 stamps written by the next nt_lstm_fwd / nt_lstm_bwd (nullptr = off)
extern "C" int nt_debug_lstm_trace(void *device_buffer) {
    g_trace = reinterpret_cast<unsigned long long *>(device_buffer);
    return 0;
}

extern "C" int nt_lstm_sizes(int R, int T, int L, int H, int E, nt_lstm_sizes_t *out) {
    NT_REQUIRE(out != nullptr, "nt_lstm_sizes: out is null");
    Plan pl;
    if (int rc = make_plan(R, T, L, H, E, pl)) return rc;
    const Sizes z = sizes_of(pl);
    out->weights_bytes = z.weights;
    out->act_bytes = z.act;
    out->fwd_workspace_bytes = align256(z.flags_fwd);
    out->y_bytes = z.y;
    out->cs_bytes = z.cs;
    out->gates_bytes = z.gates;
    out->bwd_workspace_bytes = align256(z.dgs) + align256(z.dxbuf) + align256(z.flags_bwd) + align256(z.partial);
    out->y_ld = HP;
    return 0;
}

extern "C" int nt_lstm_prepare_weights(const float *const *w_ih, const float *const *w_hh, const float *const *b_ih,
                                       const float *const *b_hh, int L, int H, int E, void *weights, void *stream) {
    NT_REQUIRE(w_ih && w_hh && b_ih && b_hh && weights, "nt_lstm_prepare_weights: null argument");
    NT_REQUIRE(L >= 1 && L <= MAX_LAYERS && H >= 1 && H < HP && E >= 1 && E <= HP, "nt_lstm_prepare_weights: unsupported sizes");
    NT_REQUIRE((reinterpret_cast<uintptr_t>(weights) & 127u) == 0, "nt_lstm_prepare_weights: weights must be 128-byte aligned");
    WeightPtrs w{};
    for (int l = 0; l < L; ++l) {
        NT_REQUIRE(w_ih[l] && w_hh[l] && b_ih[l] && b_hh[l], "nt_lstm_prepare_weights: null layer pointer");
        w.w_ih[l] = w_ih[l]; w.w_hh[l] = w_hh[l]; w.b_ih[l] = b_ih[l]; w.b_hh[l] = b_hh[l];
    }
    uint8_t *fwd = reinterpret_cast<uint8_t *>(weights);
    uint8_t *bwd = fwd + (int64_t)L * SLICES * FWD_W_BYTES;
    float *bias = reinterpret_cast<float *>(bwd + (int64_t)L * SLICES * BWD_W_BYTES);
    const int64_t total = (int64_t)L * SLICES * (FWD_KSTEPS * 2 * FWD_N + BWD_KSTEPS * 2 * BWD_N + FWD_N);
    lstm_prepare_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, L, H, E, fwd, bwd, bias);
    return check_launch("nt_lstm_prepare_weights");
}

extern "C" int nt_lstm_fwd(const float *x, int ldx, const float *h0, const float *c0, const void *weights, int R, int T, int L, int H,
                           int E, float *y, void *act, float *cs, float *gates, void *workspace, void *stream) {
    NT_REQUIRE(x && h0 && c0 && weights && y && act && workspace, "nt_lstm_fwd: null argument");
    NT_REQUIRE((cs == nullptr) == (gates == nullptr), "nt_lstm_fwd: cs and gates must both be given (training) or both be null");
    NT_REQUIRE(ldx >= E, "nt_lstm_fwd: ldx < E");
    NT_REQUIRE(aligned16(y) && aligned16(cs) && aligned16(gates) && (reinterpret_cast<uintptr_t>(act) & 255u) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
               "nt_lstm_fwd: y / cs / gates must be 16-byte aligned, act / workspace 256-byte aligned");
    Plan pl;
    if (int rc = make_plan(R, T, L, H, E, pl)) return rc;
    const Dims &d = pl.d;
    const Sizes z = sizes_of(pl);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FwdParams p{};
    p.d = d;
    p.x = x; p.ldx = ldx; p.h0 = h0; p.c0 = c0;
    p.w = reinterpret_cast<const uint8_t *>(weights);
    p.bias = reinterpret_cast<const float *>(p.w + (int64_t)L * SLICES * (FWD_W_BYTES + BWD_W_BYTES));
    p.out.act = reinterpret_cast<uint8_t *>(act);
    p.out.y = y; p.out.cs = cs; p.out.gates = gates;
    p.flags = reinterpret_cast<uint32_t *>(workspace);
    p.trace = g_trace;
    if (cudaMemsetAsync(p.flags, 0, (size_t)z.flags_fwd, st) != cudaSuccess) return fail("nt_lstm_fwd: cudaMemsetAsync failed%s", "");
    for (int rt0 = 0; rt0 < d.RT; rt0 += pl.rtb) {
        p.rt0 = rt0;
        p.nrt = d.RT - rt0 < pl.rtb ? d.RT - rt0 : pl.rtb;
        if (int rc = launch_coop(lstm_fwd_kernel, p, L * p.nrt * SLICES, st, "nt_lstm_fwd")) return rc;
    }
    return 0;
}

extern "C" int nt_lstm_bwd(const float *dy, int ld_dy, const void *act, const float *cs, const float *gates, const void *weights, int R,
                           int T, int L, int H, int E, void *workspace, float *dx, int ld_dx, float *const *dw_ih, float *const *dw_hh,
                           float *const *db_ih, float *const *db_hh, void *stream) {
    NT_REQUIRE(dy && act && cs && gates && weights && workspace, "nt_lstm_bwd: null argument");
    NT_REQUIRE(ld_dy >= H && (dx == nullptr || ld_dx >= E), "nt_lstm_bwd: bad leading dimension");
    NT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "nt_lstm_bwd: workspace must be 256-byte aligned");
    Plan pl;
    if (int rc = make_plan(R, T, L, H, E, pl)) return rc;
    const Dims &d = pl.d;
    const Sizes z = sizes_of(pl);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
    BwdParams p{};
    p.d = d;
    p.w = reinterpret_cast<const uint8_t *>(weights) + (int64_t)L * SLICES * FWD_W_BYTES;
    p.io.cs = cs; p.io.gates = gates; p.io.dy = dy; p.io.ld_dy = ld_dy;
    p.io.dgs = ws;                                       ws += align256(z.dgs);
    p.io.dxbuf = reinterpret_cast<float *>(ws);          ws += align256(z.dxbuf);
    p.flags_dg = reinterpret_cast<uint32_t *>(ws);
    p.flags_dx = p.flags_dg + (int64_t)d.L * d.T * d.RT * SLICES;
    ws += align256(z.flags_bwd);
    float *partial = reinterpret_cast<float *>(ws);
    p.io.dx0 = dx; p.io.ld_dx0 = ld_dx;
    p.trace = g_trace;
    if (cudaMemsetAsync(p.flags_dg, 0, (size_t)z.flags_bwd, st) != cudaSuccess) return fail("nt_lstm_bwd: cudaMemsetAsync failed%s", "");
    for (int rt0 = 0; rt0 < d.RT; rt0 += pl.rtb) {
        p.rt0 = rt0;
        p.nrt = d.RT - rt0 < pl.rtb ? d.RT - rt0 : pl.rtb;
        if (int rc = launch_coop(lstm_bwd_kernel, p, L * p.nrt * SLICES, st, "nt_lstm_bwd")) return rc;
    }
    // ---- weight gradients
    bool any = false;
    GradPtrs g{};
    for (int l = 0; l < L; ++l) {
        g.dw_ih[l] = dw_ih ? dw_ih[l] : nullptr; g.dw_hh[l] = dw_hh ? dw_hh[l] : nullptr;
        g.db_ih[l] = db_ih ? db_ih[l] : nullptr; g.db_hh[l] = db_hh ? db_hh[l] : nullptr;
        any = any || g.dw_ih[l] || g.dw_hh[l] || g.db_ih[l] || g.db_hh[l];
    }
    if (!any) return 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lstm_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SMEM_BYTES);
        if (e != cudaSuccess) return fail("nt_lstm_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    DwParams q{};
    q.d = d;
    q.dgs = p.io.dgs;
    q.act = reinterpret_cast<const uint8_t *>(act);
    q.partial = partial;
    q.splits = pl.dw_splits;
    lstm_dw_kernel<<<dim3((unsigned)(L * 2 * DW_M_TILES), (unsigned)pl.dw_splits), DW_THREADS, DW_SMEM_BYTES, st>>>(q);
    if (int rc = check_launch("nt_lstm_bwd(dw)")) return rc;
    const int64_t total = (int64_t)L * 4 * H * 2 * HP;
    lstm_finish_grads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, pl.dw_splits, g, L, H, E);
    return check_launch("nt_lstm_bwd(finish)");
}
