// Layouts and per-thread cell arithmetic of the persistent LSTM decoder kernels (lstm.cu).
//
// Everything in this header is plain C++ (`__host__ __device__`, no PTX): the kernels call these functions for every address
// they compute and for the cell arithmetic, and tools/lstm_emulate.cu executes THE SAME functions on the host, emulating only
// the tensor-core product (operands read back through the UMMA descriptor addressing) and the order of the cells -- so the
// permutations, offsets and formulas below are testable without a GPU.
//
// Reference semantics: nn.LSTM(batch_first=True) as used by LSTMDecoderModule (nn/net_blocks.py:363-402): L layers, hidden
// H, gates in PyTorch order i, f, g, o; the input of layer 0 is the SAME encoding at every step (net_blocks.py:388).
//
// Decomposition.  CTA (l, rt, c) owns layer l, row tile rt (128 * NSUB rows) and unit slice c (units 16c .. 16c+15; 16 slices
// cover H <= 256).  Operands travel between CTAs through L2 in the exact byte layout the tensor core reads from shared memory
// (UMMA K-major, no swizzle: 8 rows x 16 B core matrices), split into bf16 hi + lo planes (x = hi + lo + O(2^-17 |x|); three
// products hi.hi + hi.lo + lo.hi per K-step, ~1e-5 relative like the rest of the library's BF16x3 GEMMs):
//
//   "stage" = one MMA K-step (16 K-elements) of a row tile:  [sub NSUB][plane hi|lo][chunk 2][row 128][8 bf16] = NSUB * 8 KB
//   forward : K-step kc of an activation block holds units 16kc .. 16kc+15, i.e. exactly what slice kc produces;
//   backward: the K dimension are the 4H gate columns in the order (slice c, gate g, unit u): K-step 4c + g.
#pragma once
#include <stdint.h>
#include <math.h>
#ifdef __CUDACC__
#include <cuda_bf16.h>
#endif

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace nt {
namespace lstm {

constexpr int HP = 256;            // padded hidden / input width (K of one operand block)
constexpr int UNITS = 16;          // units per slice
constexpr int SLICES = 16;         // slices per layer
constexpr int TILE_M = 128;        // UMMA M
constexpr int KSTEP = 16;          // bf16 elements per MMA K-step
constexpr int SUB_BYTES = 8192;    // one 128-row sub tile of a stage: [plane 2][chunk 2][row 128][16 B]
constexpr int PLANE_BYTES = 4096;
constexpr int CHUNK_BYTES = 2048;  // = LBO of the A operand; SBO = 128
constexpr int FWD_N = 64;          // gate columns per CTA: 4 gates x 16 units
constexpr int FWD_KSTEPS = 32;     // 16 (input part) + 16 (recurrent part)
constexpr int FWD_W_KSTEP_BYTES = 4096;   // [plane 2][chunk 2][n 64][16 B]; LBO = 1024
constexpr int FWD_W_BYTES = FWD_KSTEPS * FWD_W_KSTEP_BYTES;       // 128 KB per (layer, slice)
constexpr int BWD_N = 32;          // output columns per CTA: 16 units of d(input) + 16 units of d(h_prev)
constexpr int BWD_KSTEPS = 64;     // 4H / 16 (padded: 16 slices x 4 gates)
constexpr int BWD_W_KSTEP_BYTES = 2048;   // [plane 2][chunk 2][n 32][16 B]; LBO = 512
constexpr int BWD_W_BYTES = BWD_KSTEPS * BWD_W_KSTEP_BYTES;       // 128 KB per (layer, slice)
constexpr int MAX_LAYERS = 4;

struct Dims {
    int R, T, L, H, E;             // rows, steps, layers, hidden, input width of layer 0
    int nsub;                      // 128-row sub tiles per row tile (1 or 2)
    int RT;                        // row tiles
    __host__ __device__ int tile_rows() const { return TILE_M * nsub; }
    __host__ __device__ int stage_bytes() const { return SUB_BYTES * nsub; }
    __host__ __device__ int64_t act_block_bytes() const { return (int64_t)SLICES * stage_bytes(); }      // 16 K-steps
    __host__ __device__ int64_t dg_block_bytes() const { return (int64_t)BWD_KSTEPS * stage_bytes(); }   // 64 K-steps
};

// ---- split (bf16 hi/lo) activation blocks: src 0 = layer-0 input (slot 0 only), src l+1 = h of layer l; slot 0 = initial state,
//      slot t+1 = h_t -------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t act_block_index(const Dims &d, int src, int slot, int rt) {
    return ((int64_t)src * (d.T + 1) + slot) * d.RT + rt;
}
// byte offset of element (row-in-tile, k) inside a block of `stage`s (k = K index; 16 per stage)
__host__ __device__ __forceinline__ int64_t split_offset(const Dims &d, int kstep, int row_in_tile, int k_in_step, int plane) {
    const int sub = row_in_tile >> 7, r = row_in_tile & 127;
    return (int64_t)kstep * d.stage_bytes() + (int64_t)sub * SUB_BYTES + plane * PLANE_BYTES + (k_in_step >> 3) * CHUNK_BYTES + r * 16 +
           (k_in_step & 7) * 2;
}
// flags: one per (src, slot, row tile, slice) -- set by the slice's CTA when its K-step of the block is complete
__host__ __device__ __forceinline__ int64_t act_flag_index(const Dims &d, int src, int slot, int rt, int c) {
    return act_block_index(d, src, slot, rt) * SLICES + c;
}
// backward: dG blocks and flags per (layer, t, row tile); dX hand-off buffers per (layer, t, row tile, slice)
__host__ __device__ __forceinline__ int64_t dg_block_index(const Dims &d, int l, int t, int rt) { return ((int64_t)l * d.T + t) * d.RT + rt; }
__host__ __device__ __forceinline__ int64_t dg_flag_index(const Dims &d, int l, int t, int rt, int c) { return dg_block_index(d, l, t, rt) * SLICES + c; }
// dX[l][t][rt][c]: row-inner block [u4 4][row tile_rows][4] fp32 -- gradient w.r.t. the input of layer l (= h of layer l-1),
// units 16c..16c+15, handed from CTA (l, rt, c) to CTA (l-1, rt, c)
__host__ __device__ __forceinline__ int64_t dx_block_offset(const Dims &d, int l, int t, int rt, int c) {
    return (dg_block_index(d, l, t, rt) * SLICES + c) * (int64_t)(d.tile_rows() * UNITS);
}

// ---- fp32 state ------------------------------------------------------------------------------------------------------------------
//   y    [T][R][HP] row-major, rows NOT padded: the top layer's h_t = the module output (time-major); columns >= H are scratch.
//   cs / gates: "row-inner" layout private to the kernels (thread = row writes / reads float4s that are contiguous across the
//        rows of a warp):  cs[l][slot][rt][c][u4 4][row][4]  (slot 0 = c0),  gates[l][t][rt][c][q 4][u4 4][row][4]  (i, f, g, o)
// The split activation blocks double as the B operand of the weight-gradient kernel; unit H of every h block holds 1.0 (a "ones
// column": the weight-gradient product then yields the bias gradient as its column H; the recurrence multiplies it with the
// zero-padded weight column H, so it is inert there).
__host__ __device__ __forceinline__ int64_t y_offset(const Dims &d, int t, int64_t row) { return ((int64_t)t * d.R + row) * HP; }
__host__ __device__ __forceinline__ int64_t cs_offset(const Dims &d, int l, int slot, int rt, int c) {       // start of [u4][row][4]
    return ((((int64_t)l * (d.T + 1) + slot) * d.RT + rt) * SLICES + c) * (int64_t)(UNITS * d.tile_rows());
}
__host__ __device__ __forceinline__ int64_t gates_offset(const Dims &d, int l, int t, int rt, int c, int q) {
    return (((((int64_t)l * d.T + t) * d.RT + rt) * SLICES + c) * 4 + q) * (int64_t)(UNITS * d.tile_rows());
}

// ---- weight-gradient kernel: dW = dG^T . [X | H_prev] with the contraction over ROWS --------------------------------------------
// Both operands are the split blocks above read "sideways": for a fixed 8-element group of gate columns / units (one 16-byte
// chunk), the 128 rows of a sub tile are 2 KB of contiguous 16-byte rows -- exactly a column of UMMA *MN-major* no-swizzle core
// matrices (8 k-rows x 16 B; descriptor LBO = 128 B between 8-row groups along K, SBO = distance between the 8-element groups
// along M / N; verified on a B200 with tools/microbench/mn_major_test.cu).  Group g of a block: K-step g / 2, chunk g % 2.
constexpr int DW_TILE = 256;                 // output tile: 256 gate columns x 256 units (two M = 128 accumulators, 512 TMEM columns)
constexpr int DW_GROUPS = DW_TILE / 8;       // 32 eight-element groups per operand
constexpr int DW_KB_ROWS = 32;               // rows (K) per pipeline stage
constexpr int DW_GROUP_BYTES = DW_KB_ROWS * 16;                   // 512 B per (group, plane, stage)
constexpr int DW_PLANE_BYTES = DW_GROUPS * DW_GROUP_BYTES;        // 16 KB
constexpr int DW_STAGE_BYTES = 4 * DW_PLANE_BYTES;                // A hi | A lo | B hi | B lo = 64 KB
constexpr int DW_M_TILES = 4 * HP / DW_TILE;                      // 4 tiles of gate columns (permuted order, see gate_perm_index)
__host__ __device__ __forceinline__ int64_t group_offset(const Dims &d, int group, int sub, int plane) {
    return (int64_t)(group >> 1) * d.stage_bytes() + (int64_t)sub * SUB_BYTES + plane * PLANE_BYTES + (group & 1) * CHUNK_BYTES;
}
// position of gate row g*H + unit in the K order of the dG blocks: (slice, gate, unit-in-slice)
__host__ __device__ __forceinline__ int gate_perm_index(int unit, int g) { return ((unit >> 4) * 4 + g) * UNITS + (unit & 15); }
struct DwKBlock { int t, rt, sub, rowblk; };
__host__ __device__ __forceinline__ DwKBlock dw_kblock(const Dims &d, int kb) {       // kb in [0, T * RT * nsub * 4)
    DwKBlock k;
    k.rowblk = kb & 3;
    const int piece = kb >> 2;
    k.sub = piece % d.nsub;
    const int trt = piece / d.nsub;
    k.rt = trt % d.RT;
    k.t = trt / d.RT;
    return k;
}
// source of the B operand of (layer l, which) at step t: which = 0 -> the layer's input, which = 1 -> its own previous h
__host__ __device__ __forceinline__ int64_t dw_b_block(const Dims &d, int l, int which, int t, int rt) {
    return which == 0 ? act_block_index(d, l, l == 0 ? 0 : t + 1, rt) : act_block_index(d, l + 1, t, rt);
}

// ---- weights -----------------------------------------------------------------------------------------------------------------
// forward B operand of CTA (l, c): column n = g*16 + u  <->  row g*H + 16c + u of W_ih / W_hh;  K-step ks < 16: input part
// (k = 16 ks + kk over the layer's input), ks >= 16: recurrent part.
__host__ __device__ __forceinline__ int64_t fwd_w_offset(int l, int c, int ks, int plane, int n, int kk) {
    return ((int64_t)l * SLICES + c) * FWD_W_BYTES + (int64_t)ks * FWD_W_KSTEP_BYTES + plane * 2048 + (kk >> 3) * 1024 + n * 16 + (kk & 7) * 2;
}
// backward B operand of CTA (l, c): column n < 16: W_ih^l[., 16c + n] (-> d input), n >= 16: W_hh^l[., 16c + n - 16] (-> d h_prev);
// K-step ks = 4 c' + g, element kk = u: gate row g*H + 16c' + u.
__host__ __device__ __forceinline__ int64_t bwd_w_offset(int l, int c, int ks, int plane, int n, int kk) {
    return ((int64_t)l * SLICES + c) * BWD_W_BYTES + (int64_t)ks * BWD_W_KSTEP_BYTES + plane * 1024 + (kk >> 3) * 512 + n * 16 + (kk & 7) * 2;
}

// bf16 round-to-nearest-even on the host and the device (bit pattern), without cuda_bf16.h so that the emulator shares it
__host__ __device__ __forceinline__ uint16_t f32_to_bf16_bits(float x) {
    union { float f; uint32_t u; } v;
    v.f = x;
    if ((v.u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(v.u >> 16);          // inf / nan: truncate
    const uint32_t lsb = (v.u >> 16) & 1u;
    v.u += 0x7FFFu + lsb;
    return (uint16_t)(v.u >> 16);
}
__host__ __device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) {
    union { float f; uint32_t u; } v;
    v.u = (uint32_t)b << 16;
    return v.f;
}
__host__ __device__ __forceinline__ void split_hi_lo(float x, uint16_t &hi, uint16_t &lo) {
    hi = f32_to_bf16_bits(x);
    lo = f32_to_bf16_bits(x - bf16_bits_to_f32(hi));
}

// element of the forward / backward B operand (value before the split); w_ih: [4H, in], w_hh: [4H, H]
__host__ __device__ __forceinline__ float fwd_w_value(const float *w_ih, const float *w_hh, int in_dim, int H, int c, int ks, int n, int kk) {
    const int g = n >> 4, u = n & 15, unit = 16 * c + u;
    if (unit >= H) return 0.f;
    const int64_t row = (int64_t)g * H + unit;
    if (ks < SLICES) {
        const int k = 16 * ks + kk;
        return k < in_dim ? w_ih[row * in_dim + k] : 0.f;
    }
    const int k = 16 * (ks - SLICES) + kk;
    return k < H ? w_hh[row * H + k] : 0.f;
}
__host__ __device__ __forceinline__ float bwd_w_value(const float *w_ih, const float *w_hh, int in_dim, int H, int c, int ks, int n, int kk) {
    const int cp = ks >> 2, g = ks & 3, unit = 16 * cp + kk;
    if (unit >= H) return 0.f;
    const int64_t row = (int64_t)g * H + unit;
    if (n < UNITS) {
        const int col = 16 * c + n;
        return col < in_dim ? w_ih[row * in_dim + col] : 0.f;
    }
    const int col = 16 * c + n - UNITS;
    return col < H ? w_hh[row * H + col] : 0.f;
}

// ---- cell arithmetic -----------------------------------------------------------------------------------------------------------
// Gate non-linearities.  On the device: ex2.approx / rcp.approx based (a handful of instructions, ~1e-6 relative; the libm
// versions cost ~10x more and made the gate math the longest phase of a step in the first cycle trace).
// tanh(x) = 2 sigmoid(2x) - 1: absolute error ~2e-7, which is what matters for O(1) activations.
__host__ __device__ __forceinline__ float sigmoidf_(float x) {
#ifdef __CUDA_ARCH__
    return __fdividef(1.f, 1.f + __expf(-x));
#else
    return 1.f / (1.f + expf(-x));
#endif
}
__host__ __device__ __forceinline__ float tanhf_(float x) {
#ifdef __CUDA_ARCH__
    return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f);
#else
    return tanhf(x);
#endif
}
__host__ __device__ __forceinline__ void return_h(const float (&h)[UNITS], float (&out)[UNITS]) {
#pragma unroll
    for (int u = 0; u < UNITS; ++u) out[u] = h[u];
}

// write 16 values of one row into a split block (two 16-byte chunks per plane)
__host__ __device__ __forceinline__ void store_split16(uint8_t *block, const Dims &d, int kstep, int row_in_tile, const float (&v)[UNITS]) {
#ifndef __CUDA_ARCH__
    uint16_t hi[UNITS], lo[UNITS];
#pragma unroll
    for (int u = 0; u < UNITS; ++u) split_hi_lo(v[u], hi[u], lo[u]);
#endif
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#ifdef __CUDA_ARCH__
            // packed conversion (one F2FP per pair and term); same round-to-nearest-even bits as split_hi_lo for finite values
            const float a = v[8 * ch + 2 * e], b = v[8 * ch + 2 * e + 1];
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            ph[e] = *reinterpret_cast<uint32_t *>(&h2);
            __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(ph[e] << 16), b - __uint_as_float(ph[e] & 0xFFFF0000u));
            pl[e] = *reinterpret_cast<uint32_t *>(&l2);
#else
            ph[e] = (uint32_t)hi[8 * ch + 2 * e] | ((uint32_t)hi[8 * ch + 2 * e + 1] << 16);
            pl[e] = (uint32_t)lo[8 * ch + 2 * e] | ((uint32_t)lo[8 * ch + 2 * e + 1] << 16);
#endif
        }
        uint32_t *dh = reinterpret_cast<uint32_t *>(block + split_offset(d, kstep, row_in_tile, 8 * ch, 0));
        uint32_t *dl = reinterpret_cast<uint32_t *>(block + split_offset(d, kstep, row_in_tile, 8 * ch, 1));
#ifdef __CUDA_ARCH__
        *reinterpret_cast<uint4 *>(dh) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4 *>(dl) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
#else
        for (int e = 0; e < 4; ++e) { dh[e] = ph[e]; dl[e] = pl[e]; }
#endif
    }
}

__host__ __device__ __forceinline__ void store16(float *dst, const float (&v)[UNITS]) {       // dst 16-byte aligned
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4 *>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
#else
    for (int i = 0; i < UNITS; ++i) dst[i] = v[i];
#endif
}
__host__ __device__ __forceinline__ void load16(const float *src, float (&v)[UNITS]) {        // src 16-byte aligned
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 a = reinterpret_cast<const float4 *>(src)[i];
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
#else
    for (int i = 0; i < UNITS; ++i) v[i] = src[i];
#endif
}

// "row-inner" blocks: [u4 4][row tile_rows][4 floats]
__host__ __device__ __forceinline__ void store16_rowinner(float *blk, int tile_rows, int row_in_tile, const float (&v)[UNITS]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float *dst = blk + ((int64_t)q * tile_rows + row_in_tile) * 4;
#ifdef __CUDA_ARCH__
        *reinterpret_cast<float4 *>(dst) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#else
        for (int e = 0; e < 4; ++e) dst[e] = v[4 * q + e];
#endif
    }
}
__host__ __device__ __forceinline__ void load16_rowinner(const float *blk, int tile_rows, int row_in_tile, float (&v)[UNITS]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float *src = blk + ((int64_t)q * tile_rows + row_in_tile) * 4;
#ifdef __CUDA_ARCH__
        const float4 a = __ldcg(reinterpret_cast<const float4 *>(src));      // L2 only: other CTAs of this launch wrote it
        v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
#else
        for (int e = 0; e < 4; ++e) v[4 * q + e] = src[e];
#endif
    }
}

struct FwdOut {                 // global buffers of the forward kernel (see the offsets above)
    uint8_t *act;               // split activation blocks
    float *y;                   // row-major fp32 h of the top layer (see above)
    float *cs, *gates;          // row-inner saved state; may be null (inference)
};

// One row of cell (l, t) for unit slice c: `acc` = the 64 gate pre-activations of this row WITHOUT bias, columns g*16 + u;
// `cst` = c_{t-1} in, c_t out (kept in registers by the kernel).  Writes h_t as split K-step c of block (l+1, t+1);
// returns h_t in `hout` and the activated gates in `gsave` (fwd_store_hf / fwd_store_saved write them after the flag).
__host__ __device__ __forceinline__ void fwd_cell_row(const Dims &d, const FwdOut &o, int l, int t, int rt, int c, int row_in_tile,
                                                      const float (&acc)[FWD_N], const float *bias64, float (&cst)[UNITS],
                                                      float (&hout)[UNITS], float (&gsave)[4][UNITS]) {
    const int64_t grow = (int64_t)rt * d.tile_rows() + row_in_tile;
    const bool valid = grow < d.R;
    float gi[UNITS], gf[UNITS], gg[UNITS], go[UNITS], h[UNITS];
#pragma unroll
    for (int u = 0; u < UNITS; ++u) {
        const bool live = valid && (UNITS * c + u) < d.H;
        gi[u] = sigmoidf_(acc[u] + bias64[u]);
        gf[u] = sigmoidf_(acc[16 + u] + bias64[16 + u]);
        gg[u] = tanhf_(acc[32 + u] + bias64[32 + u]);
        go[u] = sigmoidf_(acc[48 + u] + bias64[48 + u]);
        const float cn = gf[u] * cst[u] + gi[u] * gg[u];
        cst[u] = live ? cn : 0.f;
        h[u] = live ? go[u] * tanhf_(cn) : ((UNITS * c + u == d.H) ? 1.f : 0.f);       // ones column at unit H
    }
    uint8_t *blk = o.act + act_block_index(d, l + 1, t + 1, rt) * d.act_block_bytes();
    store_split16(blk, d, c, row_in_tile, h);
    return_h(h, hout);
#pragma unroll
    for (int u = 0; u < UNITS; ++u) { gsave[0][u] = gi[u]; gsave[1][u] = gf[u]; gsave[2][u] = gg[u]; gsave[3][u] = go[u]; }
}

// state kept for the backward (row-inner layout); issued AFTER the flag of the step: off the critical path
__host__ __device__ __forceinline__ void fwd_store_saved(const Dims &d, const FwdOut &o, int l, int t, int rt, int c, int row_in_tile,
                                                        const float (&gsave)[4][UNITS], const float (&cst)[UNITS]) {
    if (o.cs) store16_rowinner(o.cs + cs_offset(d, l, t + 1, rt, c), d.tile_rows(), row_in_tile, cst);
    if (o.gates) {
#pragma unroll
        for (int q = 0; q < 4; ++q) store16_rowinner(o.gates + gates_offset(d, l, t, rt, c, q), d.tile_rows(), row_in_tile, gsave[q]);
    }
}

// fp32 copy of the top layer's h_t for the caller (row-major; issued AFTER the flag: off the critical path)
__host__ __device__ __forceinline__ void fwd_store_y(const Dims &d, float *y, int t, int rt, int c, int row_in_tile, const float (&h)[UNITS]) {
    const int64_t grow = (int64_t)rt * d.tile_rows() + row_in_tile;
    if (grow < d.R) store16(y + y_offset(d, t, grow) + UNITS * c, h);
}

struct BwdIo {
    const float *cs, *gates;          // saved by the forward (row-inner)
    const float *dy; int64_t ld_dy;   // [T][R][ld_dy] gradient of the top layer's outputs
    uint8_t *dgs;                     // split dG blocks
    float *dxbuf;                     // dX hand-off buffers (row-inner [u4][row][4] per (l, t, rt, c))
    float *dx0; int64_t ld_dx0;       // [R][ld_dx0] gradient w.r.t. the layer-0 input (sum over t), written at the end
};

// Phase 1 of backward cell (l, t): from dh (= dy or dX from the layer above, + recurrent part) and the saved state, the gate
// gradients of this row / slice; dc carried in registers.  Writes the split K-steps 4c..4c+3.
__host__ __device__ __forceinline__ void bwd_cell_row(const Dims &d, const BwdIo &io, int l, int t, int rt, int c, int row_in_tile,
                                                      const float (&dh)[UNITS], float (&dc)[UNITS], float (&dgo)[4][UNITS]) {
    const int64_t grow = (int64_t)rt * d.tile_rows() + row_in_tile;
    const bool valid = grow < d.R;
    const int tr = d.tile_rows();
    float gi[UNITS], gf[UNITS], gg[UNITS], go[UNITS], ct[UNITS], cp[UNITS];
    load16_rowinner(io.gates + gates_offset(d, l, t, rt, c, 0), tr, row_in_tile, gi);
    load16_rowinner(io.gates + gates_offset(d, l, t, rt, c, 1), tr, row_in_tile, gf);
    load16_rowinner(io.gates + gates_offset(d, l, t, rt, c, 2), tr, row_in_tile, gg);
    load16_rowinner(io.gates + gates_offset(d, l, t, rt, c, 3), tr, row_in_tile, go);
    load16_rowinner(io.cs + cs_offset(d, l, t + 1, rt, c), tr, row_in_tile, ct);
    load16_rowinner(io.cs + cs_offset(d, l, t, rt, c), tr, row_in_tile, cp);
#pragma unroll
    for (int u = 0; u < UNITS; ++u) {
        const bool live = valid && (UNITS * c + u) < d.H;
        const float tc = tanhf_(ct[u]);
        const float dcu = dc[u] + dh[u] * go[u] * (1.f - tc * tc);
        dgo[3][u] = live ? dh[u] * tc * go[u] * (1.f - go[u]) : 0.f;
        dgo[0][u] = live ? dcu * gg[u] * gi[u] * (1.f - gi[u]) : 0.f;
        dgo[1][u] = live ? dcu * cp[u] * gf[u] * (1.f - gf[u]) : 0.f;
        dgo[2][u] = live ? dcu * gi[u] * (1.f - gg[u] * gg[u]) : 0.f;
        dc[u] = live ? dcu * gf[u] : 0.f;
    }
    uint8_t *blk = io.dgs + dg_block_index(d, l, t, rt) * d.dg_block_bytes();
#pragma unroll
    for (int g = 0; g < 4; ++g) store_split16(blk, d, 4 * c + g, row_in_tile, dgo[g]);
}

}  // namespace lstm
}  // namespace nt
