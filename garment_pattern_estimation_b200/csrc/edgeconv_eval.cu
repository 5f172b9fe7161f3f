// Fused inference EdgeConv layer (BASELINE.json north_star: "EdgeConv neighbor-gather + edge-MLP + max-reduce fused into a single
// kernel"; reference nn/net_blocks.py:126-135,172-180 through torch_geometric.nn.DynamicEdgeConv):
//
//     out_i = BN3( max_{j in kNN(i)}  relu(W3' relu(W2' relu(P_i + Q_j) + b2') + b3') )          (eval mode, k neighbours)
//
// In eval mode BatchNorm is a constant affine, so BN1 / BN2 are folded into W2' / W3' (nt_bn_fold) and BN3 -- monotone per
// channel -- is applied after the max (max for s >= 0, min for s < 0).  P | Q = X . [Wa - Wb | Wb]^T is the first Linear evaluated
// once per POINT (the algebraic split of W1 . [x_i, x_j - x_i]).  Nothing edge-sized ever reaches HBM: per tile of 128 edge rows
//
//   gather  : thread = edge row; relu(P[centre] + Q[nbr]) in 32-column K blocks, split into bf16 hi / lo planes, written to the
//             shared-memory ring directly in the UMMA K-major core-matrix layout
//   GEMM 1  : tcgen05, error-compensated 3-product split (hi.hi + hi.lo + lo.hi; TF32x3 by default -- the forward path keeps fp32-class
//             accuracy so that the kNN graph of the NEXT layer matches the reference's, see DESIGN.md -- or BF16x3), D1[128, H2] in
//             TMEM; W2' K blocks arrive by cp.async.bulk
//   bridge  : D1 is read back 32 columns at a time (thread = row = TMEM lane), + b2', relu, split, and becomes -- through the same
//             ring -- the A operand of
//   GEMM 2  : D2[128, C] in TMEM; W3' K blocks by cp.async.bulk
//   finish  : + b3', relu, max / min over the k rows of every point, BN3 affine, skip-connection columns -> out[M, C + tail]
//
// HBM traffic per layer = PQ rows (L2-resident gathers) + idx + out.  Training mode cannot use this kernel: batch statistics of
// each BatchNorm need a grid-wide reduction between the layers.
#include "gemm_tc_shared.cuh"

namespace nt {

constexpr int EE_WORKERS = 256;                 // warps 0-7: gather / bridge / finish; two threads per edge row (TMEM lane = row), each
                                                // owning one half (16 columns) of every 32-column K block
constexpr int EE_WORKER_WARPS = EE_WORKERS / 32;
constexpr int EE_THREADS = EE_WORKERS + 64;     // warp 8: MMA issuer, warp 9: weight loader + TMEM
constexpr int EE_NST = 3;
constexpr int EE_A_PLANE = 4 * TC_M * 16;       // 8 KB: [chunk 4][row 128][16 B]
constexpr int EE_MAX_N = 224;
constexpr int EE_STAGE_BYTES = 2 * EE_A_PLANE + EE_MAX_N * 128;         // A hi | A lo | B (hi | lo planes of n_tile rows) = 45056

struct EEParams {
    const float *pq; int ldpq; int H1;
    const int32_t *idx; int k; int n_per_cloud; int64_t M;
    const uint8_t *w2; const float *b2; int H2; int n2;     // n2 = padded N of GEMM 1 (multiple of 16)
    const uint8_t *w3; const float *b3; int C; int n3;
    const float *s_out, *t_out;
    const float *tail_src; int tail_ld; int tail;
    float *out; int ldo;
    int rows_per_tile; int64_t n_tiles;
};

template <bool TF32>
__global__ void __launch_bounds__(EE_THREADS, 1) edgeconv_eval_kernel(EEParams p) {
    constexpr int EPC = TF32 ? 4 : 8;               // elements per 16-byte chunk
    constexpr int EE_KB = 4 * EPC;                  // elements per K block (4 chunks): 16 tf32 or 32 bf16
    constexpr int HC = 2 * EPC;                     // columns of a K block owned by one worker half
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *tail_s = smem + (size_t)EE_NST * EE_STAGE_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(tail_s);           // [NST]
    uint64_t *empty = full + EE_NST;                                 // [NST]
    uint64_t *d1_full = empty + EE_NST, *d2_full = d1_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d2_full + 1);
    float *vec = reinterpret_cast<float *>(tail_s + 128);            // b2 [224] | b3 [160] | s [160] | t [160]
    float *vt_all = vec + 224 + 3 * 160;                             // 2 x [128][33] transposition tiles (one per worker half)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kb1 = (p.H1 + EE_KB - 1) / EE_KB, kb2 = (p.H2 + EE_KB - 1) / EE_KB;
    const uint32_t d2_col = 224;                                     // D1: TMEM columns [0, 224), D2: [224, 224 + n3)

    for (int i = tid; i < 224; i += EE_THREADS) vec[i] = i < p.H2 ? p.b2[i] : 0.f;
    for (int i = tid; i < 160; i += EE_THREADS) {
        vec[224 + i] = i < p.C ? p.b3[i] : 0.f;
        vec[384 + i] = i < p.C ? p.s_out[i] : 0.f;
        vec[544 + i] = i < p.C ? p.t_out[i] : 0.f;
    }
    if (warp == EE_WORKER_WARPS && lane == 0) {
        for (int s = 0; s < EE_NST; ++s) { mbar_init(&full[s], EE_WORKERS + 1); mbar_init(&empty[s], 1); }
        mbar_init(d1_full, 1);
        mbar_init(d2_full, 1);
        mbar_fence_init();
    }
    if (warp == EE_WORKER_WARPS + 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < EE_WORKER_WARPS) {
        // =========================== workers: (row r, half) -- 16 of the 32 columns of every K block ===========================
        const int r = tid & (TC_M - 1), half = tid >> 7, quad = warp & 3;
        float *vt = vt_all + half * (TC_M * 33);
        uint32_t it = 0;                                             // ring position (shared numbering with the other roles)
        uint32_t tile_i = 0;
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_i) {
            const int64_t row0 = tile * p.rows_per_tile;
            const int64_t E = p.M * p.k;
            const bool live = r < p.rows_per_tile && row0 + r < E;
            const float *pp = nullptr, *qq = nullptr;
            if (live) {
                const int64_t e = row0 + r, centre = e / p.k;
                const int64_t base = (centre / p.n_per_cloud) * (int64_t)p.n_per_cloud;
                pp = p.pq + centre * p.ldpq;
                qq = p.pq + (base + p.idx[e]) * p.ldpq + p.H1;
            }
            // ---- gather -> A operand of GEMM 1 (the loads of K block kb + 1 are in flight while block kb is converted)
            float4 pa[HC / 4], pb[HC / 4];
            auto fetch = [&](int kb) {
#pragma unroll
                for (int j = 0; j < HC / 4; ++j) {
                    const int c = kb * EE_KB + half * HC + 4 * j;
                    pa[j] = pb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && kb < kb1 && c < p.H1) {              // H1 % 4 == 0 and 16-byte aligned rows (checked by the launcher)
                        pa[j] = __ldg(reinterpret_cast<const float4 *>(pp + c));
                        pb[j] = __ldg(reinterpret_cast<const float4 *>(qq + c));
                    }
                }
            };
            fetch(0);
            for (int kb = 0; kb < kb1; ++kb, ++it) {
                float v[HC];
#pragma unroll
                for (int j = 0; j < HC / 4; ++j) {
                    v[4 * j] = fmaxf(pa[j].x + pb[j].x, 0.f); v[4 * j + 1] = fmaxf(pa[j].y + pb[j].y, 0.f);
                    v[4 * j + 2] = fmaxf(pa[j].z + pb[j].z, 0.f); v[4 * j + 3] = fmaxf(pa[j].w + pb[j].w, 0.f);
                }
                fetch(kb + 1);
                const uint32_t st = it % EE_NST;
                mbar_wait(&empty[st], ((it / EE_NST) & 1u) ^ 1u);
                uint8_t *a_hi = smem + (size_t)st * EE_STAGE_BYTES, *a_lo = a_hi + EE_A_PLANE;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float t8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) t8[e] = e < EPC ? v[EPC * ch + e] : 0.f;
                    uint4 h, l;
                    pack_chunk(t8, TF32, h, l);
                    *reinterpret_cast<uint4 *>(a_hi + (2 * half + ch) * (TC_M * 16) + r * 16) = h;
                    *reinterpret_cast<uint4 *>(a_lo + (2 * half + ch) * (TC_M * 16) + r * 16) = l;
                }
                fence_proxy_async();
                mbar_arrive(&full[st]);
            }
            // ---- bridge: D1 -> relu(D1 + b2') -> A operand of GEMM 2
            mbar_wait(d1_full, tile_i & 1u);
            tc_fence_after();
            for (int kb = 0; kb < kb2; ++kb, ++it) {
                float acc[HC];
                if constexpr (TF32) tmem_ld8(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kb * EE_KB + half * HC), acc);
                else tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kb * EE_KB + half * HC), acc);
                const uint32_t st = it % EE_NST;
                mbar_wait(&empty[st], ((it / EE_NST) & 1u) ^ 1u);
                uint8_t *a_hi = smem + (size_t)st * EE_STAGE_BYTES, *a_lo = a_hi + EE_A_PLANE;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    float t8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = kb * EE_KB + half * HC + EPC * ch + e;
                        t8[e] = (e < EPC && live && c < p.H2) ? fmaxf(acc[(EPC * ch + e) % HC] + vec[c < 224 ? c : 223], 0.f) : 0.f;
                    }
                    uint4 h, l;
                    pack_chunk(t8, TF32, h, l);
                    *reinterpret_cast<uint4 *>(a_hi + (2 * half + ch) * (TC_M * 16) + r * 16) = h;
                    *reinterpret_cast<uint4 *>(a_lo + (2 * half + ch) * (TC_M * 16) + r * 16) = l;
                }
                fence_proxy_async();
                tc_fence_before();               // the TMEM reads of this K block precede the arrive the MMA thread waits on
                mbar_arrive(&full[st]);
            }
            // ---- finish: D2 -> relu(+ b3') -> max / min over the k rows of a point -> BN3 -> out; the two halves take alternate
            //      32-column chunks, each with its own transposition tile and named barrier
            mbar_wait(d2_full, tile_i & 1u);
            tc_fence_after();
            const int nodes_here = (int)min((int64_t)(p.rows_per_tile / p.k), p.M - row0 / p.k);
            const int64_t node0 = row0 / p.k;
            for (int c0 = half * 32; c0 < p.C; c0 += 64) {
                float acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + d2_col + (uint32_t)c0, acc);
#pragma unroll
                for (int i = 0; i < 32; ++i) vt[r * 33 + i] = fmaxf(acc[i] + vec[224 + ((c0 + i) < 160 ? c0 + i : 159)], 0.f);
                if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                for (int t = r; t < nodes_here * 32; t += TC_M) {
                    const int nd = t >> 5, cc = t & 31, c = c0 + cc;
                    if (c >= p.C) continue;
                    float mx = vt[(nd * p.k) * 33 + cc], mn = mx;
                    for (int sl = 1; sl < p.k; ++sl) {
                        const float x = vt[(nd * p.k + sl) * 33 + cc];
                        mx = fmaxf(mx, x);
                        mn = fminf(mn, x);
                    }
                    const float s = vec[384 + c];
                    p.out[(node0 + nd) * (int64_t)p.ldo + c] = fmaf(s, s >= 0.f ? mx : mn, vec[544 + c]);
                }
                if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
            }
            if (p.tail)
                for (int t = tid; t < nodes_here * p.tail; t += EE_WORKERS) {
                    const int nd = t / p.tail, cc = t % p.tail;
                    p.out[(node0 + nd) * (int64_t)p.ldo + p.C + cc] = p.tail_src[(node0 + nd) * (int64_t)p.tail_ld + cc];
                }
            tc_fence_before();
            asm volatile("bar.sync 3, 256;" ::: "memory");          // every worker is done with D1 / D2 before the next tile's MMAs
        }
    } else if (warp == EE_WORKER_WARPS) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc1 = TF32 ? make_idesc_tf32(TC_M, (uint32_t)p.n2, 0, 0) : make_idesc_bf16(TC_M, (uint32_t)p.n2, 0, 0);
            const uint32_t idesc2 = TF32 ? make_idesc_tf32(TC_M, (uint32_t)p.n3, 0, 0) : make_idesc_bf16(TC_M, (uint32_t)p.n3, 0, 0);
            const uint32_t lbo_a = TC_M * 16, sbo = 128;
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int phase = 0; phase < 2; ++phase) {
                    const int nkb = phase == 0 ? kb1 : kb2;
                    const uint32_t n_tile = (uint32_t)(phase == 0 ? p.n2 : p.n3), lbo_b = n_tile * 16;
                    const uint32_t idesc = phase == 0 ? idesc1 : idesc2;
                    const uint32_t dst = tmem_base + (phase == 0 ? 0u : d2_col);
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const uint32_t st = it % EE_NST;
                        mbar_wait(&full[st], (it / EE_NST) & 1u);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + (size_t)st * EE_STAGE_BYTES), a_lo = a_hi + EE_A_PLANE;
                        const uint32_t b_hi = a_lo + EE_A_PLANE, b_lo = b_hi + 4 * lbo_b;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t dah = make_smem_desc(a_hi + kk * 2 * lbo_a, lbo_a, sbo), dal = make_smem_desc(a_lo + kk * 2 * lbo_a, lbo_a, sbo);
                            const uint64_t dbh = make_smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, sbo), dbl = make_smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, sbo);
                            if constexpr (TF32) {
                                umma_tf32(dst, dah, dbh, idesc, (kb | kk) ? 1u : 0u);
                                umma_tf32(dst, dah, dbl, idesc, 1u);
                                umma_tf32(dst, dal, dbh, idesc, 1u);
                            } else {
                                umma_bf16(dst, dah, dbh, idesc, (kb | kk) ? 1u : 0u);
                                umma_bf16(dst, dah, dbl, idesc, 1u);
                                umma_bf16(dst, dal, dbh, idesc, 1u);
                            }
                        }
                        umma_commit(&empty[st]);
                    }
                    umma_commit(phase == 0 ? d1_full : d2_full);
                }
            }
        }
    } else {
        // =========================== weight loader: one cp.async.bulk per K block ===========================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int phase = 0; phase < 2; ++phase) {
                    const int nkb = phase == 0 ? kb1 : kb2;
                    const uint32_t bytes = (uint32_t)(phase == 0 ? p.n2 : p.n3) * 128u;
                    const uint8_t *w = phase == 0 ? p.w2 : p.w3;
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const uint32_t st = it % EE_NST;
                        mbar_wait(&empty[st], ((it / EE_NST) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(&full[st], bytes);
                        bulk_g2s(smem + (size_t)st * EE_STAGE_BYTES + 2 * EE_A_PLANE, w + (size_t)kb * bytes, bytes, &full[st]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EE_WORKER_WARPS + 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace nt

using namespace nt;

// 1 if nt_edgeconv_eval_fwd can run these sizes (the host falls back to the layer-by-layer kernels otherwise)
extern "C" int nt_edgeconv_eval_supported(int H1, int H2, int C, int k, int ldpq) {
    return H1 >= 4 && H1 <= 224 && (H1 % 4) == 0 && H2 >= 1 && H2 <= 224 && C >= 1 && C <= 160 && k >= 1 && k <= TC_M && (ldpq % 4) == 0 &&
           ldpq >= 2 * H1;
}

extern "C" int nt_edgeconv_eval_fwd(const float *pq, int ldpq, int H1, const int32_t *idx, int k, int n_per_cloud, int64_t M,
                                    const void *w2_split, const float *b2, int H2, const void *w3_split, const float *b3, int C,
                                    int precision, const float *s_out, const float *t_out, const float *tail_src, int tail_ld, int tail,
                                    float *out, int ldo, void *stream) {
    NT_REQUIRE(precision == NT_PREC_BF16X3 || precision == NT_PREC_TF32X3, "nt_edgeconv_eval_fwd: bad precision");
    NT_REQUIRE(pq && idx && w2_split && b2 && w3_split && b3 && s_out && t_out && out, "nt_edgeconv_eval_fwd: null argument");
    NT_REQUIRE(nt_edgeconv_eval_supported(H1, H2, C, k, ldpq), "nt_edgeconv_eval_fwd: unsupported sizes (see nt_edgeconv_eval_supported)");
    NT_REQUIRE(aligned16(pq) && M >= 1 && n_per_cloud >= 1 && ldo >= C + tail && (tail == 0 || tail_src), "nt_edgeconv_eval_fwd: bad arguments");
    EEParams p{};
    p.pq = pq; p.ldpq = ldpq; p.H1 = H1; p.idx = idx; p.k = k; p.n_per_cloud = n_per_cloud; p.M = M;
    p.w2 = reinterpret_cast<const uint8_t *>(w2_split); p.b2 = b2; p.H2 = H2; p.n2 = tc_geometry(H2, H1, precision).n_tile;
    p.w3 = reinterpret_cast<const uint8_t *>(w3_split); p.b3 = b3; p.C = C; p.n3 = tc_geometry(C, H2, precision).n_tile;
    p.s_out = s_out; p.t_out = t_out; p.tail_src = tail_src; p.tail_ld = tail_ld; p.tail = tail; p.out = out; p.ldo = ldo;
    p.rows_per_tile = (TC_M / k) * k;
    const int64_t nodes_per_tile = p.rows_per_tile / k;
    p.n_tiles = (M + nodes_per_tile - 1) / nodes_per_tile;
    static int sms = 0;
    const size_t smem = (size_t)EE_NST * EE_STAGE_BYTES + 128 + (224 + 3 * 160) * sizeof(float) + 2 * 128 * 33 * sizeof(float);
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1)
            return fail("nt_edgeconv_eval_fwd: cannot query the SM count%s", "");
        if (cudaFuncSetAttribute(edgeconv_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(edgeconv_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("nt_edgeconv_eval_fwd: cudaFuncSetAttribute failed%s", "");
        sms = n;
    }
    const int ctas = (int)(p.n_tiles < sms ? p.n_tiles : sms);
    if (precision == NT_PREC_TF32X3) edgeconv_eval_kernel<true><<<ctas, EE_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    else edgeconv_eval_kernel<false><<<ctas, EE_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("nt_edgeconv_eval_fwd");
}
