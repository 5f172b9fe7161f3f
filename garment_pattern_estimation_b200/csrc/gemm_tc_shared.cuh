// Helpers shared by the tensor-core row-GEMM kernels (gemm_tc.cu: one tile per CTA; gemm_tc3.cu: persistent streaming engine).
#pragma once
#include "gemm_params.cuh"
#include "tc_common.cuh"

namespace nt {
using namespace tc;

constexpr int TC_THREADS = 192;
constexpr int TC_M = 128;          // UMMA M
// a stage holds 4 chunks of 16 B along K: 32 bf16 elements (NT_PREC_BF16X3) or 16 tf32 elements (NT_PREC_TF32X3)
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = 4 * TC_M * 16;          // one of hi / lo: [4 chunks][128 rows][16 B]

struct TCGeom {
    int epc;             // elements per 16-byte chunk: 8 (bf16) or 4 (tf32)
    int n_tile;          // columns per CTA (multiple of 16, <= 256)
    int n_tiles;         // column tiles
    int num_kb;          // K blocks
    int tmem_cols;       // pow2 >= 32 allocation
};

__host__ __device__ inline TCGeom tc_geometry(int n_out, int K, int precision) {
    TCGeom g;
    g.epc = precision == NT_PREC_TF32X3 ? 4 : 8;
    g.n_tiles = (n_out + 255) / 256;
    int per = (n_out + g.n_tiles - 1) / g.n_tiles;
    g.n_tile = ((per + 15) / 16) * 16;
    g.num_kb = (K + 4 * g.epc - 1) / (4 * g.epc);
    int c = 32;
    while (c < g.n_tile) c <<= 1;
    g.tmem_cols = c;
    return g;
}
__host__ __device__ inline size_t tc_stage_bytes(int n_tile) { return 2 * TC_A_BYTES + (size_t)n_tile * 128; }

// one 16-byte chunk (8 bf16 or 4 tf32 values) of the hi and lo planes
__device__ __forceinline__ void pack_chunk(const float (&v)[8], bool tf32, uint4 &h, uint4 &l) {
    if (tf32) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
        h = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        l = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        h = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        l = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
template <int EPC>
__device__ __forceinline__ void load_chunk(const float *src, int k, int K, bool vec, float (&v)[8]) {
    if (vec && k + EPC - 1 < K) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(src + k));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        if (EPC == 8) {
            float4 b = __ldg(reinterpret_cast<const float4 *>(src + k + 4));
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPC; ++e) v[e] = (k + e < K) ? __ldg(src + k + e) : 0.f;
    }
}


// streaming engine entry (gemm_tc3.cu); returns -1 when the call is not eligible (caller falls back to gemm_tc.cu)
int launch_nt_tc3(const NTParams &p, int producer, int epilogue, const void *w_split, cudaStream_t st);
bool tc3_eligible(const NTParams &p, int producer, int epilogue);
// second-generation streaming engine (gemm_tc4.cu); -1 when the call is not eligible
int launch_nt_tc4(const NTParams &p, int producer, int epilogue, const void *w_split, cudaStream_t st);
bool tc4_eligible(const NTParams &p, int producer, int epilogue);

}  // namespace nt
