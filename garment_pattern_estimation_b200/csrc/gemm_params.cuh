// Parameter blocks shared by the CUDA-core (gemm.cu) and tensor-core (gemm_tc.cu) GEMM engines.
#pragma once
#include "common.cuh"

namespace nt {

struct EdgeSrc {
    const float *pq; int ldpq; int qoff; const int32_t *idx; int k; int n_per_cloud;
};

// value of the EDGE producer at (row r, column c)
__device__ __forceinline__ void edge_row_ptrs(const EdgeSrc &s, int64_t r, const float *&p, const float *&q) {
    if (s.idx) {
        int64_t centre = r / s.k;
        int64_t base = (centre / s.n_per_cloud) * (int64_t)s.n_per_cloud;
        int64_t j = base + s.idx[r];
        p = s.pq + centre * s.ldpq;
        q = s.pq + j * s.ldpq + s.qoff;
    } else {
        p = s.pq + r * s.ldpq;
        q = nullptr;
    }
}

struct NTParams {
    int64_t rows; int K; int n_out; int rows_per_tile;
    const float *a; int lda;
    EdgeSrc e;
    const float *w; int ldw; const float *bias;
    float *out; int ldo;
    double *stats;
    float *vmax, *vmin; uint8_t *imax, *imin; int k_agg;
    const float *aux; int ldaux; int aux_edge; EdgeSrc ae;
    const float *k0, *k1, *mu; double *colsum;
    float *scatter; int ldscatter;               // fused edge scatter (BNRELU_BWD on the streaming engine only)
    int engine;                                  // nt_gemm_args.engine: 0 = auto, 1 = one tile per CTA, 3 / 4 / 5 = streaming configurations
};


// tensor-core engine entry (gemm_tc.cu); w_split = weights pre-split by nt_gemm_prepare_weights
int launch_nt_tc(const NTParams &p, int producer, int epilogue, int precision, const void *w_split, cudaStream_t st);
// true if launch_nt_tc would hand this call to the streaming engine (gemm_tc3.cu)
bool nt_tc_would_stream(const NTParams &p, int producer, int epilogue, int precision);

// tensor-core weight-gradient engine (gemm_tn_tc.cu)
int gemm_tn_tc(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows, const EdgeSrc &e, int b_edge,
               const float *mu, void *out, int out_double, int ldo, float *workspace, cudaStream_t st);
// the same product for plain (not gathered) operands through MN-major BF16x3 tensor-core operands (gemm_tn_mn.cu)
int gemm_tn_mn(const float *a, int lda, int m, const float *b, int ldb, int n, int64_t rows, const float *mu, void *out, int out_double,
               int ldo, float *workspace, cudaStream_t st);

}  // namespace nt
