// The non-network pieces of the reference training step (nn/trainer.py:96-101) as two kernels each:
//
//   * nt_pattern_loss_fwd / _bwd -- the four loss terms active in the shipped attention config (models/att/att.yaml:124;
//     nn/metrics/composed_loss.py:294-321, nn/metrics/losses.py:19-51): MSE on outlines / rotations / translations + PanelLoopLoss.
//     The reference spends B*23 Python iterations with a device sync each on the loop term and ~25 element-wise launches on the
//     rest; the tensors are a few hundred KB.
//   * nt_adam_step -- torch.optim.Adam (nn/trainer.py:64, 98-99) on ONE flat parameter / gradient buffer: gradient scaling (the
//     1 / world_size of the data-parallel average), optional L2 weight decay, moment updates, bias-corrected parameter update and
//     the zeroing of the gradient buffer (optimizer.zero_grad) in a single pass.  The learning rate and the step counter live in
//     device memory, so the kernel is CUDA-graph safe under a stepping scheduler (OneCycleLR, nn/trainer.py:73-80).
#include "common.cuh"

namespace nt {

// ---------------------------------------------------------------------------------------------------------------------
// pattern loss
// ---------------------------------------------------------------------------------------------------------------------
struct LossArgs {
    const float *outl; int64_t so_b, so_p, so_e;        // predicted outlines [B, P, Lp, D], element strides (last dim dense)
    const float *rot; int64_t sr_b, sr_p;               // predicted rotations [B, P, Dr]
    const float *tr; int64_t st_b, st_p;                // predicted translations [B, P, Dt]
    const float *gt_outl, *gt_rot, *gt_tr;              // ground truth, contiguous
    const int64_t *num_edges;                           // [B, P]
    int B, P, Lp, D, Dr, Dt;
    float pad_x, pad_y, loop_weight;
    int use_shape, use_loop, use_rot, use_tr;
};

// result[0..4] = total, pattern_loss, loop_loss, rotation_loss, translation_loss  (double accumulators, zeroed by the caller)
__global__ void pattern_loss_fwd_kernel(LossArgs a, double *__restrict__ acc) {
    const int panels = a.B * a.P;
    double s_shape = 0.0, s_loop = 0.0, s_rot = 0.0, s_tr = 0.0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < panels; q += gridDim.x * blockDim.x) {
        const int b = q / a.P, p = q % a.P;
        const float *o = a.outl + b * a.so_b + p * a.so_p;
        const float *g = a.gt_outl + (int64_t)q * a.Lp * a.D;
        const int n = (int)a.num_edges[q];
        float sx = 0.f, sy = 0.f, sq = 0.f;
        for (int e = 0; e < a.Lp; ++e) {
            for (int c = 0; c < a.D; ++c) {
                const float v = o[e * a.so_e + c], d = v - g[e * a.D + c];
                sq = fmaf(d, d, sq);
            }
            if (e < n && n >= 3) { sx += o[e * a.so_e] - a.pad_x; sy += o[e * a.so_e + 1] - a.pad_y; }
        }
        s_shape += sq;
        s_loop += (double)sx * sx + (double)sy * sy;
        const float *r = a.rot + b * a.sr_b + p * a.sr_p, *gr = a.gt_rot + (int64_t)q * a.Dr;
        for (int c = 0; c < a.Dr; ++c) { const float d = r[c] - gr[c]; s_rot += (double)d * d; }
        const float *t = a.tr + b * a.st_b + p * a.st_p, *gtr = a.gt_tr + (int64_t)q * a.Dt;
        for (int c = 0; c < a.Dt; ++c) { const float d = t[c] - gtr[c]; s_tr += (double)d * d; }
    }
    // block reduction (4 values)
    __shared__ double red[4][32];
    double v[4] = {s_shape, s_loop, s_rot, s_tr};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
        if (lane == 0) red[i][warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
        const double denom[4] = {(double)panels * a.Lp * a.D, (double)panels * 2.0, (double)panels * a.Dr, (double)panels * a.Dt};
        const int use[4] = {a.use_shape, a.use_loop, a.use_rot, a.use_tr};
        if (use[threadIdx.x]) {
            const double part = t / denom[threadIdx.x];
            atomicAdd(acc + 1 + threadIdx.x, part);
            atomicAdd(acc, threadIdx.x == 1 ? part * (double)a.loop_weight : part);
        }
    }
}

__global__ void pattern_loss_finish_kernel(const double *__restrict__ acc, float *__restrict__ out) {
    if (threadIdx.x < 5) out[threadIdx.x] = (float)acc[threadIdx.x];
}

// gradients w.r.t. the three predictions (contiguous outputs), scaled by the upstream scalar *gscale
__global__ void pattern_loss_bwd_kernel(LossArgs a, const float *__restrict__ gscale, float *__restrict__ g_outl,
                                        float *__restrict__ g_rot, float *__restrict__ g_tr) {
    const int panels = a.B * a.P;
    const float gs = *gscale;
    const float k_shape = a.use_shape ? gs * 2.f / ((float)panels * a.Lp * a.D) : 0.f;
    const float k_loop = a.use_loop ? gs * a.loop_weight * 2.f / ((float)panels * 2.f) : 0.f;
    const float k_rot = a.use_rot ? gs * 2.f / ((float)panels * a.Dr) : 0.f;
    const float k_tr = a.use_tr ? gs * 2.f / ((float)panels * a.Dt) : 0.f;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < panels; q += gridDim.x * blockDim.x) {
        const int b = q / a.P, p = q % a.P;
        const float *o = a.outl + b * a.so_b + p * a.so_p;
        const float *g = a.gt_outl + (int64_t)q * a.Lp * a.D;
        const int n = (int)a.num_edges[q];
        float sx = 0.f, sy = 0.f;
        if (n >= 3)
            for (int e = 0; e < a.Lp && e < n; ++e) { sx += o[e * a.so_e] - a.pad_x; sy += o[e * a.so_e + 1] - a.pad_y; }
        float *go = g_outl + (int64_t)q * a.Lp * a.D;
        for (int e = 0; e < a.Lp; ++e) {
            const bool live = e < n && n >= 3;
            for (int c = 0; c < a.D; ++c) {
                float v = k_shape * (o[e * a.so_e + c] - g[e * a.D + c]);
                if (live && c == 0) v = fmaf(k_loop, sx, v);
                if (live && c == 1) v = fmaf(k_loop, sy, v);
                go[e * a.D + c] = v;
            }
        }
        const float *r = a.rot + b * a.sr_b + p * a.sr_p, *gr = a.gt_rot + (int64_t)q * a.Dr;
        for (int c = 0; c < a.Dr; ++c) g_rot[(int64_t)q * a.Dr + c] = k_rot * (r[c] - gr[c]);
        const float *t = a.tr + b * a.st_b + p * a.st_p, *gtr = a.gt_tr + (int64_t)q * a.Dt;
        for (int c = 0; c < a.Dt; ++c) g_tr[(int64_t)q * a.Dt + c] = k_tr * (t[c] - gtr[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Adam on a flat buffer
// ---------------------------------------------------------------------------------------------------------------------
// state: [0] = step count (float, incremented by the last block to finish), [1] = block counter (as uint32 bits)
__global__ void adam_step_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n,
                                 const float *__restrict__ lr_ptr, float beta1, float beta2, float eps, float weight_decay,
                                 float grad_scale, int zero_grad, float *__restrict__ state) {
    const float step = state[0] + 1.f;                       // torch counts from 1
    const float lr = *lr_ptr;
    const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], gg = reinterpret_cast<float4 *>(g)[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        float *pa = &pp.x, *ga = &gg.x, *ma = &mm.x, *va = &vv.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float gr = ga[e] * grad_scale;
            if (weight_decay != 0.f) gr = fmaf(weight_decay, pa[e], gr);
            ma[e] = fmaf(beta1, ma[e], (1.f - beta1) * gr);            // torch: exp_avg.lerp_(grad, 1 - beta1)
            va[e] = fmaf(beta2, va[e], (1.f - beta2) * gr * gr);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(va[e]) * inv_sqrt_bc2 + eps;
            pa[e] -= step_size * (ma[e] / denom);
        }
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
        if (zero_grad) reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {              // tail (n not a multiple of 4)
        const int64_t i = (n4 << 2) + threadIdx.x;
        float gr = g[i] * grad_scale;
        if (weight_decay != 0.f) gr = fmaf(weight_decay, p[i], gr);
        m[i] = fmaf(beta1, m[i], (1.f - beta1) * gr);
        v[i] = fmaf(beta2, v[i], (1.f - beta2) * gr * gr);
        p[i] -= step_size * (m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps));
        if (zero_grad) g[i] = 0.f;
    }
    // the last block to finish advances the step counter for the NEXT call (every block of this call has read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *counter = reinterpret_cast<unsigned int *>(state + 1);
        __threadfence();
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {
            state[0] = step;
            *counter = 0u;
        }
    }
}

}  // namespace nt

using namespace nt;

extern "C" int nt_pattern_loss_fwd(const nt_pattern_loss_args *args, double *acc5, float *out5, void *stream) {
    NT_REQUIRE(args && acc5 && out5, "nt_pattern_loss_fwd: null argument");
    NT_REQUIRE(args->outlines && args->rotations && args->translations && args->gt_outlines && args->gt_rotations &&
                   args->gt_translations && args->num_edges, "nt_pattern_loss_fwd: null tensor");
    NT_REQUIRE(args->B >= 1 && args->P >= 1 && args->Lp >= 1 && args->D >= 2, "nt_pattern_loss_fwd: bad sizes");
    LossArgs a;
    a.outl = args->outlines; a.so_b = args->outl_stride_b; a.so_p = args->outl_stride_p; a.so_e = args->outl_stride_e;
    a.rot = args->rotations; a.sr_b = args->rot_stride_b; a.sr_p = args->rot_stride_p;
    a.tr = args->translations; a.st_b = args->tr_stride_b; a.st_p = args->tr_stride_p;
    a.gt_outl = args->gt_outlines; a.gt_rot = args->gt_rotations; a.gt_tr = args->gt_translations; a.num_edges = args->num_edges;
    a.B = args->B; a.P = args->P; a.Lp = args->Lp; a.D = args->D; a.Dr = args->Dr; a.Dt = args->Dt;
    a.pad_x = args->pad_x; a.pad_y = args->pad_y; a.loop_weight = args->loop_weight;
    a.use_shape = args->use_shape; a.use_loop = args->use_loop; a.use_rot = args->use_rotation; a.use_tr = args->use_translation;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(acc5, 0, 5 * sizeof(double), st) != cudaSuccess) return fail("nt_pattern_loss_fwd: cudaMemsetAsync failed%s", "");
    const int panels = a.B * a.P;
    pattern_loss_fwd_kernel<<<(panels + 127) / 128, 128, 0, st>>>(a, acc5);
    if (int rc = check_launch("nt_pattern_loss_fwd")) return rc;
    pattern_loss_finish_kernel<<<1, 32, 0, st>>>(acc5, out5);
    return check_launch("nt_pattern_loss_fwd(finish)");
}

extern "C" int nt_pattern_loss_bwd(const nt_pattern_loss_args *args, const float *grad_scale, float *g_outlines, float *g_rotations,
                                   float *g_translations, void *stream) {
    NT_REQUIRE(args && grad_scale && g_outlines && g_rotations && g_translations, "nt_pattern_loss_bwd: null argument");
    LossArgs a;
    a.outl = args->outlines; a.so_b = args->outl_stride_b; a.so_p = args->outl_stride_p; a.so_e = args->outl_stride_e;
    a.rot = args->rotations; a.sr_b = args->rot_stride_b; a.sr_p = args->rot_stride_p;
    a.tr = args->translations; a.st_b = args->tr_stride_b; a.st_p = args->tr_stride_p;
    a.gt_outl = args->gt_outlines; a.gt_rot = args->gt_rotations; a.gt_tr = args->gt_translations; a.num_edges = args->num_edges;
    a.B = args->B; a.P = args->P; a.Lp = args->Lp; a.D = args->D; a.Dr = args->Dr; a.Dt = args->Dt;
    a.pad_x = args->pad_x; a.pad_y = args->pad_y; a.loop_weight = args->loop_weight;
    a.use_shape = args->use_shape; a.use_loop = args->use_loop; a.use_rot = args->use_rotation; a.use_tr = args->use_translation;
    const int panels = a.B * a.P;
    pattern_loss_bwd_kernel<<<(panels + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, grad_scale, g_outlines, g_rotations,
                                                                                                   g_translations);
    return check_launch("nt_pattern_loss_bwd");
}

extern "C" int nt_adam_step(float *params, float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, const float *lr, float beta1,
                            float beta2, float eps, float weight_decay, float grad_scale, int zero_grad, float *state2, void *stream) {
    NT_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr && state2 && n >= 1, "nt_adam_step: null argument");
    NT_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq), "nt_adam_step: buffers must be 16-byte aligned");
    const int64_t n4 = (n + 3) / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                weight_decay, grad_scale, zero_grad, state2);
    return check_launch("nt_adam_step");
}
