// Host emulation of the persistent LSTM kernels (csrc/lstm.cu): runs THE SAME layout / cell functions (csrc/lstm_layout.cuh)
// on the CPU, emulating only (a) the tensor-core product -- operands are read back from the byte images through the UMMA
// K-major no-swizzle descriptor addressing (start + (k/8)*LBO + (row/8)*SBO + (row%8)*16 + (k%8)*2), three bf16 products per
// K-step -- and (b) the order in which the cells become ready.  Compared against a plain double-precision LSTM forward /
// backward (the decomposition of oracle/lstm_decomposed.py).  No GPU needed:
//
//     g++ -O2 -std=c++17 -o /tmp/lstm_emulate tools/lstm_emulate.cpp && /tmp/lstm_emulate [R T L H E]
//
// Exit code 0 = every compared tensor within tolerance.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../garment_pattern_estimation_b200/csrc/lstm_layout.cuh"

using namespace nt::lstm;

static float bf(const uint8_t *p) { uint16_t b; memcpy(&b, p, 2); return bf16_bits_to_f32(b); }
// element (row, k) of a K-major no-swizzle operand plane as the tensor core addresses it
static float umma_elem(const uint8_t *plane, int row, int k, int lbo, int sbo) {
    return bf(plane + (k / 8) * lbo + (row / 8) * sbo + (row % 8) * 16 + (k % 8) * 2);
}
// D[row][n] += A[row][0..15] . B[n][0..15] over one K-step, bf16x3
static void mma_kstep(const uint8_t *a_stage_sub, const uint8_t *b_kstep, int n_cols, int b_plane_bytes, int b_lbo, float *D, int ldd) {
    for (int r = 0; r < TILE_M; ++r)
        for (int n = 0; n < n_cols; ++n) {
            float acc = 0.f;
            for (int k = 0; k < KSTEP; ++k) {
                const float ah = umma_elem(a_stage_sub, r, k, CHUNK_BYTES, 128), al = umma_elem(a_stage_sub + PLANE_BYTES, r, k, CHUNK_BYTES, 128);
                const float bh = umma_elem(b_kstep, n, k, b_lbo, 128), bl = umma_elem(b_kstep + b_plane_bytes, n, k, b_lbo, 128);
                acc += ah * bh + ah * bl + al * bh;
            }
            D[r * ldd + n] += acc;
        }
}

struct Net {
    int R, T, L, H, E;
    std::vector<std::vector<double>> w_ih, w_hh, b_ih, b_hh;
    std::vector<double> x, h0, c0;
};
static double sig(double v) { return 1.0 / (1.0 + exp(-v)); }

int main(int argc, char **argv) {
    Net n;
    n.R = argc > 1 ? atoi(argv[1]) : 150; n.T = argc > 2 ? atoi(argv[2]) : 3; n.L = argc > 3 ? atoi(argv[3]) : 2;
    n.H = argc > 4 ? atoi(argv[4]) : 250; n.E = argc > 5 ? atoi(argv[5]) : 250;
    const int nsub_arg = argc > 6 ? atoi(argv[6]) : 1;
    const int R = n.R, T = n.T, L = n.L, H = n.H, E = n.E;
    std::mt19937 rng(7);
    std::normal_distribution<double> nd(0.0, 1.0);
    n.w_ih.resize(L); n.w_hh.resize(L); n.b_ih.resize(L); n.b_hh.resize(L);
    for (int l = 0; l < L; ++l) {
        const int in = l == 0 ? E : H;
        n.w_ih[l].resize((size_t)4 * H * in); n.w_hh[l].resize((size_t)4 * H * H); n.b_ih[l].resize(4 * H); n.b_hh[l].resize(4 * H);
        for (auto &v : n.w_ih[l]) v = nd(rng) * sqrt(2.0 / in);
        for (auto &v : n.w_hh[l]) v = nd(rng) * sqrt(2.0 / H);
        for (auto &v : n.b_ih[l]) v = nd(rng) * 0.1;
        for (auto &v : n.b_hh[l]) v = nd(rng) * 0.1;
    }
    n.x.resize((size_t)R * E); n.h0.resize((size_t)L * R * H); n.c0.resize((size_t)L * R * H);
    for (auto &v : n.x) v = nd(rng);
    for (auto &v : n.h0) v = nd(rng) * 0.3;
    for (auto &v : n.c0) v = nd(rng) * 0.3;
    std::vector<double> dy((size_t)T * R * H);
    for (auto &v : dy) v = nd(rng);

    // ------------------------------------------------------------------ reference (double) ------------------------------------------
    // states[l][slot][r][H]; gates act[l][t][r][4H]
    std::vector<double> Hs((size_t)L * (T + 1) * R * H), Cs((size_t)L * (T + 1) * R * H), act((size_t)L * T * R * 4 * H);
    auto HS = [&](int l, int s, int r, int u) -> double & { return Hs[(((size_t)l * (T + 1) + s) * R + r) * H + u]; };
    auto CS = [&](int l, int s, int r, int u) -> double & { return Cs[(((size_t)l * (T + 1) + s) * R + r) * H + u]; };
    auto ACT = [&](int l, int t, int r, int q) -> double & { return act[(((size_t)l * T + t) * R + r) * 4 * H + q]; };
    for (int l = 0; l < L; ++l)
        for (int r = 0; r < R; ++r)
            for (int u = 0; u < H; ++u) { HS(l, 0, r, u) = n.h0[((size_t)l * R + r) * H + u]; CS(l, 0, r, u) = n.c0[((size_t)l * R + r) * H + u]; }
    for (int l = 0; l < L; ++l) {
        const int in = l == 0 ? E : H;
        for (int t = 0; t < T; ++t)
            for (int r = 0; r < R; ++r) {
                std::vector<double> G(4 * H);
                for (int q = 0; q < 4 * H; ++q) {
                    double a = n.b_ih[l][q] + n.b_hh[l][q];
                    for (int k = 0; k < in; ++k) a += n.w_ih[l][(size_t)q * in + k] * (l == 0 ? n.x[(size_t)r * E + k] : HS(l - 1, t + 1, r, k));
                    for (int k = 0; k < H; ++k) a += n.w_hh[l][(size_t)q * H + k] * HS(l, t, r, k);
                    G[q] = a;
                }
                for (int u = 0; u < H; ++u) {
                    const double i = sig(G[u]), f = sig(G[H + u]), g = tanh(G[2 * H + u]), o = sig(G[3 * H + u]);
                    const double c = f * CS(l, t, r, u) + i * g;
                    CS(l, t + 1, r, u) = c; HS(l, t + 1, r, u) = o * tanh(c);
                    ACT(l, t, r, u) = i; ACT(l, t, r, H + u) = f; ACT(l, t, r, 2 * H + u) = g; ACT(l, t, r, 3 * H + u) = o;
                }
            }
    }
    // backward reference
    std::vector<std::vector<double>> rdw_ih(L), rdw_hh(L), rdb(L);
    std::vector<double> rdx((size_t)R * E, 0.0);
    {
        std::vector<double> dH(dy);                                   // [T][R][H] gradient w.r.t. the outputs of the current layer
        for (int l = L - 1; l >= 0; --l) {
            const int in = l == 0 ? E : H;
            rdw_ih[l].assign((size_t)4 * H * in, 0.0); rdw_hh[l].assign((size_t)4 * H * H, 0.0); rdb[l].assign(4 * H, 0.0);
            std::vector<double> dHbelow((size_t)T * R * H, 0.0), dhrec((size_t)R * H, 0.0), dcrec((size_t)R * H, 0.0);
            for (int t = T - 1; t >= 0; --t)
                for (int r = 0; r < R; ++r) {
                    std::vector<double> dG(4 * H);
                    for (int u = 0; u < H; ++u) {
                        const double i = ACT(l, t, r, u), f = ACT(l, t, r, H + u), g = ACT(l, t, r, 2 * H + u), o = ACT(l, t, r, 3 * H + u);
                        const double tc = tanh(CS(l, t + 1, r, u));
                        const double dh = dH[((size_t)t * R + r) * H + u] + dhrec[(size_t)r * H + u];
                        const double dc = dcrec[(size_t)r * H + u] + dh * o * (1 - tc * tc);
                        dG[u] = dc * g * i * (1 - i); dG[H + u] = dc * CS(l, t, r, u) * f * (1 - f);
                        dG[2 * H + u] = dc * i * (1 - g * g); dG[3 * H + u] = dh * tc * o * (1 - o);
                        dcrec[(size_t)r * H + u] = dc * f;
                    }
                    for (int u = 0; u < H; ++u) dhrec[(size_t)r * H + u] = 0.0;
                    for (int q = 0; q < 4 * H; ++q) {
                        const double gq = dG[q];
                        rdb[l][q] += gq;
                        for (int k = 0; k < H; ++k) {
                            dhrec[(size_t)r * H + k] += gq * n.w_hh[l][(size_t)q * H + k];
                            rdw_hh[l][(size_t)q * H + k] += gq * HS(l, t, r, k);
                        }
                        for (int k = 0; k < in; ++k) {
                            const double xin = l == 0 ? n.x[(size_t)r * E + k] : HS(l - 1, t + 1, r, k);
                            rdw_ih[l][(size_t)q * in + k] += gq * xin;
                            if (l == 0) rdx[(size_t)r * E + k] += gq * n.w_ih[l][(size_t)q * in + k];
                            else dHbelow[((size_t)t * R + r) * H + k] += gq * n.w_ih[l][(size_t)q * in + k];
                        }
                    }
                }
            dH.swap(dHbelow);
        }
    }

    // ------------------------------------------------------------------ emulated kernels --------------------------------------------
    Dims d;
    d.R = R; d.T = T; d.L = L; d.H = H; d.E = E; d.nsub = nsub_arg;
    d.RT = (R + d.tile_rows() - 1) / d.tile_rows();
    // float copies of the parameters (what the kernels see)
    std::vector<std::vector<float>> fw_ih(L), fw_hh(L), fb_ih(L), fb_hh(L);
    for (int l = 0; l < L; ++l) {
        fw_ih[l].assign(n.w_ih[l].begin(), n.w_ih[l].end()); fw_hh[l].assign(n.w_hh[l].begin(), n.w_hh[l].end());
        fb_ih[l].assign(n.b_ih[l].begin(), n.b_ih[l].end()); fb_hh[l].assign(n.b_hh[l].begin(), n.b_hh[l].end());
    }
    std::vector<float> fx(n.x.begin(), n.x.end()), fh0(n.h0.begin(), n.h0.end()), fc0(n.c0.begin(), n.c0.end()), fdy(dy.begin(), dy.end());
    // weight preparation (what lstm_prepare_weights_kernel writes)
    std::vector<uint8_t> wf((size_t)L * SLICES * FWD_W_BYTES), wb((size_t)L * SLICES * BWD_W_BYTES);
    std::vector<float> bias((size_t)L * SLICES * FWD_N);
    for (int l = 0; l < L; ++l)
        for (int c = 0; c < SLICES; ++c) {
            const int in = l == 0 ? E : H;
            for (int ks = 0; ks < FWD_KSTEPS; ++ks)
                for (int nn = 0; nn < FWD_N; ++nn)
                    for (int kk = 0; kk < KSTEP; ++kk) {
                        uint16_t hi, lo;
                        split_hi_lo(fwd_w_value(fw_ih[l].data(), fw_hh[l].data(), in, H, c, ks, nn, kk), hi, lo);
                        memcpy(&wf[fwd_w_offset(l, c, ks, 0, nn, kk)], &hi, 2); memcpy(&wf[fwd_w_offset(l, c, ks, 1, nn, kk)], &lo, 2);
                    }
            for (int ks = 0; ks < BWD_KSTEPS; ++ks)
                for (int nn = 0; nn < BWD_N; ++nn)
                    for (int kk = 0; kk < KSTEP; ++kk) {
                        uint16_t hi, lo;
                        split_hi_lo(bwd_w_value(fw_ih[l].data(), fw_hh[l].data(), in, H, c, ks, nn, kk), hi, lo);
                        memcpy(&wb[bwd_w_offset(l, c, ks, 0, nn, kk)], &hi, 2); memcpy(&wb[bwd_w_offset(l, c, ks, 1, nn, kk)], &lo, 2);
                    }
            for (int nn = 0; nn < FWD_N; ++nn) {
                const int g = nn >> 4, unit = UNITS * c + (nn & 15);
                bias[((size_t)l * SLICES + c) * FWD_N + nn] = unit < H ? fb_ih[l][g * H + unit] + fb_hh[l][g * H + unit] : 0.f;
            }
        }
    const int TR = d.tile_rows();
    std::vector<uint8_t> actbuf((size_t)(L + 1) * (T + 1) * d.RT * d.act_block_bytes(), 0xFF);       // 0xFF: unwritten bytes show up as NaN
    std::vector<float> y((size_t)T * R * HP, NAN), cs((size_t)L * (T + 1) * d.RT * SLICES * UNITS * TR, NAN),
        gates((size_t)L * T * d.RT * SLICES * 4 * UNITS * TR, NAN);
    FwdOut fo{actbuf.data(), y.data(), cs.data(), gates.data()};
    std::vector<float> cstate((size_t)L * d.RT * SLICES * TR * UNITS);
    auto CST = [&](int l, int rt, int c, int row) { return &cstate[((((size_t)l * d.RT + rt) * SLICES + c) * TR + row) * UNITS]; };
    // step -1
    for (int l = 0; l < L; ++l)
        for (int rt = 0; rt < d.RT; ++rt)
            for (int c = 0; c < SLICES; ++c)
                for (int row = 0; row < TR; ++row) {
                    const int64_t grow = (int64_t)rt * TR + row;
                    float h[UNITS], xv[UNITS], *cst = CST(l, rt, c, row);
                    for (int u = 0; u < UNITS; ++u) {
                        const int unit = UNITS * c + u;
                        const bool live = grow < R && unit < H;
                        cst[u] = live ? fc0[((size_t)l * R + grow) * H + unit] : 0.f;
                        h[u] = live ? fh0[((size_t)l * R + grow) * H + unit] : (unit == H ? 1.f : 0.f);
                        xv[u] = (grow < R && unit < E) ? fx[(size_t)grow * E + unit] : 0.f;
                    }
                    float (&hh)[UNITS] = *reinterpret_cast<float (*)[UNITS]>(h);
                    float (&cc)[UNITS] = *reinterpret_cast<float (*)[UNITS]>(cst);
                    float (&xx)[UNITS] = *reinterpret_cast<float (*)[UNITS]>(xv);
                    store_split16(actbuf.data() + act_block_index(d, l + 1, 0, rt) * d.act_block_bytes(), d, c, row, hh);
                    if (l == 0) store_split16(actbuf.data() + act_block_index(d, 0, 0, rt) * d.act_block_bytes(), d, c, row, xx);
                    store16_rowinner(cs.data() + cs_offset(d, l, 0, rt, c), TR, row, cc);
                }
    std::vector<float> D((size_t)TR * FWD_N);
    for (int t = 0; t < T; ++t)
        for (int l = 0; l < L; ++l)
            for (int rt = 0; rt < d.RT; ++rt)
                for (int c = 0; c < SLICES; ++c) {
                    std::fill(D.begin(), D.end(), 0.f);
                    for (int part = 0; part < 2; ++part) {
                        const int src = part == 0 ? l : l + 1, slot = part == 0 ? (l == 0 ? 0 : t + 1) : t;
                        const uint8_t *blk = actbuf.data() + act_block_index(d, src, slot, rt) * d.act_block_bytes();
                        for (int j = 0; j < SLICES; ++j) {
                            const int cc = (c + j) & (SLICES - 1);
                            const uint8_t *stage = blk + (size_t)cc * d.stage_bytes();
                            const uint8_t *b = wf.data() + ((size_t)l * SLICES + c) * FWD_W_BYTES + (size_t)(part * SLICES + cc) * FWD_W_KSTEP_BYTES;
                            for (int sub = 0; sub < d.nsub; ++sub)
                                mma_kstep(stage + (size_t)sub * SUB_BYTES, b, FWD_N, 2048, 1024, &D[(size_t)sub * TILE_M * FWD_N], FWD_N);
                        }
                    }
                    for (int row = 0; row < TR; ++row) {
                        float h[UNITS], gsave[4][UNITS];
                        float (&acc)[FWD_N] = *reinterpret_cast<float (*)[FWD_N]>(&D[(size_t)row * FWD_N]);
                        float (&cc)[UNITS] = *reinterpret_cast<float (*)[UNITS]>(CST(l, rt, c, row));
                        fwd_cell_row(d, fo, l, t, rt, c, row, acc, &bias[((size_t)l * SLICES + c) * FWD_N], cc, h, gsave);
                        if (l == L - 1) fwd_store_y(d, y.data(), t, rt, c, row, h);
                        fwd_store_saved(d, fo, l, t, rt, c, row, gsave, cc);
                    }
                }
    int bad = 0;
    auto report = [&](const char *what, double err, double tol) {
        printf("%-28s max rel err %.3e (tol %.1e) %s\n", what, err, tol, err <= tol ? "ok" : "FAIL");
        if (!(err <= tol)) ++bad;
    };
    // h of every layer / step, read back from the split blocks (hi + lo), and the fp32 output of the top layer
    auto act_h = [&](int src, int slot, int64_t r, int unit) {
        const int rt = (int)(r / TR), row = (int)(r % TR);
        const uint8_t *blk = actbuf.data() + act_block_index(d, src, slot, rt) * d.act_block_bytes();
        return bf(blk + split_offset(d, unit / 16, row, unit % 16, 0)) + bf(blk + split_offset(d, unit / 16, row, unit % 16, 1));
    };
    {
        double err = 0, scale = 0, ones_bad = 0, erry = 0;
        for (int l = 0; l < L; ++l)
            for (int sl = 0; sl <= T; ++sl)
                for (int r = 0; r < R; ++r) {
                    for (int u = 0; u < H; ++u) {
                        const double want = HS(l, sl, r, u), got = act_h(l + 1, sl, r, u);
                        err = fmax(err, std::isnan(got) ? 1e30 : fabs(want - got)); scale = fmax(scale, fabs(want));
                        if (l == L - 1 && sl > 0) { const double gy = y[y_offset(d, sl - 1, r) + u]; erry = fmax(erry, std::isnan(gy) ? 1e30 : fabs(want - gy)); }
                    }
                    if (act_h(l + 1, sl, r, H) != 1.f) ones_bad = 1;
                }
        report("forward h (all layers)", err / scale, 5e-5);
        report("forward y (top layer)", erry / scale, 5e-5);
        report("ones column", ones_bad, 0.0);
    }

    // ---- backward
    std::vector<uint8_t> dgs((size_t)L * T * d.RT * d.dg_block_bytes(), 0xFF);
    std::vector<float> dxbuf((size_t)L * T * d.RT * SLICES * TR * UNITS, NAN), dx0((size_t)R * E, NAN);
    BwdIo io{cs.data(), gates.data(), fdy.data(), H, dgs.data(), dxbuf.data(), dx0.data(), E};
    std::vector<float> dcst((size_t)L * d.RT * SLICES * TR * UNITS, 0.f), dhrec(dcst.size(), 0.f), dxsum(dcst.size(), 0.f);
    auto ST = [&](std::vector<float> &v, int l, int rt, int c, int row) { return &v[((((size_t)l * d.RT + rt) * SLICES + c) * TR + row) * UNITS]; };
    std::vector<float> D2((size_t)TR * BWD_N);
    for (int t = T - 1; t >= 0; --t)
        for (int l = L - 1; l >= 0; --l) {
            for (int rt = 0; rt < d.RT; ++rt)
                for (int c = 0; c < SLICES; ++c)
                    for (int row = 0; row < TR; ++row) {
                        const int64_t grow = (int64_t)rt * TR + row;
                        float dh[UNITS], dgo[4][UNITS];
                        if (l == L - 1) {
                            for (int u = 0; u < UNITS; ++u) {
                                const int unit = UNITS * c + u;
                                dh[u] = (grow < R && unit < H) ? fdy[((size_t)t * R + grow) * H + unit] : 0.f;
                            }
                        } else {
                            load16_rowinner(dxbuf.data() + dx_block_offset(d, l + 1, t, rt, c), TR, row, dh);
                        }
                        float *rec = ST(dhrec, l, rt, c, row);
                        for (int u = 0; u < UNITS; ++u) dh[u] += rec[u];
                        float (&dcr)[UNITS] = *reinterpret_cast<float (*)[UNITS]>(ST(dcst, l, rt, c, row));
                        bwd_cell_row(d, io, l, t, rt, c, row, dh, dcr, dgo);
                    }
            for (int rt = 0; rt < d.RT; ++rt)
                for (int c = 0; c < SLICES; ++c) {
                    std::fill(D2.begin(), D2.end(), 0.f);
                    const uint8_t *blk = dgs.data() + dg_block_index(d, l, t, rt) * d.dg_block_bytes();
                    for (int j = 0; j < SLICES; ++j) {
                        const int cc = (c + j) & (SLICES - 1);
                        for (int g = 0; g < 4; ++g) {
                            const int ks = 4 * cc + g;
                            const uint8_t *b = wb.data() + ((size_t)l * SLICES + c) * BWD_W_BYTES + (size_t)ks * BWD_W_KSTEP_BYTES;
                            for (int sub = 0; sub < d.nsub; ++sub)
                                mma_kstep(blk + (size_t)ks * d.stage_bytes() + (size_t)sub * SUB_BYTES, b, BWD_N, 1024, 512,
                                          &D2[(size_t)sub * TILE_M * BWD_N], BWD_N);
                        }
                    }
                    for (int row = 0; row < TR; ++row) {
                        const float *v = &D2[(size_t)row * BWD_N];
                        float *rec = ST(dhrec, l, rt, c, row);
                        for (int u = 0; u < UNITS; ++u) rec[u] = v[UNITS + u];
                        if (l > 0) {
                            float dx[UNITS];
                            for (int u = 0; u < UNITS; ++u) dx[u] = v[u];
                            store16_rowinner(dxbuf.data() + dx_block_offset(d, l, t, rt, c), TR, row, dx);
                        } else {
                            float *sum = ST(dxsum, l, rt, c, row);
                            for (int u = 0; u < UNITS; ++u) sum[u] += v[u];
                        }
                    }
                }
        }
    for (int rt = 0; rt < d.RT; ++rt)
        for (int c = 0; c < SLICES; ++c)
            for (int row = 0; row < TR; ++row) {
                const int64_t grow = (int64_t)rt * TR + row;
                if (grow >= R) continue;
                for (int u = 0; u < UNITS; ++u)
                    if (UNITS * c + u < E) dx0[(size_t)grow * E + UNITS * c + u] = ST(dxsum, 0, rt, c, row)[u];
            }
    {
        double err = 0, scale = 0;
        for (size_t i = 0; i < rdx.size(); ++i) { err = fmax(err, std::isnan(dx0[i]) ? 1e30 : fabs(rdx[i] - dx0[i])); scale = fmax(scale, fabs(rdx[i])); }
        report("backward dx", err / scale, 5e-5);
    }
    // weight gradients the way lstm_dw_kernel + lstm_finish_grads_kernel compute them: per (layer, which, m-tile) a 256 x 256
    // tile, K-blocks of 32 rows copied group by group into the stage image, operands read through MN-major descriptors
    // (element (m, k): start + (m/8)*SBO + (k/8)*LBO + (k%8)*16 + (m%8)*2 with LBO = 128, SBO = DW_GROUP_BYTES)
    {
        const int tiles = L * 2 * DW_M_TILES, n_kb = T * d.RT * d.nsub * 4;
        std::vector<float> partial((size_t)tiles * DW_TILE * DW_TILE, 0.f);
        std::vector<uint8_t> stage(DW_STAGE_BYTES);
        auto mn_elem = [&](const uint8_t *plane, int m, int k) { return bf(plane + (m / 8) * DW_GROUP_BYTES + (k / 8) * 128 + (k % 8) * 16 + (m % 8) * 2); };
        for (int tile = 0; tile < tiles; ++tile) {
            const int l = tile / (2 * DW_M_TILES), which = (tile / DW_M_TILES) & 1, mt = tile % DW_M_TILES;
            float *D = &partial[(size_t)tile * DW_TILE * DW_TILE];
            for (int kb = 0; kb < n_kb; ++kb) {
                const DwKBlock k = dw_kblock(d, kb);
                const uint8_t *a_blk = dgs.data() + dg_block_index(d, l, k.t, k.rt) * d.dg_block_bytes();
                const uint8_t *b_blk = actbuf.data() + dw_b_block(d, l, which, k.t, k.rt) * d.act_block_bytes();
                for (int lane = 0; lane < 32; ++lane)
                    for (int plane = 0; plane < 2; ++plane) {
                        memcpy(&stage[plane * DW_PLANE_BYTES + lane * DW_GROUP_BYTES],
                               a_blk + group_offset(d, mt * DW_GROUPS + lane, k.sub, plane) + k.rowblk * DW_GROUP_BYTES, DW_GROUP_BYTES);
                        memcpy(&stage[(2 + plane) * DW_PLANE_BYTES + lane * DW_GROUP_BYTES],
                               b_blk + group_offset(d, lane, k.sub, plane) + k.rowblk * DW_GROUP_BYTES, DW_GROUP_BYTES);
                    }
                for (int m = 0; m < DW_TILE; ++m)
                    for (int nn = 0; nn < DW_TILE; ++nn) {
                        float acc = 0.f;
                        for (int kk = 0; kk < DW_KB_ROWS; ++kk) {
                            const float ah = mn_elem(&stage[0], m, kk), al = mn_elem(&stage[DW_PLANE_BYTES], m, kk);
                            const float bh = mn_elem(&stage[2 * DW_PLANE_BYTES], nn, kk), bl = mn_elem(&stage[3 * DW_PLANE_BYTES], nn, kk);
                            acc += ah * bh + ah * bl + al * bh;
                        }
                        D[(size_t)m * DW_TILE + nn] += acc;
                    }
            }
        }
        for (int l = 0; l < L; ++l) {
            const int in = l == 0 ? E : H;
            double e_ih = 0, s_ih = 0, e_hh = 0, s_hh = 0, e_b = 0, s_b = 0;
            for (int row = 0; row < 4 * H; ++row) {
                const int gate = row / H, unit = row % H, m = gate_perm_index(unit, gate);
                for (int which = 0; which < 2; ++which) {
                    const int tile = (l * 2 + which) * DW_M_TILES + m / DW_TILE;
                    const float *src = &partial[((size_t)tile * DW_TILE + (m % DW_TILE)) * DW_TILE];
                    if (which == 0)
                        for (int k = 0; k < in; ++k) { e_ih = fmax(e_ih, fabs(src[k] - rdw_ih[l][(size_t)row * in + k])); s_ih = fmax(s_ih, fabs(rdw_ih[l][(size_t)row * in + k])); }
                    else {
                        for (int k = 0; k < H; ++k) { e_hh = fmax(e_hh, fabs(src[k] - rdw_hh[l][(size_t)row * H + k])); s_hh = fmax(s_hh, fabs(rdw_hh[l][(size_t)row * H + k])); }
                        e_b = fmax(e_b, fabs(src[H] - rdb[l][row])); s_b = fmax(s_b, fabs(rdb[l][row]));
                    }
                }
            }
            char name[64];
            snprintf(name, sizeof(name), "backward dW_ih[%d]", l); report(name, e_ih / s_ih, 5e-5);
            snprintf(name, sizeof(name), "backward dW_hh[%d]", l); report(name, e_hh / s_hh, 5e-5);
            snprintf(name, sizeof(name), "backward db[%d]", l); report(name, e_b / s_b, 5e-5);
        }
    }
    printf(bad ? "EMULATION FAILED (%d)\n" : "emulation ok\n", bad);
    return bad ? 1 : 0;
}
