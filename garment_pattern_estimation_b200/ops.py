"""Host-side operator layer: torch tensors in, libnt_b200.so kernels out.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every FLOP of the operators below runs in the
hand-written sm_100a kernels behind the C ABI of include/nt_b200.h.  Nothing in this file has a CPU or library
fallback: tensors must live on a CUDA device and the extension must load.

Operators (reference call sites in parentheses, relative to /root/reference):
  * ``knn_graph``        kNN indices of every point inside its cloud (torch_cluster.knn via nn/net_blocks.py:174)
  * ``fused_mlp``        MLP() = [Linear -> ReLU -> BatchNorm1d] x L  (nn/net_blocks.py:43-47), in two row modes:
                         'edge'  -- DynamicEdgeConv: message MLP on [x_i, x_j - x_i] + max aggregation over k
                                    (nn/net_blocks.py:127-135,174), optional skip-concat of the positions (:178-180)
                         'plain' -- per-point MLP (point_segment_mlp, nn/nets.py:223-226)
  * ``sparsemax``        sparsemax.Sparsemax(dim=1) on [rows, P<=32] (nn/nets.py:225)
  * ``attention_pool``   the per-panel weighted global_mean_pool loop as one contraction (nn/nets.py:263-276)
  * ``linear``           nn.Linear on rows (panel_dec_lin nn/nets.py:232, placement_decoder nn/nets.py:128)
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import (GemmArgs, NT_EPI_BIAS, NT_EPI_BNRELU_BWD, NT_EPI_RELU_MAXMIN, NT_EPI_RELU_STATS, NT_PROD_EDGE,
                   NT_PROD_PLAIN)


# ----------------------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


FLOP_SINK = None      # dict: kernel group -> useful flops (bench.py roofline_tensor)
BYTES_SINK = None     # dict: kernel group -> algorithmic HBM bytes (operands read once + results written once; bench.py roofline)
EVENT_SINK = None     # set to a dict to collect (start, end) CUDA-event pairs per kernel group (bench.py roofline)


def _call(name, fn, *args, group=None):
    """Invoke one C-ABI entry point, raise RuntimeError on a non-zero status, and -- when EVENT_SINK is a dict --
    bracket the launch with CUDA events on the launching stream."""
    if EVENT_SINK is None:
        _lib.check(fn(*args), name)
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    _lib.check(fn(*args), name)
    end.record()
    EVENT_SINK.setdefault(group or name, []).append((start, end))


def _require_cuda(*tensors):
    """Every operand on a CUDA device -- and on the CURRENT one: the kernels are enqueued on torch.cuda.current_stream(), so a
    tensor living on another GPU would be addressed from the wrong device."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('garment_pattern_estimation_b200: tensors must be on a CUDA device '
                               '(the B200 hot path has no CPU fallback)')
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError('garment_pattern_estimation_b200: tensor on cuda:{} but the current device is cuda:{} -- call '
                               'torch.cuda.set_device() / use torch.cuda.device() around the model'.format(t.device.index, cur))


def _pad4(c):
    return (c + 3) // 4 * 4


def _rowbuf(rows, cols, device):
    """[rows, cols] fp32 workspace whose row stride is padded to a multiple of 4 floats (16-byte aligned rows for the
    vectorised producers / epilogues).  Only the first `cols` columns are ever read or written."""
    return torch.empty(rows, _pad4(cols), dtype=torch.float32, device=device)


def _out_rows(rows, cols, device, padded):
    """Result tensor [rows, cols]: contiguous, or (padded) a view of a buffer with 16-byte aligned rows."""
    if padded:
        return _rowbuf(rows, cols, device)[:, :cols]
    return torch.empty(rows, cols, dtype=torch.float32, device=device)


def _rows2d(t):
    """(tensor, row stride) for a 2-D fp32 tensor whose last dim is dense."""
    assert t.dim() == 2 and t.dtype == torch.float32
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0)))


_EPI_NAMES = {NT_EPI_BIAS: 'bias', NT_EPI_RELU_STATS: 'relu_stats', NT_EPI_RELU_MAXMIN: 'relu_maxmin',
              NT_EPI_BNRELU_BWD: 'bnrelu_bwd'}


class EdgeSrc:
    """Operand description of NT_PROD_EDGE: relu(pq[centre, :H] + pq[nbr, qoff:qoff+H])."""

    def __init__(self, pq, H, idx=None, k=1, n_per_cloud=1):
        self.pq, self.H, self.idx, self.k, self.n_per_cloud = pq, H, idx, k, n_per_cloud
        self.ldpq = pq.stride(0)
        self.qoff = H if idx is not None else 0


GRAD_PRECISION = _lib.NT_PREC_BF16X3 if os.environ.get('NT_GRAD_PRECISION', 'tf32x3') == 'bf16x3' else _lib.NT_PREC_TF32X3
# EdgeConv: materialise the first edge activation a1 = relu(P[i] + Q[j]) once ([E, H] fp32) instead of re-gathering two
# random PQ rows per edge in every consumer (next GEMM, weight-gradient GEMM, BN/ReLU backward epilogue).  '0' keeps the
# gather fused into those kernels (smaller footprint, ~2x slower consumers).
EDGE_MATERIALIZE = os.environ.get('NT_EDGE_MATERIALIZE', '1') != '0'
# EdgeConv backward: scatter dz_1 into the per-point gradient dPQ inside the epilogue of the last data-gradient GEMM (dz_1 is
# never written) instead of a separate nt_edge_scatter pass.  Correct (tests) but OFF by default: measured at C2 the 16-byte
# reductions in the epilogue cost the GEMM +0.13 ms per launch while the separate pass costs 0.10 ms (8.89 vs 8.83 ms/step).
FUSED_SCATTER = os.environ.get('NT_FUSED_SCATTER', '0') != '0'
# Inference EdgeConv: one kernel per layer (gather -> GEMM -> GEMM -> max -> BN, csrc/edgeconv_eval.cu) instead of the layer-by-layer
# path; '0' keeps the latter (A/B measurements, and the reference point of the parity test).
EDGE_EVAL_FUSED = os.environ.get('NT_EDGE_EVAL_FUSED', '1') != '0'
EDGE_EVAL_PRECISION = _lib.NT_PREC_BF16X3 if os.environ.get('NT_EDGE_EVAL_PREC', 'tf32x3') == 'bf16x3' else _lib.NT_PREC_TF32X3
# Training-mode EdgeConv layers with three Linear stages run through the library's composite entry points (nt_edgeconv_train_fwd /
# _bwd: the whole layer per call, orchestrated in C++).  '0' keeps the kernel-by-kernel orchestration below (same kernels, same
# order; also what runs while bench.py brackets the individual kernels with CUDA events, and for other depths / the plain MLP).
EDGECONV_COMPOSITE = os.environ.get('NT_EDGECONV_COMPOSITE', '1') != '0'
NT_ENGINE = 0         # nt_gemm_args.engine of every row GEMM launched from here: 0 = auto (product); tests set 1 / 3 / 4 / 5


def prepare_weights(w, ldw, n_out, K, precision):
    """hi/lo split of W in UMMA core-matrix layout for the tensor-core engine (a few hundred KB)."""
    lib = _lib.load()
    nbytes = int(lib.nt_gemm_weights_bytes(n_out, K, precision))
    buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    _call('nt_gemm_prepare_weights', lib.nt_gemm_prepare_weights, _p(w), ldw, n_out, K, precision, _p(buf), _stream())
    return buf


def gemm_nt(rows, K, n_out, w, ldw, epilogue, a=None, lda=0, edge=None, bias=None, out=None, ldo=0, stats=None,
            agg=None, k_agg=0, aux=None, ldaux=0, aux_edge=False, k0=None, k1=None, mu=None, colsum=None, grad_gemm=False,
            scatter=None):
    """scatter: optional [M, 2*n_out] tensor (second half zeroed) for the fused edge scatter of NT_EPI_BNRELU_BWD (needs
    `edge` for idx / k / n_per_cloud).  Returns False -- without launching anything -- when the library cannot fuse the
    scatter for this call, True otherwise."""
    # TF32x3 (fp32-like, ~1e-6) everywhere: the tensor pipe is far from being the limiter of these gather-bound GEMMs,
    # and BatchNorm's backward is cancellation-heavy (sum_r da = 0), which amplifies BF16x3's 1e-5 to ~5e-3 on bias
    # gradients.  GRAD_PRECISION can be switched to NT_PREC_BF16X3 for the data-gradient GEMMs.
    precision = GRAD_PRECISION if (epilogue == NT_EPI_BNRELU_BWD or grad_gemm) else _lib.NT_PREC_TF32X3
    w_split = prepare_weights(w, ldw, n_out, K, precision)
    g = GemmArgs()
    g.w_split, g.precision, g.engine = _p(w_split), precision, int(NT_ENGINE)
    g.rows, g.K, g.n_out = int(rows), int(K), int(n_out)
    g.producer = NT_PROD_PLAIN if edge is None or a is not None else NT_PROD_EDGE
    g.epilogue = epilogue
    g.a, g.lda = _p(a), int(lda)
    if edge is not None:
        g.pq, g.ldpq, g.qoff = _p(edge.pq), int(edge.ldpq), int(edge.qoff)
        g.idx, g.k, g.n_per_cloud = _p(edge.idx), int(edge.k), int(edge.n_per_cloud)
    if k_agg:
        g.k = int(k_agg)
    g.w, g.ldw, g.bias = _p(w), int(ldw), _p(bias)
    g.out, g.ldo = _p(out), int(ldo)
    g.stats = _p(stats)
    if agg is not None:
        g.vmax, g.vmin, g.imax, g.imin = (_p(t) for t in agg)
    g.aux, g.ldaux, g.aux_edge = _p(aux), int(ldaux), int(bool(aux_edge))
    g.k0, g.k1, g.mu, g.colsum = _p(k0), _p(k1), _p(mu), _p(colsum)
    if scatter is not None:
        g.scatter_dpq, g.ldscatter = _p(scatter), int(scatter.stride(0))
        if not _lib.load().nt_gemm_nt_scatter_supported(ctypes.byref(g)):
            return False
    group = 'nt_gemm_nt[%s,%s]' % (_EPI_NAMES[epilogue], 'edge' if g.producer == NT_PROD_EDGE else 'plain')
    if FLOP_SINK is not None:
        FLOP_SINK[group] = FLOP_SINK.get(group, 0.0) + 2.0 * rows * K * n_out
    if BYTES_SINK is not None:
        cols = K + (n_out if out is not None else 0) + (n_out if epilogue == NT_EPI_BNRELU_BWD else 0)
        extra = 4.0 * scatter.shape[0] * scatter.shape[1] if scatter is not None else 0.0
        BYTES_SINK[group] = BYTES_SINK.get(group, 0.0) + 4.0 * rows * cols + 4.0 * n_out * K + extra
    _call('nt_gemm_nt', _lib.load().nt_gemm_nt, ctypes.byref(g), _stream(), group=group)
    return True


def gemm_tn(a, lda, m, rows, out, b=None, ldb=0, n=0, edge=None, mu=None):
    """out[m, n] += sum_r a[r, m] * Bop[r, n].  With `mu` the B operand is centred and `out` must be float64."""
    lib = _lib.load()
    ws = torch.empty(int(lib.nt_gemm_tn_workspace_bytes()), dtype=torch.uint8, device=a.device)
    if edge is not None:
        bop = (None, 0, n, rows, _p(edge.pq), edge.ldpq, edge.qoff, _p(edge.idx), edge.k, edge.n_per_cloud)
    else:
        bop = (_p(b), ldb, n, rows, None, 0, 0, None, 1, 1)
    if FLOP_SINK is not None:
        name = 'nt_gemm_tn_centered' if mu is not None else 'nt_gemm_tn'
        FLOP_SINK[name] = FLOP_SINK.get(name, 0.0) + 2.0 * rows * m * n
    if BYTES_SINK is not None:
        name = 'nt_gemm_tn_centered' if mu is not None else 'nt_gemm_tn'
        BYTES_SINK[name] = BYTES_SINK.get(name, 0.0) + 4.0 * rows * (m + n) + 4.0 * m * n
    if mu is not None:
        assert out.dtype == torch.float64
        _call('nt_gemm_tn_centered', lib.nt_gemm_tn_centered, _p(a), lda, m, *bop, _p(mu), _p(out), out.stride(0),
              _p(ws), _stream())
    else:
        _call('nt_gemm_tn', lib.nt_gemm_tn, _p(a), lda, m, *bop, _p(out), out.stride(0), _p(ws), _stream())


# ----------------------------------------------------------------------------------------------------------
# kNN
# ----------------------------------------------------------------------------------------------------------
def knn_graph(x, B, N, k):
    """x: [B*N, D] fp32 (row stride free).  Returns idx [B*N, k] int32, LOCAL to each cloud, ascending by
    (squared distance, index), self included.  Bit-exact with oracle/knn_oracle.c."""
    _require_cuda(x)
    x, ldx = _rows2d(x)
    if x.shape[0] != B * N:
        raise RuntimeError('knn_graph: expected {} rows, got {}'.format(B * N, x.shape[0]))
    idx = torch.empty(B * N, k, dtype=torch.int32, device=x.device)
    lib = _lib.load()
    ws = torch.empty(int(lib.nt_knn_workspace_bytes(B, N, k)), dtype=torch.uint8, device=x.device)
    _call('nt_knn', lib.nt_knn, _p(x), B, N, x.shape[1], ldx, k, _p(idx), _p(ws), _stream(),
          group='nt_knn[D=%d]' % x.shape[1])
    return idx


def edge_index_from_knn(idx, B, N):
    """PyG-style edge_index [2, E] int64 (row 0 = source/neighbour j, row 1 = target/centre i), the layout
    DynamicEdgeConv builds from torch_cluster.knn(...).flip([0])."""
    M, k = idx.shape
    base = (torch.arange(M, device=idx.device) // N * N).unsqueeze(1)
    src = (idx.long() + base).reshape(-1)
    dst = torch.arange(M, device=idx.device).repeat_interleave(k)
    return torch.stack([src, dst], dim=0)


# ----------------------------------------------------------------------------------------------------------
# fused MLP  (Linear -> ReLU -> BN) x L, edge or plain rows
# ----------------------------------------------------------------------------------------------------------
class _BNBuffers:
    """Non-differentiable BatchNorm state handed to the autograd function (updated in place like nn.BatchNorm1d)."""

    def __init__(self, running_mean, running_var, num_batches_tracked, momentum, eps):
        self.running_mean, self.running_var, self.nbt = running_mean, running_var, num_batches_tracked
        if momentum is None:
            raise NotImplementedError('BatchNorm1d(momentum=None) (cumulative moving average) is not built on the B200 path; the '
                                      'reference uses the default momentum 0.1 (nn/net_blocks.py:46)')
        self.momentum = float(momentum)
        self.eps = float(eps)


def _ptr3(ts):
    arr = (ctypes.c_void_p * 3)()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _edgeconv_args(x, ldx, idx, tail_src, k, n_per_cloud, Ws, bs, gammas, betas, bn_bufs=None):
    """nt_edgeconv_args for one EdgeConv layer (three stages); the caller fills out / saved / scratch / gradient pointers."""
    g = _lib.EdgeConvArgs()
    g.M, g.C = int(x.shape[0]), int(x.shape[1])
    g.H1, g.H2, g.H3 = (int(W.shape[0]) for W in Ws)
    g.k, g.n_per_cloud = int(k), int(n_per_cloud)
    g.x, g.ldx, g.idx = _p(x), int(ldx), _p(idx)
    if tail_src is not None:
        ts, tld = _rows2d(tail_src)
        g.tail_src, g.tail_ld, g.tail = _p(ts), int(tld), int(ts.shape[1])
    g.W, g.b, g.gamma, g.beta = _ptr3(Ws), _ptr3(bs), _ptr3(gammas), _ptr3(betas)
    if bn_bufs is not None:
        g.running_mean, g.running_var = _ptr3([b.running_mean for b in bn_bufs]), _ptr3([b.running_var for b in bn_bufs])
        g.num_batches_tracked = _ptr3([b.nbt for b in bn_bufs])
        g.momentum, g.eps = bn_bufs[0].momentum, bn_bufs[0].eps
    return g


class _FusedMLPFunction(torch.autograd.Function):
    """params = (W_0, b_0, gamma_1, beta_1, W_1, b_1, gamma_2, beta_2, ...): 4 tensors per layer."""

    @staticmethod
    def forward(ctx, x, idx, tail_src, meta, *params):
        lib = _lib.load()
        mode, training, bn_bufs = meta['mode'], meta['training'], meta['bn']
        L = len(params) // 4
        if L < 2:
            raise NotImplementedError('fused_mlp needs at least two Linear layers (EConv_hidden_depth >= 1)')
        Ws = [params[4 * l] for l in range(L)]
        bs = [params[4 * l + 1] for l in range(L)]
        gammas = [params[4 * l + 2] for l in range(L)]
        betas = [params[4 * l + 3] for l in range(L)]
        _require_cuda(x, *params)
        dev = x.device
        x, ldx = _rows2d(x)
        M, C = x.shape
        widths = [W.shape[0] for W in Ws]           # H_1 .. H_L
        f32 = dict(dtype=torch.float32, device=dev)

        # ---- training, EdgeConv with three Linear layers: the whole layer is ONE library call (csrc/edgeconv_train.cu)
        if (mode == 'edge' and training and L == 3 and EDGECONV_COMPOSITE and EDGE_MATERIALIZE and not FUSED_SCATTER
                and EVENT_SINK is None and all(b is not None for b in bs)
                and all(b.momentum == bn_bufs[0].momentum and b.eps == bn_bufs[0].eps for b in bn_bufs)):
            cWs, cbs = [W.contiguous() for W in Ws], [b.contiguous() for b in bs]
            cgs, cbetas = [t.contiguous() for t in gammas], [t.contiguous() for t in betas]
            if Ws[0].shape[1] != 2 * C:
                raise RuntimeError('edge MLP expects first Linear with {} inputs, got {}'.format(2 * C, Ws[0].shape[1]))
            k, N = meta['k'], meta['n_per_cloud']
            g = _edgeconv_args(x, ldx, idx, tail_src, k, N, cWs, cbs, cgs, cbetas, bn_bufs)
            tail = int(g.tail)
            out = _out_rows(M, widths[2] + tail, dev, meta.get('pad_out'))
            saved = torch.empty(int(lib.nt_edgeconv_saved_bytes(ctypes.byref(g))), dtype=torch.uint8, device=dev)
            scratch = torch.empty(int(lib.nt_edgeconv_scratch_bytes(ctypes.byref(g), 0)), dtype=torch.uint8, device=dev)
            g.out, g.ldo, g.saved, g.scratch = _p(out), out.stride(0), _p(saved), _p(scratch)
            _call('nt_edgeconv_train_fwd', lib.nt_edgeconv_train_fwd, ctypes.byref(g), _stream())
            ctx.meta = None
            if any(ctx.needs_input_grad):
                ctx.meta = dict(composite=True, k=k, N=N, tail=tail, ldx=ldx)
                ctx.save_for_backward(x, idx, saved, *cWs, *cbs, *cgs, *cbetas)
            return out

        # ---- first Linear, evaluated per POINT (edge mode: split W_0 = [W_a | W_b] on [x_i, x_j - x_i])
        H1 = widths[0]
        if mode == 'edge':
            k, N = meta['k'], meta['n_per_cloud']
            W0 = Ws[0]
            if W0.shape[1] != 2 * C:
                raise RuntimeError('edge MLP expects first Linear with {} inputs, got {}'.format(2 * C, W0.shape[1]))
            Wc = torch.cat([W0[:, :C] - W0[:, C:], W0[:, C:]], dim=0).contiguous()      # [2*H1, C]
            bc = torch.cat([bs[0], torch.zeros_like(bs[0])])
            pq = _rowbuf(M, 2 * H1, dev)
            gemm_nt(M, C, 2 * H1, Wc, C, NT_EPI_BIAS, a=x, lda=ldx, bias=bc, out=pq, ldo=pq.stride(0))
            R = M * k
            src = EdgeSrc(pq, H1, idx, k, N)
        else:
            k, N = 1, 1
            Wc = Ws[0].contiguous()
            pq = _rowbuf(M, H1, dev)
            gemm_nt(M, C, H1, Wc, C, NT_EPI_BIAS, a=x, lda=ldx, bias=bs[0].contiguous(), out=pq, ldo=pq.stride(0))
            R = M
            src = EdgeSrc(pq, H1)

        def fold(layer, stats, next_layer):
            """BN `layer` (0-based, follows Linear `layer`): stats -> affine; fold into Linear `next_layer`."""
            Cn = widths[layer]
            vec = torch.empty(4, Cn, **f32)           # mean, rstd, s, t
            buf = bn_bufs[layer]
            if next_layer is not None:
                n_next = widths[next_layer]
                w_f = torch.empty(n_next, Cn, **f32)
                w_ft = torch.empty(Cn, n_next, **f32)
                b_f = torch.empty(n_next, **f32)
                wn, bn_ = Ws[next_layer].contiguous(), bs[next_layer].contiguous()
            else:
                n_next, w_f, w_ft, b_f, wn, bn_ = 0, None, None, None, None, None
            _call('nt_bn_fold', _lib.load().nt_bn_fold, _p(stats), R, Cn, _p(gammas[layer].contiguous()), _p(betas[layer].contiguous()),
                                      _p(buf.running_mean), _p(buf.running_var), _p(buf.nbt), buf.momentum, buf.eps,
                                      int(training), _p(vec[0]), _p(vec[1]), _p(vec[2]), _p(vec[3]),
                                      _p(wn), _p(bn_), n_next, _p(w_f), _p(w_ft), _p(b_f), _stream())
            return vec, w_f, w_ft, b_f

        # ---- inference, EdgeConv with three Linear layers: everything after PQ is ONE kernel; no edge-sized tensor touches HBM
        if (mode == 'edge' and not training and L == 3 and EDGE_EVAL_FUSED
                and lib.nt_edgeconv_eval_supported(H1, widths[1], widths[2], k, pq.stride(0))):
            _, w2f, _, b2f = fold(0, None, 1)
            _, w3f, _, b3f = fold(1, None, 2)
            vec3, _, _, _ = fold(2, None, None)
            w2s = prepare_weights(w2f, H1, widths[1], H1, EDGE_EVAL_PRECISION)
            w3s = prepare_weights(w3f, widths[1], widths[2], widths[1], EDGE_EVAL_PRECISION)
            tail = 0 if tail_src is None else tail_src.shape[1]
            ts, tld = (None, 0) if not tail else _rows2d(tail_src)
            out = _out_rows(M, widths[2] + tail, dev, meta.get('pad_out'))
            _call('nt_edgeconv_eval_fwd', lib.nt_edgeconv_eval_fwd, _p(pq), pq.stride(0), H1, _p(idx), k, N, M, _p(w2s), _p(b2f),
                  widths[1], _p(w3s), _p(b3f), widths[2], EDGE_EVAL_PRECISION, _p(vec3[2]), _p(vec3[3]), _p(ts), tld, tail, _p(out),
                  out.stride(0),
                  _stream())
            ctx.meta = None
            return out

        # one zero-fill for the BatchNorm statistics of every layer of this call (was one fill kernel per layer)
        stats_all = torch.zeros(2 * sum(widths), dtype=torch.float64, device=dev) if training else None
        stats_off = [2 * sum(widths[:l]) for l in range(L)]

        def stats_of(layer):
            return stats_all[stats_off[layer]:stats_off[layer] + 2 * widths[layer]] if training else None

        # ---- BN_1 statistics of a_1 = relu(P + Q) (no GEMM at row level); edge mode also materialises a_1
        stats = stats_of(0)
        a1 = None
        if EDGE_MATERIALIZE:           # both row modes: every consumer then streams a plain, 16-byte aligned operand
            a1 = _rowbuf(R, H1, dev)
            _call('nt_edge_activation', _lib.load().nt_edge_activation, _p(pq), src.ldpq, src.qoff, _p(src.idx), src.k,
                  src.n_per_cloud, R, H1, _p(a1), a1.stride(0), _p(stats), _stream())
        elif training:
            _call('nt_edge_stats', _lib.load().nt_edge_stats, _p(pq), src.ldpq, src.qoff, _p(src.idx), src.k, src.n_per_cloud, R, H1,
                                         _p(stats), _stream())
        first_in = dict(edge=src) if a1 is None else dict(a=a1, lda=a1.stride(0))
        bn_vec = [None] * L
        w_fts = [None] * L            # w_fts[l] = (W_l . diag(s_l))^T, l >= 1
        bn_vec[0], w_f, w_fts[1], b_f = fold(0, stats, 1)

        # ---- middle layers: a_{l+1} = relu(a_l . W_l'^T + b_l'), statistics in the epilogue
        acts = [None] * (L + 1)       # acts[l] = a_l for l >= 2 (a_1 is recomputed from pq)
        for l in range(1, L - 1):
            Hn = widths[l]
            stats = stats_of(l)
            out = _rowbuf(R, Hn, dev)
            if l == 1:
                gemm_nt(R, widths[0], Hn, w_f, widths[0], NT_EPI_RELU_STATS, bias=b_f, out=out,
                        ldo=out.stride(0), stats=stats, **first_in)
            else:
                gemm_nt(R, widths[l - 1], Hn, w_f, widths[l - 1], NT_EPI_RELU_STATS, a=acts[l], lda=acts[l].stride(0),
                        bias=b_f, out=out, ldo=out.stride(0), stats=stats)
            acts[l + 1] = out
            bn_vec[l], w_f, w_fts[l + 1], b_f = fold(l, stats, l + 1)

        # ---- last layer
        HL, Kin = widths[L - 1], widths[L - 2]
        stats = stats_of(L - 1)
        last_in = first_in if L == 2 else dict(a=acts[L - 1], lda=acts[L - 1].stride(0))
        need_bwd = training and any(ctx.needs_input_grad)
        if mode == 'edge':
            tail = 0 if tail_src is None else tail_src.shape[1]
            a_last = _rowbuf(R, HL, dev) if need_bwd else None
            ld_last = a_last.stride(0) if need_bwd else 0
            agg = (torch.empty(M, HL, **f32), torch.empty(M, HL, **f32),
                   torch.empty(M, HL, dtype=torch.uint8, device=dev), torch.empty(M, HL, dtype=torch.uint8, device=dev))
            gemm_nt(R, Kin, HL, w_f, Kin, NT_EPI_RELU_MAXMIN, bias=b_f, out=a_last, ldo=ld_last, stats=stats, agg=agg,
                    k_agg=k, **last_in)
            bn_vec[L - 1], _, _, _ = fold(L - 1, stats, None)
            out = _out_rows(M, HL + tail, dev, meta.get('pad_out'))
            sel = torch.empty(M, HL, dtype=torch.uint8, device=dev) if need_bwd else None
            vsel = torch.empty(M, HL, **f32) if need_bwd else None
            ts, tld = (None, 0)
            if tail:
                ts, tld = _rows2d(tail_src)
            _call('nt_maxmin_finish', _lib.load().nt_maxmin_finish, _p(agg[0]), _p(agg[1]), _p(agg[2]), _p(agg[3]), _p(bn_vec[L - 1][2]),
                                            _p(bn_vec[L - 1][3]), M, HL, _p(out), out.stride(0), _p(sel), _p(vsel),
                                            _p(ts), tld, tail, _stream())
        else:
            a_last = _rowbuf(R, HL, dev)
            gemm_nt(R, Kin, HL, w_f, Kin, NT_EPI_RELU_STATS, bias=b_f, out=a_last, ldo=a_last.stride(0), stats=stats,
                    **last_in)
            bn_vec[L - 1], _, _, _ = fold(L - 1, stats, None)
            out = torch.empty(M, HL, **f32)
            _call('nt_bn_apply', _lib.load().nt_bn_apply, _p(a_last), a_last.stride(0), _p(bn_vec[L - 1][2]), _p(bn_vec[L - 1][3]), R, HL, _p(out), HL,
                                       _stream())
            sel = vsel = None
            tail = 0
        acts[L] = a_last

        ctx.meta = None
        if need_bwd:
            ctx.meta = dict(mode=mode, k=k, N=N, L=L, widths=widths, M=M, C=C, R=R, tail=tail, ldx=ldx)
            ctx.save_for_backward(x, idx, pq, Wc, sel, vsel, a1, *Ws, *[a for a in acts[2:]], *bn_vec, *w_fts[1:], *betas)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        m = ctx.meta
        if m is None:
            raise RuntimeError('fused_mlp: backward through an eval-mode (running-statistics) forward is not supported')
        if m.get('composite'):
            sv = ctx.saved_tensors
            x, idx, saved = sv[:3]
            Ws, bs, gammas, betas = sv[3:6], sv[6:9], sv[9:12], sv[12:15]
            dev = x.device
            g = _edgeconv_args(x, m['ldx'], idx, None, m['k'], m['N'], Ws, bs, gammas, betas)
            gout, ldg = _rows2d(gout)
            scratch = torch.empty(int(lib.nt_edgeconv_scratch_bytes(ctypes.byref(g), 1)), dtype=torch.uint8, device=dev)
            gW = [torch.empty_like(W) for W in Ws]
            gb = [torch.empty_like(b) for b in bs]
            gg = [torch.empty_like(t) for t in gammas]
            gbeta = [torch.empty_like(t) for t in betas]
            gx = _rowbuf(x.shape[0], x.shape[1], dev)[:, :x.shape[1]] if ctx.needs_input_grad[0] else None
            g.saved, g.scratch, g.gout, g.ldg = _p(saved), _p(scratch), _p(gout), int(ldg)
            g.gx, g.ldgx = _p(gx), int(gx.stride(0)) if gx is not None else 0
            g.gW, g.gb, g.ggamma, g.gbeta = _ptr3(gW), _ptr3(gb), _ptr3(gg), _ptr3(gbeta)
            _call('nt_edgeconv_train_bwd', lib.nt_edgeconv_train_bwd, ctypes.byref(g), _stream())
            g_tail = None
            H3 = Ws[2].shape[0]
            if m['tail'] and ctx.needs_input_grad[2]:
                g_tail = gout[:, H3:H3 + m['tail']].contiguous()
            flat = []
            for l in range(3):
                flat += [gW[l], gb[l], gg[l], gbeta[l]]
            return (gx, None, g_tail, None, *flat)
        mode, k, N, L, widths, M, C, R, tail = (m[key] for key in ('mode', 'k', 'N', 'L', 'widths', 'M', 'C', 'R', 'tail'))
        sv = ctx.saved_tensors
        x, idx, pq, Wc, sel, vsel, a1 = sv[:7]
        sv = sv[1:]                                                       # (a1 shifts the remaining slots by one)
        Ws = list(sv[6:6 + L])
        acts = [None, a1] + list(sv[6 + L:6 + L + (L - 1)])               # acts[1] (edge mode, materialised), acts[2..L]
        bn_vec = list(sv[6 + 2 * L - 1:6 + 3 * L - 1])
        w_fts = [None] + list(sv[6 + 3 * L - 1:6 + 4 * L - 2])            # w_fts[1..L-1]
        betas = list(sv[6 + 4 * L - 2:6 + 5 * L - 2])
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        H1, HL = widths[0], widths[L - 1]
        src = EdgeSrc(pq, H1, idx, k, N) if mode == 'edge' else EdgeSrc(pq, H1)
        gout, ldg = _rows2d(gout)

        grads_W, grads_b, grads_g, grads_beta = [None] * L, [None] * L, [None] * L, [None] * L

        # ---- trailing BN (after the aggregation in edge mode): column sums, then dz_L for every row
        # one zero-fill for every double accumulator of the backward: [sums 2HL | csum_l (l = L-1 .. 0) | raw_l (l = L-1 .. 1)]
        n_acc = 2 * HL + sum(widths) + sum(widths[l] * widths[l - 1] for l in range(1, L))
        acc_all = torch.zeros(n_acc, **f64)
        acc_pos = [0]

        def take(n):
            out = acc_all[acc_pos[0]:acc_pos[0] + n]
            acc_pos[0] += n
            return out

        mean, rstd, s, t = bn_vec[L - 1]
        sums = take(2 * HL)
        v_ref = vsel if mode == 'edge' else acts[L]
        _call('nt_bn_bwd_reduce', _lib.load().nt_bn_bwd_reduce, _p(gout), ldg, _p(v_ref), v_ref.stride(0), _p(mean), _p(rstd), M, HL, _p(sums), _stream())
        grads_beta[L - 1] = sums[:HL].float()
        grads_g[L - 1] = sums[HL:].float()
        dz = _rowbuf(R, HL, dev)
        csum = take(HL)
        _call('nt_bn_relu_bwd_last', _lib.load().nt_bn_relu_bwd_last, _p(acts[L]), acts[L].stride(0), _p(gout), ldg, _p(sel), k, _p(s), _p(mean), _p(rstd),
                                           _p(sums), R, R, HL, _p(dz), dz.stride(0), _p(csum), _stream())

        # ---- walk down the Linear layers L-1 .. 1
        dpq = None
        for l in range(L - 1, 0, -1):
            Hout, Hin = widths[l], widths[l - 1]
            pmean, prstd, ps, pt = bn_vec[l - 1]
            raw = take(Hout * Hin).view(Hout, Hin)         # dz^T . (a_l - mean_l), accumulated in double
            if l == 1 and a1 is None:
                gemm_tn(dz, dz.stride(0), Hout, R, raw, n=Hin, edge=src, mu=pmean)
            else:
                gemm_tn(dz, dz.stride(0), Hout, R, raw, b=acts[l], ldb=acts[l].stride(0), n=Hin, mu=pmean)
            dW = torch.empty(Hout, Hin, **f32)
            db = torch.empty(Hout, **f32)
            vecs = torch.empty(4, Hin, **f32)          # dgamma, dbeta, k0, k1
            _call('nt_linear_bn_bwd', _lib.load().nt_linear_bn_bwd, _p(raw), _p(csum), Hout, Hin, _p(Ws[l].contiguous()), _p(ps),
                                            _p(betas[l - 1].contiguous()), _p(prstd), R, _p(dW), _p(db), _p(vecs[0]),
                                            _p(vecs[1]), _p(vecs[2]), _p(vecs[3]), _stream())
            grads_W[l], grads_b[l] = dW, db
            grads_g[l - 1], grads_beta[l - 1] = vecs[0], vecs[1]
            csum_prev = take(Hin)
            if l == 1 and mode == 'edge' and a1 is not None and FUSED_SCATTER:
                dpq = torch.zeros(M, 2 * H1, **f32)
                if gemm_nt(R, Hout, Hin, w_fts[l], Hout, NT_EPI_BNRELU_BWD, a=dz, lda=dz.stride(0), edge=src, aux=acts[l],
                           ldaux=acts[l].stride(0), k0=vecs[2], k1=vecs[3], mu=pmean, colsum=csum_prev, scatter=dpq):
                    dz, csum = None, csum_prev
                    continue
                dpq = None
            if l == 1 and a1 is None:
                dz_prev = _rowbuf(R, Hin, dev)
                gemm_nt(R, Hout, Hin, w_fts[l], Hout, NT_EPI_BNRELU_BWD, a=dz, lda=dz.stride(0), edge=src, out=dz_prev,
                        ldo=dz_prev.stride(0), aux_edge=True, k0=vecs[2], k1=vecs[3], mu=pmean, colsum=csum_prev)
            else:
                dz_prev = _rowbuf(R, Hin, dev)         # saved activations stay intact (retain_graph-safe)
                gemm_nt(R, Hout, Hin, w_fts[l], Hout, NT_EPI_BNRELU_BWD, a=dz, lda=dz.stride(0), out=dz_prev,
                        ldo=dz_prev.stride(0), aux=acts[l], ldaux=acts[l].stride(0), k0=vecs[2], k1=vecs[3], mu=pmean, colsum=csum_prev)
            dz, csum = dz_prev, csum_prev

        # ---- first Linear (per point)
        grads_b[0] = csum.float()
        gx = None
        if mode == 'edge':
            if dpq is None:
                dpq = torch.zeros(M, 2 * H1, **f32)
                _call('nt_edge_scatter', _lib.load().nt_edge_scatter, _p(dz), dz.stride(0), _p(idx), k, N, M, H1, _p(dpq), 2 * H1, _stream())
            dWc = torch.zeros(2 * H1, C, **f32)
            gemm_tn(dpq, 2 * H1, 2 * H1, M, dWc, b=x, ldb=m['ldx'], n=C)
            grads_W[0] = torch.cat([dWc[:H1], dWc[H1:] - dWc[:H1]], dim=1)
            if ctx.needs_input_grad[0]:
                gx = _rowbuf(M, C, dev)[:, :C]         # 16-byte aligned rows: the call streams (TMA result boxes)
                WcT = Wc.t().contiguous()              # [C, 2*H1]
                gemm_nt(M, 2 * H1, C, WcT, 2 * H1, NT_EPI_BIAS, a=dpq, lda=2 * H1, out=gx, ldo=gx.stride(0), grad_gemm=True)
        else:
            dW0 = torch.zeros(H1, C, **f32)
            gemm_tn(dz, dz.stride(0), H1, M, dW0, b=x, ldb=m['ldx'], n=C)
            grads_W[0] = dW0
            if ctx.needs_input_grad[0]:
                gx = torch.empty(M, C, **f32)
                WT = Wc.t().contiguous()               # [C, H1]
                gemm_nt(M, H1, C, WT, H1, NT_EPI_BIAS, a=dz, lda=dz.stride(0), out=gx, ldo=C, grad_gemm=True)

        g_tail = None
        if tail and ctx.needs_input_grad[2]:
            g_tail = gout[:, HL:HL + tail].contiguous()
        flat = []
        for l in range(L):
            flat += [grads_W[l], grads_b[l], grads_g[l], grads_beta[l]]
        return (gx, None, g_tail, None, *flat)


def fused_mlp(x, layers, training, mode='plain', idx=None, k=1, n_per_cloud=1, tail_src=None, pad_out=False):
    """layers: list of (nn.Linear, nn.BatchNorm1d) pairs (the reference's Sequential(Linear, ReLU, BatchNorm1d)).
    pad_out (edge mode): return a [M, C] view of a buffer whose rows are padded to a multiple of 4 floats, so that the NEXT
    EdgeConv layer streams 16-byte aligned rows (set for the layers whose output only feeds the next layer)."""
    params, bufs = [], []
    for lin, bn in layers:
        params += [lin.weight, lin.bias, bn.weight, bn.bias]
        bufs.append(_BNBuffers(bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps))
    meta = dict(mode=mode, training=bool(training), bn=bufs, k=int(k), n_per_cloud=int(n_per_cloud), pad_out=bool(pad_out))
    return _FusedMLPFunction.apply(x, idx, tail_src, meta, *params)


# ----------------------------------------------------------------------------------------------------------
# sparsemax / attention pooling / linear
# ----------------------------------------------------------------------------------------------------------
class _SparsemaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        _require_cuda(z)
        z = z.contiguous()
        rows, P = z.shape
        out = torch.empty_like(z)
        _call('nt_sparsemax_fwd', _lib.load().nt_sparsemax_fwd, _p(z), rows, P, _p(out), _stream())
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        out, = ctx.saved_tensors
        g = g.contiguous()
        gz = torch.empty_like(out)
        _call('nt_sparsemax_bwd', _lib.load().nt_sparsemax_bwd, _p(out), _p(g), out.shape[0], out.shape[1], _p(gz), _stream())
        return gz


def sparsemax(z):
    """Sparsemax over dim 1 of a [rows, P] tensor, P <= 32."""
    if z.dim() != 2:
        raise RuntimeError('sparsemax expects [rows, P]')
    return _SparsemaxFunction.apply(z)


class _AttnPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, feat, B, N, scale):
        _require_cuda(w, feat)
        w = w.contiguous()
        feat, ldf = _rows2d(feat)
        P, F = w.shape[1], feat.shape[1]
        enc = torch.zeros(B, P, F, dtype=torch.float32, device=w.device)
        _call('nt_attn_pool_fwd', _lib.load().nt_attn_pool_fwd, _p(w), _p(feat), ldf, B, N, P, F, scale, _p(enc), _stream())
        ctx.save_for_backward(w, feat)
        ctx.dims = (B, N, P, F, scale, ldf)
        return enc

    @staticmethod
    def backward(ctx, genc):
        w, feat = ctx.saved_tensors
        B, N, P, F, scale, ldf = ctx.dims
        genc = genc.contiguous()
        gw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        gfeat = torch.empty(B * N, F, dtype=torch.float32, device=w.device) if ctx.needs_input_grad[1] else None
        _call('nt_attn_pool_bwd', _lib.load().nt_attn_pool_bwd, _p(genc), _p(w), _p(feat), ldf, B, N, P, F, scale, _p(gw), _p(gfeat), F,
                                                0, _stream())
        return gw, gfeat, None, None, None


def attention_pool(weights, feats, B, N, scale):
    """enc[b, p, :] = scale * sum_n weights[b*N+n, p] * feats[b*N+n, :]  ->  [B, P, F]."""
    return _AttnPoolFunction.apply(weights, feats, B, N, float(scale))


_POOL_MODES = {'mean': 0, 'max': 1, 'add': 2}


class _GlobalPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, N, mode):
        _require_cuda(x)
        x, ldx = _rows2d(x)
        F = x.shape[1]
        out = torch.empty(B, F, dtype=torch.float32, device=x.device)
        arg = torch.empty(B, F, dtype=torch.int32, device=x.device) if mode == 1 else None
        _call('nt_global_pool_fwd', _lib.load().nt_global_pool_fwd, _p(x), ldx, B, N, F, mode, _p(out), _p(arg), _stream())
        ctx.save_for_backward(arg)
        ctx.dims = (B, N, F, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        arg, = ctx.saved_tensors
        B, N, F, mode = ctx.dims
        g = g.contiguous()
        gx = torch.empty(B * N, F, dtype=torch.float32, device=g.device)
        _call('nt_global_pool_bwd', _lib.load().nt_global_pool_bwd, _p(g), _p(arg), B, N, F, mode, _p(gx), F, _stream())
        return gx, None, None, None


def global_pool(x, B, N, mode):
    """torch_geometric.nn.global_{mean,max,add}_pool on the dense equal-size layout: [B*N, F] -> [B, F]."""
    if mode not in _POOL_MODES:
        raise ValueError('{} pooling is not supported'.format(mode))
    if x.shape[0] != B * N:
        raise RuntimeError('global_pool: expected {} rows, got {}'.format(B * N, x.shape[0]))
    return _GlobalPoolFunction.apply(x, int(B), int(N), _POOL_MODES[mode])


# ----------------------------------------------------------------------------------------------------------
# LSTM decoder (persistent tcgen05 kernels, csrc/lstm.cu)
# ----------------------------------------------------------------------------------------------------------
LSTM_MAX_HIDDEN, LSTM_MAX_INPUT, LSTM_MAX_LAYERS = 255, 256, 4
_LSTM_SKIP_DW = False          # development knob (tools/lstm_check.py): time the recurrence kernel without the dW GEMMs
_LSTM_WEIGHT_CACHE = {}


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def lstm_supported(input_size, hidden_size, num_layers):
    return hidden_size <= LSTM_MAX_HIDDEN and input_size <= LSTM_MAX_INPUT and num_layers <= LSTM_MAX_LAYERS


def _lstm_sizes(R, T, L, H, E):
    z = _lib.LstmSizes()
    _lib.check(_lib.load().nt_lstm_sizes(R, T, L, H, E, ctypes.byref(z)), 'nt_lstm_sizes')
    return z


def _lstm_prepared_weights(params, L, H, E, nbytes, cacheable):
    """hi / lo split of the LSTM weights in the kernels' operand layouts.  Re-done whenever the parameters may have changed:
    always while gradients are recorded or a CUDA graph is being captured (a replayed graph must contain the preparation of
    the weights it runs on), cached per parameter version otherwise (inference)."""
    key = tuple((p.data_ptr(), p._version) for p in params)
    if cacheable:
        hit = _LSTM_WEIGHT_CACHE.get('w')
        if hit is not None and hit[0] == key:
            return hit[1]
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device=params[0].device)
    ps = [p.detach().contiguous() for p in params]
    _call('nt_lstm_prepare_weights', _lib.load().nt_lstm_prepare_weights, _ptr_array(ps[0::4]), _ptr_array(ps[1::4]),
          _ptr_array(ps[2::4]), _ptr_array(ps[3::4]), L, H, E, _p(buf), _stream())
    if cacheable:
        _LSTM_WEIGHT_CACHE['w'] = (key, buf, ps)
    return buf


class _LSTMDecoderFunction(torch.autograd.Function):
    """params = (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, weight_ih_l1, ...)."""

    @staticmethod
    def forward(ctx, x, h0, c0, T, *params):
        lib = _lib.load()
        _require_cuda(x, h0, c0, *params)
        L = len(params) // 4
        H = params[1].shape[1]
        x, ldx = _rows2d(x)
        R, E = x.shape
        if tuple(h0.shape) != (L, R, H) or tuple(c0.shape) != (L, R, H):
            raise RuntimeError('lstm_decoder: initial states must be [{}, {}, {}]'.format(L, R, H))
        h0, c0 = h0.contiguous().float(), c0.contiguous().float()
        z = _lstm_sizes(R, T, L, H, E)
        need_bwd = any(ctx.needs_input_grad)          # (grad mode is off inside Function.forward: ask autograd instead)
        capturing = torch.cuda.is_current_stream_capturing()
        w = _lstm_prepared_weights(params, L, H, E, z.weights_bytes, cacheable=not need_bwd and not capturing)
        dev = x.device
        y = torch.empty(T, R, z.y_ld, dtype=torch.float32, device=dev)
        act = torch.empty(int(z.act_bytes), dtype=torch.uint8, device=dev)
        cs = torch.empty(z.cs_bytes // 4, dtype=torch.float32, device=dev) if need_bwd else None
        gates = torch.empty(z.gates_bytes // 4, dtype=torch.float32, device=dev) if need_bwd else None
        ws = torch.empty(int(z.fwd_workspace_bytes), dtype=torch.uint8, device=dev)
        _call('nt_lstm_fwd', lib.nt_lstm_fwd, _p(x), ldx, _p(h0), _p(c0), _p(w), R, T, L, H, E, _p(y), _p(act), _p(cs), _p(gates),
              _p(ws), _stream())
        if FLOP_SINK is not None:
            FLOP_SINK['nt_lstm_fwd'] = FLOP_SINK.get('nt_lstm_fwd', 0.0) + 2.0 * R * T * 4 * H * sum((E if l == 0 else H) + H for l in range(L))
        ctx.dims = (R, T, L, H, E)
        if need_bwd:
            ctx.save_for_backward(act, cs, gates, w)
        return y[:, :, :H]                            # [T, R, H] time-major view (row stride y_ld)

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        R, T, L, H, E = ctx.dims
        act, cs, gates, w = ctx.saved_tensors
        gout = gout.contiguous()
        z = _lstm_sizes(R, T, L, H, E)
        dev = gout.device
        ws = torch.empty(int(z.bwd_workspace_bytes), dtype=torch.uint8, device=dev)
        dx = torch.empty(R, E, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        f32 = dict(dtype=torch.float32, device=dev)
        dw_ih = [torch.empty(4 * H, E if l == 0 else H, **f32) for l in range(L)]
        dw_hh = [torch.empty(4 * H, H, **f32) for l in range(L)]
        db_ih = [torch.empty(4 * H, **f32) for l in range(L)]
        db_hh = [torch.empty(4 * H, **f32) for l in range(L)]
        grads = (None,) * 4 if _LSTM_SKIP_DW else (_ptr_array(dw_ih), _ptr_array(dw_hh), _ptr_array(db_ih), _ptr_array(db_hh))
        _call('nt_lstm_bwd', lib.nt_lstm_bwd, _p(gout), H, _p(act), _p(cs), _p(gates), _p(w), R, T, L, H, E, _p(ws), _p(dx), E,
              *grads, _stream())
        if FLOP_SINK is not None:
            FLOP_SINK['nt_lstm_bwd'] = FLOP_SINK.get('nt_lstm_bwd', 0.0) + 4.0 * R * T * 4 * H * sum((E if l == 0 else H) + H for l in range(L))
        flat = []
        for l in range(L):
            flat += [dw_ih[l], dw_hh[l], db_ih[l], db_hh[l]]
        return (dx, None, None, None, *flat)


def lstm_decoder(x, h0, c0, T, params):
    """nn.LSTM(batch_first=True) on the input x [R, E] repeated T times (LSTMDecoderModule, nn/net_blocks.py:382-402).
    params: flat list (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, weight_ih_l1, ...); h0, c0: [L, R, H].
    Returns the top layer's hidden states TIME-MAJOR, [T, R, H] (a strided view; row stride 256 floats)."""
    return _LSTMDecoderFunction.apply(x, h0, c0, int(T), *params)


# ----------------------------------------------------------------------------------------------------------
# PointNet++ set abstraction (csrc/pointnet.cu)
# ----------------------------------------------------------------------------------------------------------
def fps(pos, B, N, ratio):
    """torch_geometric.nn.fps on the dense equal-size layout with a deterministic start (nn/net_blocks.py:19): ceil(ratio * N)
    farthest points per cloud.  Returns LOCAL indices [B, n] int32."""
    import math
    _require_cuda(pos)
    pos, ld = _rows2d(pos)
    n = int(math.ceil(ratio * N))
    idx = torch.empty(B, n, dtype=torch.int32, device=pos.device)
    _call('nt_fps', _lib.load().nt_fps, _p(pos), ld, B, N, pos.shape[1], n, _p(idx), _stream())
    return idx


def radius(pos, B, N, centres, r, max_num_neighbors):
    """torch_geometric.nn.radius(pos, pos[idx], r, batch, batch[idx], max_num_neighbors) (nn/net_blocks.py:20-21) for centres
    given as LOCAL point indices [B, M].  Returns (nbr [B*M, max] int32 local, -1 padded; count [B*M] int32)."""
    _require_cuda(pos, centres)
    pos, ld = _rows2d(pos)
    M = centres.shape[1]
    nbr = torch.empty(B * M, max_num_neighbors, dtype=torch.int32, device=pos.device)
    cnt = torch.empty(B * M, dtype=torch.int32, device=pos.device)
    _call('nt_radius', _lib.load().nt_radius, _p(pos), ld, B, N, pos.shape[1], _p(centres.contiguous()), M, float(r),
          int(max_num_neighbors), _p(nbr), _p(cnt), _stream())
    return nbr, cnt


def point_edges(pos, B, N, centres, nbr, cnt):
    """Edge list + message input of PointConv in the reference's bipartite call (see include/nt_b200.h).  One host read (the
    edge count decides the size of the result, as in the library).  Returns (src [E], dst [E] int64 global rows, msg [E, D])."""
    pos, ld = _rows2d(pos)
    D, M, max_nbr = pos.shape[1], centres.shape[1], nbr.shape[1]
    lib = _lib.load()
    keep = torch.empty(B * M, dtype=torch.int32, device=pos.device)
    _call('nt_point_edges_count', lib.nt_point_edges_count, _p(nbr), _p(cnt), B, N, M, max_nbr, _p(keep), _stream())
    ends = torch.cumsum(keep.long(), 0)
    offsets = (ends - keep.long()).contiguous()
    n_radius = int(ends[-1].item())
    E = n_radius + min(B * N, B * M)
    src = torch.empty(E, dtype=torch.int64, device=pos.device)
    dst = torch.empty(E, dtype=torch.int64, device=pos.device)
    msg = torch.empty(E, D, dtype=torch.float32, device=pos.device)
    _call('nt_point_edges_fill', lib.nt_point_edges_fill, _p(pos), ld, D, _p(centres.contiguous()), _p(nbr), _p(cnt), _p(offsets), B, N, M,
          max_nbr, n_radius, _p(src), _p(dst), _p(msg), D, _stream())
    return src, dst, msg


class _ScatterMaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, dst, T):
        _require_cuda(v, dst)
        v, ldv = _rows2d(v)
        E, F = v.shape
        out = torch.empty(T, F, dtype=torch.float32, device=v.device)
        arg = torch.empty(T, F, dtype=torch.int64, device=v.device)
        key = torch.empty(T, F, dtype=torch.int32, device=v.device)
        _call('nt_scatter_max_fwd', _lib.load().nt_scatter_max_fwd, _p(v), ldv, _p(dst), E, F, T, _p(out), _p(arg), _p(key), _stream())
        ctx.save_for_backward(arg)
        ctx.dims = (E, F, T)
        return out

    @staticmethod
    def backward(ctx, g):
        arg, = ctx.saved_tensors
        E, F, T = ctx.dims
        gv = torch.zeros(E, F, dtype=torch.float32, device=g.device)
        _call('nt_scatter_max_bwd', _lib.load().nt_scatter_max_bwd, _p(g.contiguous()), _p(arg), T, F, _p(gv), F, _stream())
        return gv, None, None


def scatter_max(values, dst, n_targets):
    """max aggregation of per-edge rows `values` [E, F] into n_targets rows by the int64 target index `dst` (PyG aggr='max')."""
    return _ScatterMaxFunction.apply(values, dst.contiguous(), int(n_targets))


def all_edge_pairs(edges, num_edges):
    """Device batching of ``NNSewingPattern.all_edge_pairs`` (nn/data/pattern_converter.py:458-499): every pair of 3D edges that
    belong to DIFFERENT panels, in the reference's order (panel i, panel j > i, edges of i x edges of j row-major).
    edges: [P, Lmax, F] fp32 on the device; num_edges: the per-panel edge counts (host list / tensor).  Returns (pairs
    [n_pairs, 2F], mapping [n_pairs, 4] int32 = (panel i, edge, panel j, edge)) -- the input of StitchOnEdge3DPairs and the table
    ``stitches_from_pair_classifier`` indexes."""
    _require_cuda(edges)
    edges = edges.contiguous().float()
    P, Lmax, F = edges.shape
    counts = [int(c) for c in (num_edges.tolist() if torch.is_tensor(num_edges) else num_edges)]
    bi, bj, bc, off, total = [], [], [], [], 0
    for i in range(P):
        for j in range(i + 1, P):
            if counts[i] > 0 and counts[j] > 0:
                bi.append(i); bj.append(j); bc.append(counts[j]); off.append(total)
                total += counts[i] * counts[j]
    if total == 0:
        raise ValueError('No edges to construct')            # InvalidPatternDefError in the reference (pattern_converter.py:494)
    dev = edges.device
    t = lambda v, dt: torch.tensor(v, dtype=dt, device=dev)
    bi_t, bj_t, bc_t, off_t = t(bi, torch.int32), t(bj, torch.int32), t(bc, torch.int32), t(off, torch.int64)
    pairs = torch.empty(total, 2 * F, dtype=torch.float32, device=dev)
    mapping = torch.empty(total, 4, dtype=torch.int32, device=dev)
    _call('nt_edge_pairs', _lib.load().nt_edge_pairs, _p(edges), Lmax, F, _p(bi_t), _p(bj_t), _p(bc_t), _p(off_t), len(bi), total,
          _p(pairs), _p(mapping), _stream())
    return pairs, mapping


# ----------------------------------------------------------------------------------------------------------
# pattern loss (shape / loop / rotation / translation) and Adam on a flat buffer (csrc/train_step.cu)
# ----------------------------------------------------------------------------------------------------------
class _PatternLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outl, rot, tr, gt_outl, gt_rot, gt_tr, num_edges, cfg):
        _require_cuda(outl, rot, tr, gt_outl, gt_rot, gt_tr, num_edges)
        for t in (outl, rot, tr):
            if t.dtype != torch.float32 or t.stride(-1) != 1:
                raise RuntimeError('pattern_loss: predictions must be fp32 with a dense last dimension')
        B, P, Lp, D = outl.shape
        a = _lib.PatternLossArgs()
        a.outlines, a.outl_stride_b, a.outl_stride_p, a.outl_stride_e = outl.data_ptr(), outl.stride(0), outl.stride(1), outl.stride(2)
        a.rotations, a.rot_stride_b, a.rot_stride_p = rot.data_ptr(), rot.stride(0), rot.stride(1)
        a.translations, a.tr_stride_b, a.tr_stride_p = tr.data_ptr(), tr.stride(0), tr.stride(1)
        gt_outl, gt_rot, gt_tr = gt_outl.contiguous().float(), gt_rot.contiguous().float(), gt_tr.contiguous().float()
        num_edges = num_edges.contiguous().long()
        a.gt_outlines, a.gt_rotations, a.gt_translations, a.num_edges = (gt_outl.data_ptr(), gt_rot.data_ptr(), gt_tr.data_ptr(),
                                                                           num_edges.data_ptr())
        a.B, a.P, a.Lp, a.D, a.Dr, a.Dt = B, P, Lp, D, rot.shape[-1], tr.shape[-1]
        a.pad_x, a.pad_y, a.loop_weight = cfg['pad_x'], cfg['pad_y'], cfg['loop_weight']
        a.use_shape, a.use_loop, a.use_rotation, a.use_translation = (int(cfg[k]) for k in ('shape', 'loop', 'rotation', 'translation'))
        acc = torch.empty(5, dtype=torch.float64, device=outl.device)
        out = torch.empty(5, dtype=torch.float32, device=outl.device)
        _call('nt_pattern_loss_fwd', _lib.load().nt_pattern_loss_fwd, ctypes.byref(a), _p(acc), _p(out), _stream())
        ctx.args = a
        ctx.save_for_backward(outl, rot, tr, gt_outl, gt_rot, gt_tr, num_edges)          # keeps the pointers in `a` alive
        return out

    @staticmethod
    def backward(ctx, g):
        outl, rot, tr = ctx.saved_tensors[:3]
        # only the total (element 0) carries gradient in the training step; the parts are reported values
        gscale = g[0:1].contiguous()
        g_outl = torch.empty(outl.shape, dtype=torch.float32, device=outl.device)
        g_rot = torch.empty(rot.shape, dtype=torch.float32, device=outl.device)
        g_tr = torch.empty(tr.shape, dtype=torch.float32, device=outl.device)
        _call('nt_pattern_loss_bwd', _lib.load().nt_pattern_loss_bwd, ctypes.byref(ctx.args), _p(gscale), _p(g_outl), _p(g_rot),
              _p(g_tr), _stream())
        return g_outl, g_rot, g_tr, None, None, None, None, None


def pattern_loss(outlines, rotations, translations, gt_outlines, gt_rotations, gt_translations, num_edges, components,
                 loop_weight=1.0, pad_xy=(0.0, 0.0)):
    """The main loss terms of ComposedPatternLoss (composed_loss.py:294-321) in one kernel.  Returns a [5] tensor:
    (total, pattern_loss, loop_loss, rotation_loss, translation_loss); gradients flow from element 0 only."""
    cfg = dict(shape='shape' in components, loop='loop' in components, rotation='rotation' in components,
               translation='translation' in components, loop_weight=float(loop_weight), pad_x=float(pad_xy[0]), pad_y=float(pad_xy[1]))
    return _PatternLossFunction.apply(outlines, rotations, translations, gt_outlines, gt_rotations, gt_translations, num_edges, cfg)


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, state, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0,
              zero_grad=True):
    """torch.optim.Adam on flat fp32 buffers; `lr` is a 1-element device tensor, `state` a 2-element device tensor (zeros at start)."""
    _require_cuda(params, grads, exp_avg, exp_avg_sq, lr, state)
    _call('nt_adam_step', _lib.load().nt_adam_step, _p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), _p(lr),
          float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(grad_scale), int(bool(zero_grad)), _p(state),
          _stream())


class _LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _require_cuda(x, weight)
        x, ldx = _rows2d(x)
        rows, K = x.shape
        n_out = weight.shape[0]
        w = weight.contiguous()
        out = torch.empty(rows, n_out, dtype=torch.float32, device=x.device)
        gemm_nt(rows, K, n_out, w, K, NT_EPI_BIAS, a=x, lda=ldx, bias=None if bias is None else bias.contiguous(),
                out=out, ldo=n_out)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g, ldg = _rows2d(g)
        rows, K = x.shape
        n_out = w.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(rows, K, dtype=torch.float32, device=x.device)
            wt = w.t().contiguous()
            gemm_nt(rows, n_out, K, wt, n_out, NT_EPI_BIAS, a=g, lda=ldg, out=gx, ldo=K, grad_gemm=True)
        if ctx.needs_input_grad[1]:
            gw = torch.zeros(n_out, K, dtype=torch.float32, device=x.device)
            gemm_tn(g, ldg, n_out, rows, gw, b=x, ldb=x.stride(0) if rows > 1 else K, n=K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            sums = torch.zeros(n_out, dtype=torch.float64, device=x.device)
            _call('nt_bn_bwd_reduce', _lib.load().nt_bn_bwd_reduce, _p(g), ldg, None, 0, None, None, rows, n_out, _p(sums), _stream())
            gb = sums.float()
        return gx, gw, gb


def linear(x, weight, bias=None):
    """nn.Linear on a [rows, K] tensor through the library's GEMM."""
    lead = x.shape[:-1]
    out = _LinearFunction.apply(x.reshape(-1, x.shape[-1]), weight, bias)
    return out.view(*lead, weight.shape[0])
