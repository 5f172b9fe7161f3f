"""GPU parity tests of the individual operators (call through the C ABI via ops.py).

kNN: bit-exact against the oracle C restatement and the committed known-answer fixtures.
Float operators: within 1e-3 relative (tests/helpers.py REL_TOL) of a plain-PyTorch fp32 reference of the same op.
"""
import os

import pytest
import torch

from helpers import REL_TOL, assert_close, assert_grad_close, global_index, ref_edgeconv, rel_err, torch_mlp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda_device):
    from garment_pattern_estimation_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------------------------
# kNN
# ------------------------------------------------------------------------------------------------------------
def _knn_gpu(ops, x, k, dev):
    B, N, D = x.shape
    return ops.knn_graph(x.to(dev).reshape(B * N, D), B, N, k).view(B, N, k).cpu()


def test_knn_known_answers(ops, cuda_device, golden_dir):
    cases = torch.load(os.path.join(golden_dir, 'knn_kat.pt'))
    for name, c in cases.items():
        got = _knn_gpu(ops, c['x'], c['k'], cuda_device)
        assert torch.equal(got, c['idx']), 'kNN indices differ from the golden vector: ' + name


@pytest.mark.parametrize('B,N,D,k', [(1, 5, 3, 5), (2, 33, 3, 4), (3, 1000, 3, 5), (2, 777, 150, 5), (2, 1024, 150, 16),
                                      (1, 300, 64, 32), (2, 257, 153, 8), (4, 50, 1, 3)])
def test_knn_bit_exact_vs_oracle(ops, cuda_device, B, N, D, k):
    from oracle import knn as oknn
    x = torch.randn(B, N, D, generator=torch.Generator().manual_seed(B * 1000 + N + D))
    got = _knn_gpu(ops, x, k, cuda_device)
    want = oknn.knn_indices(x, k, nthreads=8)
    assert torch.equal(got, want)


def test_knn_quantised_ties(ops, cuda_device):
    """coordinates on a coarse grid -> many exactly equal distances; order must be (distance, index)."""
    from oracle import knn as oknn
    x = torch.randint(0, 4, (2, 400, 3), generator=torch.Generator().manual_seed(5)).float()
    assert torch.equal(_knn_gpu(ops, x, 8, cuda_device), oknn.knn_indices(x, 8))


def test_knn_strided_input(ops, cuda_device):
    """feature matrix embedded in a wider buffer (row stride > D)."""
    from oracle import knn as oknn
    B, N, D = 2, 200, 150
    wide = torch.randn(B * N, 160, generator=torch.Generator().manual_seed(3)).to(cuda_device)
    got = ops.knn_graph(wide[:, :D], B, N, 5).view(B, N, 5).cpu()
    want = oknn.knn_indices(wide[:, :D].cpu().reshape(B, N, D), 5)
    assert torch.equal(got, want)


def test_knn_full_size_properties(ops, cuda_device):
    """BASELINE C2 shape (N=2048, 150-d): bit-exact on 2 clouds, plus size-independent properties on 8."""
    from oracle import knn as oknn
    B, N, D, k = 8, 2048, 150, 5
    x = torch.randn(B, N, D, generator=torch.Generator().manual_seed(77))
    got = _knn_gpu(ops, x, k, cuda_device)
    assert torch.equal(got[:2], oknn.knn_indices(x[:2], k, nthreads=8))
    assert torch.equal(got[:, :, 0], torch.arange(N).expand(B, N).int())          # self is the nearest (distance 0)
    assert int(got.min()) >= 0 and int(got.max()) < N
    srt = got.sort(dim=-1).values
    assert bool((srt[..., 1:] != srt[..., :-1]).all())                            # no neighbour listed twice
    d = (x.unsqueeze(2) - torch.gather(x.unsqueeze(1).expand(B, N, N, D), 2,
                                        got.long().unsqueeze(-1).expand(B, N, k, D))).pow(2).sum(-1)
    assert bool((d[..., 1:] >= d[..., :-1] - 1e-3).all())                         # ascending distances


def test_knn_large_cloud_c4_shape(ops, cuda_device):
    """BASELINE C4 shape: N = 10 000 points, k = 16 (one cloud in 150-d, two in 3-d), bit-exact against the oracle."""
    from oracle import knn as oknn
    x3 = torch.randn(2, 10000, 3, generator=torch.Generator().manual_seed(10))
    assert torch.equal(_knn_gpu(ops, x3, 16, cuda_device), oknn.knn_indices(x3, 16, nthreads=os.cpu_count() or 8))
    x150 = torch.randn(1, 10000, 150, generator=torch.Generator().manual_seed(11))
    assert torch.equal(_knn_gpu(ops, x150, 16, cuda_device), oknn.knn_indices(x150, 16, nthreads=os.cpu_count() or 8))


def test_knn_clustered_data_exercises_pruning(ops, cuda_device):
    """tight clusters far apart: most candidates are pruned after the first dims; results must stay bit-exact."""
    from oracle import knn as oknn
    g = torch.Generator().manual_seed(12)
    centres = torch.randn(16, 150, generator=g) * 20
    x = (centres[torch.randint(0, 16, (2, 1500), generator=g)] + torch.randn(2, 1500, 150, generator=g) * 0.05)
    assert torch.equal(_knn_gpu(ops, x, 5, cuda_device), oknn.knn_indices(x, 5, nthreads=8))
    x[:, 100:140] = x[:, 0:40]                                  # exact duplicates across tiles
    assert torch.equal(_knn_gpu(ops, x, 8, cuda_device), oknn.knn_indices(x, 8, nthreads=8))


# ---- tensor-core filter + exact re-rank path (csrc/knn_tc.cu; taken for 8 <= D <= 157, N >= 200) -----------------------
@pytest.mark.parametrize('B,N,D,k', [(2, 2048, 150, 5), (3, 640, 157, 5), (2, 1000, 112, 5), (2, 500, 8, 3), (1, 3000, 150, 16),
                                      (2, 384, 150, 32), (5, 200, 33, 1), (1, 1153, 96, 9)])
def test_knn_tensor_core_filter_bit_exact(ops, cuda_device, B, N, D, k):
    from oracle import knn as oknn
    x = torch.randn(B, N, D, generator=torch.Generator().manual_seed(7 * B + N + D + k))
    assert torch.equal(_knn_gpu(ops, x, k, cuda_device), oknn.knn_indices(x, k, nthreads=8))


def test_knn_tensor_core_filter_hard_inputs(ops, cuda_device):
    """inputs on which the bf16 filter cannot separate the candidates: the exact fallback must take over, bit-exact."""
    from oracle import knn as oknn
    g = torch.Generator().manual_seed(21)
    # (a) a large common offset: norms ~1.5e8, distances ~3e2 -> every candidate is inside the error band
    x = 1000.0 + torch.randn(2, 400, 150, generator=g)
    assert torch.equal(_knn_gpu(ops, x, 5, cuda_device), oknn.knn_indices(x, 5, nthreads=8))
    # (b) masses of exact duplicates (ties on distance 0 far beyond k) and a few distinct points
    base = torch.randn(7, 150, generator=g)
    x = base[torch.randint(0, 7, (2, 600), generator=g)].clone()
    x[:, ::50] += torch.randn(2, 12, 150, generator=g)
    assert torch.equal(_knn_gpu(ops, x, 8, cuda_device), oknn.knn_indices(x, 8, nthreads=8))
    # (c) spatially ORDERED points (every new candidate is nearer than the previous ones for a while) and outliers
    t = torch.linspace(0, 1, 1500).view(1, 1500, 1)
    x = torch.cat([t * torch.randn(1, 1, 150, generator=g) * 30, ], dim=-1) + 0.01 * torch.randn(1, 1500, 150, generator=g)
    x[0, 700] *= 50
    assert torch.equal(_knn_gpu(ops, x, 5, cuda_device), oknn.knn_indices(x, 5, nthreads=8))
    # (d) tiny magnitudes and an all-zero cloud
    x = 1e-20 * torch.randn(1, 300, 64, generator=g)
    assert torch.equal(_knn_gpu(ops, x, 4, cuda_device), oknn.knn_indices(x, 4, nthreads=8))
    x = torch.zeros(1, 256, 16)
    assert torch.equal(_knn_gpu(ops, x, 6, cuda_device), oknn.knn_indices(x, 6, nthreads=8))


def test_knn_tensor_core_filter_on_model_features(ops, cuda_device):
    """layer-2 graph on real EdgeConv features (shipped-checkpoint scale: norms ~1e3, outliers ~3e4)."""
    from oracle import knn as oknn
    g = torch.Generator().manual_seed(33)
    feat = torch.relu(torch.randn(2, 2048, 150, generator=g)) * torch.rand(1, 1, 150, generator=g) * 8
    feat[:, ::97] *= 6.0
    assert torch.equal(_knn_gpu(ops, feat, 5, cuda_device), oknn.knn_indices(feat, 5, nthreads=8))


def test_knn_rejects_bad_k(ops, cuda_device):
    x = torch.randn(40, 3, device=cuda_device)
    with pytest.raises(RuntimeError):
        ops.knn_graph(x, 1, 40, 33)
    with pytest.raises(RuntimeError):
        ops.knn_graph(x, 4, 10, 11)


# ------------------------------------------------------------------------------------------------------------
# fused MLP: edge rows (DynamicEdgeConv) and plain rows
# ------------------------------------------------------------------------------------------------------------
def _copy_mlp(dst_mlp, src_mlp):
    dst_mlp.load_state_dict(src_mlp.state_dict())


@pytest.mark.parametrize('C,widths,k,B,N,tail', [
    (3, [200, 200, 150], 5, 2, 300, 0),
    (150, [200, 200, 150], 5, 2, 257, 3),
    (6, [32, 24], 4, 3, 64, 0),            # two Linear layers (EConv_hidden_depth = 1)
    (5, [16, 48, 40, 20], 16, 1, 130, 0),  # four Linear layers, k = 16
])
@pytest.mark.parametrize('training', [True, False])
def test_edgeconv_matches_torch(ops, cuda_device, C, widths, k, B, N, tail, training):
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    torch.manual_seed(11)
    ref_mlp = torch_mlp([2 * C] + widths).to(dev)
    with torch.no_grad():       # non-trivial BN affine incl. negative gammas (the shipped ckpt has 5.5 % negative)
        for blk in ref_mlp:
            blk[2].weight.copy_(torch.randn_like(blk[2].weight))
            blk[2].bias.copy_(torch.randn_like(blk[2].bias) * 0.3)
            blk[2].running_mean.copy_(torch.rand_like(blk[2].running_mean))
            blk[2].running_var.copy_(torch.rand_like(blk[2].running_var) + 0.5)
    mine = nb.DynamicEdgeConv(nb.MLP([2 * C] + widths), k=k).to(dev)
    _copy_mlp(mine.nn, ref_mlp)
    ref_mlp.train(training)
    mine.train(training)

    x = torch.randn(B * N, C, device=dev)
    pos = torch.randn(B * N, 3, device=dev) if tail else None
    x1 = x.clone().requires_grad_(training)
    x2 = x.clone().requires_grad_(training)
    out = mine(x1, cloud_shape=(B, N), tail_src=pos)
    idx = global_index(mine.last_index, N)
    want = ref_edgeconv(x2, idx, ref_mlp)
    if tail:
        want = torch.cat([want, pos], dim=-1)
    assert out.shape == want.shape
    assert_close(out, want, what='edgeconv forward')
    if not training:
        return
    g = torch.randn_like(want)
    out.backward(g)
    want.backward(g)
    assert_grad_close(x1.grad, x2.grad, what='grad wrt input features')
    for (n1, p1), (n2, p2) in zip(mine.nn.named_parameters(), ref_mlp.named_parameters()):
        assert n1 == n2
        assert_grad_close(p1.grad, p2.grad, what='grad ' + n1)
    for (n1, b1), (n2, b2) in zip(mine.nn.named_buffers(), ref_mlp.named_buffers()):
        if b1.dtype.is_floating_point:
            assert_close(b1, b2, what='BN buffer ' + n1)
        else:
            assert torch.equal(b1, b2), n1


@pytest.mark.parametrize('C,widths,k,B,N,tail', [(3, [200, 200, 150], 5, 2, 300, 0), (150, [200, 200, 150], 5, 3, 257, 3),
                                                 (7, [24, 40, 18], 16, 2, 130, 0)])
def test_edgeconv_composite_entry_points_match_the_kernel_by_kernel_path(ops, cuda_device, C, widths, k, B, N, tail):
    """nt_edgeconv_train_fwd / _bwd (the whole training-mode EdgeConv layer per call, orchestrated in C++: csrc/edgeconv_train.cu)
    against the kernel-by-kernel orchestration of the same kernels in ops.py: outputs, every gradient, BatchNorm buffers."""
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    res = {}
    for composite in (True, False):
        torch.manual_seed(5)
        conv = nb.DynamicEdgeConv(nb.MLP([2 * C] + widths), k=k).to(dev).train()
        with torch.no_grad():
            for blk in conv.nn:
                blk[2].weight.copy_(torch.randn_like(blk[2].weight))
                blk[2].bias.copy_(torch.randn_like(blk[2].bias) * 0.3)
        g = torch.Generator().manual_seed(9)
        x = torch.randn(B * N, C, generator=g).to(dev).requires_grad_(True)
        pos = torch.randn(B * N, 3, generator=g).to(dev).requires_grad_(True) if tail else None
        gout = torch.randn(B * N, widths[-1] + tail, generator=g).to(dev)
        ops.EDGECONV_COMPOSITE = composite
        before = ops._lib.launch_count()
        try:
            out = conv(x, cloud_shape=(B, N), tail_src=pos)
            out.backward(gout)
        finally:
            ops.EDGECONV_COMPOSITE = True
        torch.cuda.synchronize()
        res[composite] = dict(out=out.detach(), gx=x.grad, gpos=None if pos is None else pos.grad, launches=ops._lib.launch_count() - before,
                              grads={n: p.grad for n, p in conv.nn.named_parameters()}, bufs=dict(conv.nn.named_buffers()))
    a, b = res[True], res[False]
    assert a['launches'] > 0 and abs(a['launches'] - b['launches']) <= 4      # same kernels (+ the three small glue kernels)
    assert rel_err(a['out'], b['out']) < 1e-6
    assert rel_err(a['gx'], b['gx']) < 1e-5
    if tail:
        assert torch.equal(a['gpos'], b['gpos'])
    for n in a['grads']:
        assert rel_err(a['grads'][n], b['grads'][n]) < 1e-5, n
    for n in a['bufs']:
        if a['bufs'][n].dtype.is_floating_point:
            assert rel_err(a['bufs'][n], b['bufs'][n]) < 1e-6, n
        else:
            assert torch.equal(a['bufs'][n], b['bufs'][n]), n


@pytest.mark.parametrize('widths,rows', [([153, 153, 153, 23], 1000), ([16, 200, 200, 200, 1], 333), ([7, 9, 5], 64)])
@pytest.mark.parametrize('training', [True, False])
def test_plain_mlp_matches_torch(ops, cuda_device, widths, rows, training):
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    torch.manual_seed(3)
    ref_mlp = torch_mlp(widths).to(dev)
    with torch.no_grad():
        for blk in ref_mlp:
            blk[2].weight.copy_(torch.randn_like(blk[2].weight))
            blk[2].bias.copy_(torch.randn_like(blk[2].bias) * 0.3)
    mine = nb.MLP(widths).to(dev)
    _copy_mlp(mine, ref_mlp)
    ref_mlp.train(training)
    mine.train(training)
    x = torch.randn(rows, widths[0], device=dev)
    x1 = x.clone().requires_grad_(training)
    x2 = x.clone().requires_grad_(training)
    out, want = mine(x1), ref_mlp(x2)
    assert_close(out, want, what='mlp forward')
    if not training:
        return
    g = torch.randn_like(want)
    out.backward(g)
    want.backward(g)
    assert_grad_close(x1.grad, x2.grad, what='grad wrt input')
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref_mlp.named_parameters()):
        assert_grad_close(p1.grad, p2.grad, what='grad ' + n1)
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ref_mlp.named_buffers()):
        if b1.dtype.is_floating_point:
            assert_close(b1, b2, what='BN buffer ' + n1)


def test_eval_mode_backward_is_refused(ops, cuda_device):
    from garment_pattern_estimation_b200 import net_blocks as nb
    mlp = nb.MLP([4, 8, 3]).to(cuda_device).eval()
    x = torch.randn(10, 4, device=cuda_device, requires_grad=True)
    with pytest.raises(RuntimeError):
        mlp(x).sum().backward()


# ------------------------------------------------------------------------------------------------------------
# sparsemax / attention pooling / linear
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('rows,P', [(1000, 23), (17, 1), (64, 32), (5, 2)])
def test_sparsemax_matches_published_algorithm(ops, cuda_device, rows, P):
    from oracle.thirdparty import Sparsemax
    z = (torch.randn(rows, P, generator=torch.Generator().manual_seed(rows)) * 2).to(cuda_device)
    z[0] = 0.5                                        # full tie row
    z1, z2 = z.clone().requires_grad_(True), z.clone().requires_grad_(True)
    out, want = ops.sparsemax(z1), Sparsemax(dim=1)(z2)
    assert_close(out, want, tol=1e-5, what='sparsemax forward')
    assert torch.allclose(out.sum(-1), torch.ones(rows, device=cuda_device), atol=1e-5)
    assert bool(((out == 0) == (want == 0)).all()), 'support set differs'
    g = torch.randn_like(out)
    out.backward(g)
    want.backward(g)
    assert_close(z1.grad, z2.grad, tol=1e-5, what='sparsemax backward')


@pytest.mark.parametrize('B,N,P,F', [(3, 300, 23, 153), (1, 64, 4, 300), (2, 129, 32, 7)])
def test_attention_pool_matches_reference_loop(ops, cuda_device, B, N, P, F):
    dev = cuda_device
    w = torch.rand(B * N, P, device=dev)
    f = torch.randn(B * N, F, device=dev)
    w1, f1 = w.clone().requires_grad_(True), f.clone().requires_grad_(True)
    w2, f2 = w.clone().requires_grad_(True), f.clone().requires_grad_(True)
    enc = ops.attention_pool(w1, f1, B, N, 1.0 / N)
    # reference order of operations: per panel, weights * features, mean over the cloud (nn/nets.py:263-276)
    want = torch.stack([(w2[:, p:p + 1] * f2).view(B, N, F).mean(dim=1) for p in range(P)], dim=1)
    assert_close(enc, want, what='attention pool forward')
    g = torch.randn_like(want)
    enc.backward(g)
    want.backward(g)
    assert_close(w1.grad, w2.grad, what='attention pool grad weights')
    assert_close(f1.grad, f2.grad, what='attention pool grad features')


@pytest.mark.parametrize('rows,K,n_out', [(46, 153, 250), (736, 250, 7), (1, 5, 3), (10304, 250, 8)])
def test_linear_matches_torch(ops, cuda_device, rows, K, n_out):
    dev = cuda_device
    lin = torch.nn.Linear(K, n_out).to(dev)
    x = torch.randn(rows, K, device=dev)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    w1, b1 = lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
    out = ops.linear(x1, w1, b1)
    want = lin(x2)
    assert_close(out, want, what='linear forward')
    g = torch.randn_like(want)
    out.backward(g)
    want.backward(g)
    assert_close(x1.grad, x2.grad, what='linear grad x')
    assert_close(w1.grad, lin.weight.grad, what='linear grad w')
    assert_close(b1.grad, lin.bias.grad, what='linear grad b')


def test_cpu_tensors_are_refused(ops):
    with pytest.raises(RuntimeError):
        ops.knn_graph(torch.randn(10, 3), 1, 10, 3)
    with pytest.raises(RuntimeError):
        ops.sparsemax(torch.randn(4, 5))


# ------------------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) engine vs float64 on identical operands
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('rows,K,n_out', [(1000, 200, 200), (257, 150, 400), (5000, 153, 23), (129, 3, 400),
                                          (640, 300, 200), (128, 32, 16), (77, 250, 7)])
def test_tc_engine_linear_matches_float64(ops, cuda_device, rows, K, n_out):
    dev = cuda_device
    g = torch.Generator(device='cpu').manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g).to(dev)
    w = (torch.randn(n_out, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(n_out, generator=g).to(dev)
    want = torch.nn.functional.linear(x.double(), w.double(), b.double())
    assert rel_err(ops.linear(x, w, b), want) < 5e-5, 'error-compensated 3-product split should be <= ~1e-5'


@pytest.mark.parametrize('rows,m,n', [(5000, 150, 200), (777, 200, 200), (300, 400, 3), (100000, 23, 153), (64, 8, 250),
                                      (4096, 400, 150), (500, 64, 10), (33, 300, 260), (1, 5, 5)])
def test_tc_weight_gradient_gemm_matches_float64(ops, cuda_device, rows, m, n):
    """out = A^T . B on the tensor cores (plain operands: MN-major BF16x3 engine of gemm_tn_mn.cu; deterministic split reduction) vs a
    float64 reference, plain and centred."""
    dev = cuda_device
    g = torch.Generator().manual_seed(rows + m)
    a = torch.randn(rows, m, generator=g).to(dev)
    b = (torch.randn(rows, n, generator=g) + 0.5).to(dev)
    mu = b.mean(0)
    want = a.double().t() @ b.double()
    want_c = a.double().t() @ (b.double() - mu.double())
    out = torch.zeros(m, n, device=dev)
    ops.gemm_tn(a, a.stride(0), m, rows, out, b=b, ldb=b.stride(0), n=n)
    out_c = torch.zeros(m, n, dtype=torch.float64, device=dev)
    ops.gemm_tn(a, a.stride(0), m, rows, out_c, b=b, ldb=b.stride(0), n=n, mu=mu)
    assert rel_err(out, want) < 2e-5
    assert rel_err(out_c, want_c) < 2e-5


@pytest.mark.parametrize('rows,m,n', [(3000, 150, 200), (40000, 200, 150), (2049, 6, 3)])
def test_tc_weight_gradient_gemm_padded_rows(ops, cuda_device, rows, m, n):
    """The same product on operands that sit in padded row buffers (stride = columns rounded up to 4 floats, the layout every
    intermediate of the EdgeConv has): rows go through the bulk-copy path, a partial last quad carries NaN padding that must not leak."""
    dev = cuda_device
    g = torch.Generator().manual_seed(rows)
    a = ops._rowbuf(rows, m, dev).fill_(float('nan'))
    b = ops._rowbuf(rows, n, dev).fill_(float('nan'))
    a[:, :m] = torch.randn(rows, m, generator=g).to(dev)
    b[:, :n] = (torch.randn(rows, n, generator=g) + 0.5).to(dev)
    mu = b[:, :n].mean(0)
    want = a[:, :m].double().t() @ b[:, :n].double()
    want_c = a[:, :m].double().t() @ (b[:, :n].double() - mu.double())
    out = torch.zeros(m, n, device=dev)
    ops.gemm_tn(a, a.stride(0), m, rows, out, b=b, ldb=b.stride(0), n=n)
    out_c = torch.zeros(m, n, dtype=torch.float64, device=dev)
    ops.gemm_tn(a, a.stride(0), m, rows, out_c, b=b, ldb=b.stride(0), n=n, mu=mu)
    assert rel_err(out, want) < 2e-5
    assert rel_err(out_c, want_c) < 2e-5


# ------------------------------------------------------------------------------------------------------------
# streaming engine (gemm_tc3.cu: persistent CTAs, cp.async operand ring, two epilogue groups) vs the one-tile-per-CTA engine
# ------------------------------------------------------------------------------------------------------------
def _set_engine(n):
    """nt_gemm_args.engine of every row GEMM the host layer launches (per call: the library keeps no engine state)."""
    from garment_pattern_estimation_b200 import ops as _ops
    _ops.NT_ENGINE = n


def _run_gemm_nt(ops, epi, a, w, K, n_out, k=5, aux=None, with_out=True):
    """One nt_gemm_nt call with every output of the given epilogue; returns a dict of result tensors."""
    from garment_pattern_estimation_b200 import _lib
    dev = a.device
    rows = a.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    g = torch.Generator().manual_seed(5)
    res = {}
    out = ops._rowbuf(rows, n_out, dev).fill_(-7.0) if with_out else None
    kw = dict(a=a, lda=a.stride(0), out=out, ldo=out.stride(0) if with_out else 0)
    if epi == _lib.NT_EPI_BIAS:
        kw['bias'] = torch.randn(n_out, generator=g).to(dev)
    elif epi == _lib.NT_EPI_RELU_STATS:
        kw['bias'] = torch.randn(n_out, generator=g).to(dev)
        kw['stats'] = res['stats'] = torch.zeros(2 * n_out, dtype=torch.float64, device=dev)
    elif epi == _lib.NT_EPI_RELU_MAXMIN:
        kw['bias'] = torch.randn(n_out, generator=g).to(dev)
        kw['stats'] = res['stats'] = torch.zeros(2 * n_out, dtype=torch.float64, device=dev)
        M = rows // k
        agg = (torch.empty(M, n_out, **f32), torch.empty(M, n_out, **f32),
               torch.empty(M, n_out, dtype=torch.uint8, device=dev), torch.empty(M, n_out, dtype=torch.uint8, device=dev))
        kw['agg'], kw['k_agg'] = agg, k
        res['vmax'], res['vmin'], res['imax'], res['imin'] = agg
    else:
        kw['aux'], kw['ldaux'] = aux, aux.stride(0)
        kw['k0'] = torch.randn(n_out, generator=g).to(dev) * 0.1
        kw['k1'] = torch.randn(n_out, generator=g).to(dev) * 0.1
        kw['mu'] = torch.randn(n_out, generator=g).to(dev)
        kw['colsum'] = res['colsum'] = torch.zeros(n_out, dtype=torch.float64, device=dev)
    ops.gemm_nt(rows, K, n_out, w, w.stride(0), epi, **kw)
    if with_out:
        res['out'] = out[:, :n_out]
    res['_kw'] = kw
    return res


@pytest.mark.parametrize('epi_name,rows,K,n_out', [
    ('bias', 70001, 200, 200), ('relu_stats', 60000, 200, 200), ('relu_stats', 148 * 128 * 2 + 77, 150, 200),
    ('relu_maxmin', 70000, 200, 150), ('relu_maxmin', 19 * 125 * 5, 200, 152), ('bnrelu_bwd', 66000, 150, 200),
    ('bnrelu_bwd', 50001, 200, 200), ('bias', 300, 16, 24), ('relu_stats', 129, 200, 23), ('bnrelu_bwd', 1000, 23, 153),
])
def test_streaming_engine_matches_one_tile_engine_and_float64(ops, cuda_device, epi_name, rows, K, n_out):
    from garment_pattern_estimation_b200 import _lib
    dev = cuda_device
    epi = {'bias': _lib.NT_EPI_BIAS, 'relu_stats': _lib.NT_EPI_RELU_STATS, 'relu_maxmin': _lib.NT_EPI_RELU_MAXMIN,
           'bnrelu_bwd': _lib.NT_EPI_BNRELU_BWD}[epi_name]
    g = torch.Generator().manual_seed(rows + 7 * K + n_out)
    a = ops._rowbuf(rows, K, dev)
    a.fill_(float('nan'))                                    # the padding columns must never reach the product
    a[:, :K] = torch.randn(rows, K, generator=g).to(dev)
    w = (torch.randn(n_out, K, generator=g) / K ** 0.5).to(dev)
    aux = None
    if epi_name == 'bnrelu_bwd':
        aux = ops._rowbuf(rows, n_out, dev)
        aux.fill_(float('nan'))
        aux[:, :n_out] = torch.relu(torch.randn(rows, n_out, generator=g)).to(dev)
    got = {}
    try:
        # engine 5 (aux rows of BNRELU_BWD through a shared-memory ring) is what engine 0 (auto) picks for large calls since round 2
        for engine in (1, 3, 4, 5, 6):
            _set_engine(engine)
            got[engine] = _run_gemm_nt(ops, epi, a, w, K, n_out, aux=aux)
            torch.cuda.synchronize()
    finally:
        _set_engine(0)
    one, stream = got[1], got[3]
    for key in ('out', 'vmax', 'vmin', 'imax', 'imin'):       # two row tiles per weight stage: same per-tile arithmetic
        if key in one:
            assert torch.equal(one[key], got[4][key]), key + ' (two tiles per stage)'
    for key in ('stats', 'colsum'):
        if key in one:
            assert rel_err(got[4][key], one[key]) < 1e-5, key + ' (two tiles per stage)'
    if 5 in got:
        for key in ('out', 'vmax', 'vmin', 'imax', 'imin'):
            if key in one:
                assert torch.equal(one[key], got[5][key]), key + ' (aux ring)'
        for key in ('stats', 'colsum'):
            if key in one:
                assert rel_err(got[5][key], one[key]) < 1e-5, key + ' (aux ring)'
    # second-generation streaming engine (gemm_tc4.cu: eight converter warps, TMA tensor-map epilogue)
    for key in ('out', 'vmax', 'vmin', 'imax', 'imin'):
        if key in one:
            assert torch.equal(one[key], got[6][key]), key + ' (TMA epilogue)'
    for key in ('stats', 'colsum'):
        if key in one:
            assert rel_err(got[6][key], one[key]) < 1e-5, key + ' (TMA epilogue)'
    # same operand split, same MMA order: element-wise results are bit-identical; the column statistics are accumulated
    # with atomics in a different order
    for key in ('out', 'vmax', 'vmin', 'imax', 'imin'):
        if key in one:
            assert torch.equal(one[key], stream[key]), key
    for key in ('stats', 'colsum'):
        if key in one:
            assert rel_err(stream[key], one[key]) < 1e-5, key
    # ... and against float64
    kw = stream['_kw']
    z = a[:, :K].double() @ w.double().t()
    if epi_name == 'bias':
        want = z + kw['bias'].double()
    elif epi_name in ('relu_stats', 'relu_maxmin'):
        want = torch.relu(z + kw['bias'].double())
        assert rel_err(stream['stats'][:n_out], want.sum(0)) < 1e-4
        assert rel_err(stream['stats'][n_out:], (want * want).sum(0)) < 1e-4
        if epi_name == 'relu_maxmin':
            v = want.view(rows // 5, 5, n_out)
            assert rel_err(stream['vmax'], v.max(1).values) < 5e-5
            assert rel_err(stream['vmin'], v.min(1).values) < 5e-5
    else:
        x = aux[:, :n_out].double()
        want = torch.where(x > 0, z - kw['k0'].double() - (x - kw['mu'].double()) * kw['k1'].double(), torch.zeros_like(z))
        assert rel_err(stream['colsum'], want.sum(0)) < 1e-4
    assert rel_err(stream['out'], want) < 5e-5
    assert torch.isnan(a[:, K:]).all() and torch.isfinite(stream['out']).all()


@pytest.mark.parametrize('engine', [0, 3, 4])
def test_edgeconv_large_clouds_through_the_streaming_engine(ops, cuda_device, engine):
    """EdgeConv layer-2 shape (C = 150, 200-200-150, k = 5) at 4 x 4096 points: 81 920 edge rows = 640 row tiles, i.e. several
    tiles per persistent CTA, forward + backward against plain PyTorch."""
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    C, widths, k, B, N = 150, [200, 200, 150], 5, 4, 4096
    torch.manual_seed(21)
    ref_mlp = torch_mlp([2 * C] + widths).to(dev)
    with torch.no_grad():
        for blk in ref_mlp:
            blk[2].weight.copy_(torch.randn_like(blk[2].weight))
            blk[2].bias.copy_(torch.randn_like(blk[2].bias) * 0.3)
    mine = nb.DynamicEdgeConv(nb.MLP([2 * C] + widths), k=k).to(dev)
    _copy_mlp(mine.nn, ref_mlp)
    x = torch.randn(B * N, C, device=dev)
    pos = torch.randn(B * N, 3, device=dev)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    try:
        _set_engine(engine)
        out = mine(x1, cloud_shape=(B, N), tail_src=pos)
        g = torch.randn_like(out)
        out.backward(g)
        torch.cuda.synchronize()
    finally:
        _set_engine(0)
    want = torch.cat([ref_edgeconv(x2, global_index(mine.last_index, N), ref_mlp), pos], dim=-1)
    want.backward(g)
    assert_close(out, want, what='edgeconv forward (streaming engine)')
    # 81 920 edge rows put ~30x more pre-activations within 1e-6 of a ReLU kink than the small cases above (see
    # helpers.assert_grad_close): same comparison, relative-L2 bound scaled to 5e-3 (measured 2.05e-3 on the BN bias gradient)
    assert_grad_close(x1.grad, x2.grad, what='grad wrt input features', l2_tol=5e-3)
    for (n1, p1), (n2, p2) in zip(mine.nn.named_parameters(), ref_mlp.named_parameters()):
        assert_grad_close(p1.grad, p2.grad, what='grad ' + n1, l2_tol=5e-3)
    for (n1, b1), (n2, b2) in zip(mine.nn.named_buffers(), ref_mlp.named_buffers()):
        if b1.dtype.is_floating_point:
            assert_close(b1, b2, what='BN buffer ' + n1)


@pytest.mark.parametrize('mode', ['mean', 'max', 'add'])
@pytest.mark.parametrize('B,N,F', [(3, 257, 153), (1, 5, 7), (32, 2048, 150)])
def test_global_pool_matches_torch(ops, cuda_device, mode, B, N, F):
    dev = cuda_device
    x = torch.randn(B * N, F, device=dev)
    if mode == 'max':
        x[1::7] = x[0]                       # exact ties: the first maximal point must receive the gradient (scatter-max)
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    got = ops.global_pool(x1, B, N, mode)
    v = x2.view(B, N, F)
    want = {'mean': lambda: v.mean(1), 'add': lambda: v.sum(1), 'max': lambda: v.max(1).values}[mode]()
    assert_close(got, want, tol=1e-5, what='global pool ' + mode)
    g = torch.randn_like(want)
    got.backward(g)
    if mode == 'max':
        first = (v == want.unsqueeze(1)).float().argmax(1)                   # first point attaining the maximum
        ref = torch.zeros_like(v).scatter_(1, first.unsqueeze(1), g.unsqueeze(1)).view(B * N, F)
    else:
        want.backward(g)
        ref = x2.grad
    assert_close(x1.grad, ref, tol=1e-6, what='global pool grad ' + mode)


def test_fused_edge_scatter_matches_separate_scatter_pass(ops, cuda_device):
    """EdgeConv backward with dz_1 scattered inside the epilogue of the last data-gradient GEMM (streaming engine) vs the
    separate nt_edge_scatter pass: same gradients up to the summation order of the atomics."""
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    C, widths, k, B, N = 150, [200, 200, 150], 5, 3, 1000
    torch.manual_seed(5)
    conv = nb.DynamicEdgeConv(nb.MLP([2 * C] + widths), k=k).to(dev).train()
    x = torch.randn(B * N, C, device=dev)
    g = torch.randn(B * N, widths[-1], device=dev)
    results = {}
    try:
        _set_engine(3)
        ops.EDGECONV_COMPOSITE = False        # both passes through the kernel-by-kernel orchestration (the launch counts are compared)
        for fused in (True, False):
            ops.FUSED_SCATTER = fused
            conv.zero_grad()
            xi = x.clone().requires_grad_(True)
            before = _launches()
            conv(xi, cloud_shape=(B, N)).backward(g)
            torch.cuda.synchronize()
            results[fused] = (xi.grad.clone(), {n: p.grad.clone() for n, p in conv.named_parameters()}, _launches() - before)
    finally:
        ops.FUSED_SCATTER = False
        ops.EDGECONV_COMPOSITE = True
        _set_engine(0)
    assert results[True][2] == results[False][2] - 1, 'the fused path must save exactly the nt_edge_scatter launch'
    assert rel_err(results[True][0], results[False][0]) < 1e-5
    for n in results[True][1]:
        assert rel_err(results[True][1][n], results[False][1][n]) < 1e-5, n


def _launches():
    from garment_pattern_estimation_b200 import _lib
    return _lib.launch_count()


# ------------------------------------------------------------------------------------------------------------
# fused inference EdgeConv (csrc/edgeconv_eval.cu, BASELINE.json north_star "single kernel" EdgeConv)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('C,widths,k,B,N,tail', [(150, [200, 200, 150], 5, 4, 2048, 3), (3, [200, 200, 150], 5, 3, 1000, 0),
                                                 (150, [200, 200, 150], 16, 2, 1500, 3), (8, [64, 48, 40], 7, 2, 333, 0)])
def test_fused_eval_edgeconv_matches_layerwise_path_and_torch(ops, cuda_device, C, widths, k, B, N, tail):
    """Eval-mode DynamicEdgeConv through the single fused kernel (gather -> GEMM -> GEMM -> max -> BN) against (a) the layer-by-layer
    kernels on the same kNN graph and (b) plain PyTorch with running-statistics BatchNorm (incl. negative gammas: max vs min)."""
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    torch.manual_seed(5)
    ref_mlp = torch_mlp([2 * C] + widths).to(dev)
    with torch.no_grad():
        for blk in ref_mlp:
            blk[2].weight.copy_(torch.randn_like(blk[2].weight))
            blk[2].bias.copy_(torch.randn_like(blk[2].bias) * 0.3)
            blk[2].running_mean.copy_(torch.rand_like(blk[2].running_mean))
            blk[2].running_var.copy_(torch.rand_like(blk[2].running_var) + 0.5)
    conv = nb.DynamicEdgeConv(nb.MLP([2 * C] + widths), k=k).to(dev).eval()
    _copy_mlp(conv.nn, ref_mlp)
    ref_mlp.eval()
    x = torch.randn(B * N, C, device=dev)
    pos = torch.randn(B * N, 3, device=dev) if tail else None
    before = _launches()
    with torch.no_grad():
        assert ops.EDGE_EVAL_FUSED
        fused = conv(x, cloud_shape=(B, N), tail_src=pos)
        n_fused = _launches() - before
        ops.EDGE_EVAL_FUSED = False
        try:
            layerwise = conv(x, cloud_shape=(B, N), tail_src=pos)
        finally:
            ops.EDGE_EVAL_FUSED = True
        want = ref_edgeconv(x, global_index(conv.last_index, N), ref_mlp)
    if tail:
        want = torch.cat([want, pos], dim=-1)
    assert fused.shape == want.shape
    assert rel_err(fused, layerwise) <= 1e-4, 'fused vs layer-by-layer {:.2e}'.format(rel_err(fused, layerwise))
    assert_close(fused, want, what='fused eval edgeconv vs torch')
    assert n_fused <= 12, 'the fused path should be kNN + PQ GEMM + 3 folds + 2 weight preps + 1 fused kernel, got {} launches'.format(n_fused)


def _launches():
    from garment_pattern_estimation_b200 import _lib
    return _lib.launch_count()
