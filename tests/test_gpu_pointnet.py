"""GPU parity of the PointNet++ extractor (csrc/pointnet.cu + net_blocks.PointNetPlusPlus) against the
oracle restatement of torch_geometric's fps / radius / PointConv and the golden produced by the UNMODIFIED reference
nn/net_blocks.py::PointNetPlusPlus (tests/golden/make_golden_pointnet.py).  Index results bit-exact; activations 1e-3 relative."""
import os

import pytest
import torch

from helpers import assert_close, assert_grad_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,N,ratio', [(3, 400, 0.2), (2, 2048, 0.2), (1, 10000, 0.05), (4, 33, 0.5)])
def test_fps_and_radius_bit_exact_vs_oracle(cuda_device, B, N, ratio):
    import math
    from garment_pattern_estimation_b200 import ops
    from oracle import knn as oknn
    dev = cuda_device
    pos = 0.35 * torch.randn(B, N, 3, generator=torch.Generator().manual_seed(N))
    pos[0, N // 2] = pos[0, N // 3]                                   # duplicate point: ties in both operators
    flat = pos.view(-1, 3).to(dev)
    idx = ops.fps(flat, B, N, ratio)
    n = int(math.ceil(ratio * N))
    assert idx.shape == (B, n)
    assert torch.equal(idx.cpu(), oknn.fps_indices(pos, n)), 'fps indices differ from the oracle'
    for r, mx in ((0.3, 25), (0.1, 4)):
        nbr, cnt = ops.radius(flat, B, N, idx, r, mx)
        onbr, ocnt = oknn.radius_neighbours(pos, idx.cpu(), r, mx)
        assert torch.equal(cnt.cpu().view(B, n), ocnt) and torch.equal(nbr.cpu().view(B, n, mx), onbr), (r, mx)


def test_fps_known_answer_and_limits(cuda_device):
    from garment_pattern_estimation_b200 import ops
    dev = cuda_device
    line = (torch.arange(12, dtype=torch.float32).view(12, 1) * torch.tensor([1., 0., 0.])).to(dev)
    assert ops.fps(line, 1, 12, 5 / 12).cpu().tolist() == [[0, 11, 5, 8, 2]]
    with pytest.raises(RuntimeError):
        ops.fps(torch.randn(10, 3), 1, 10, 0.5)                        # CPU tensors are refused


def test_point_edges_and_scatter_max(cuda_device):
    """Edge list of PointConv (radius edges minus index-equal pairs plus the appended i -> i edges) against the oracle's
    restatement of the library code, and scatter-max forward / backward against torch.scatter_reduce."""
    from garment_pattern_estimation_b200 import ops
    from oracle import thirdparty as tp
    dev = cuda_device
    B, N = 3, 400
    pos = 0.35 * torch.randn(B, N, 3, generator=torch.Generator().manual_seed(9))
    flat = pos.view(-1, 3).to(dev)
    idx = ops.fps(flat, B, N, 0.2)
    nbr, cnt = ops.radius(flat, B, N, idx, 0.3, 25)
    src, dst, msg = ops.point_edges(flat, B, N, idx, nbr, cnt)
    batch = torch.arange(B).repeat_interleave(N)
    cflat = pos.view(-1, 3)
    gidx = tp.fps(cflat, batch, ratio=0.2)
    row, col = tp.radius(cflat, cflat[gidx], 0.3, batch, batch[gidx], max_num_neighbors=25)
    keep = col != row
    loops = torch.arange(min(B * N, gidx.numel()))
    want_src, want_dst = torch.cat([col[keep], loops]), torch.cat([row[keep], loops])
    assert torch.equal(src.cpu(), want_src) and torch.equal(dst.cpu(), want_dst)
    want_msg = cflat[want_src] - cflat[gidx][want_dst]
    assert torch.equal(msg.cpu(), want_msg)
    v = torch.randn(src.numel(), 7, generator=torch.Generator().manual_seed(1)).to(dev)
    v[5] = v[9]                                                        # a tie inside one target's edge set is possible
    v1, v2 = v.clone().requires_grad_(True), v.clone().requires_grad_(True)
    T = gidx.numel()
    out = ops.scatter_max(v1, dst, T)
    ref = torch.full((T, 7), float('-inf'), device=dev).scatter_reduce(0, dst.view(-1, 1).expand(-1, 7), v2, reduce='amax')
    assert torch.equal(out, ref)
    g = torch.randn(T, 7, device=dev)
    out.backward(g)
    # reference gradient: to the FIRST edge attaining the maximum
    want = torch.zeros_like(v)
    for t in range(0, T, 17):
        e = (dst == t).nonzero().view(-1)
        for f in range(7):
            want[e[(v[e, f] == out[t, f]).nonzero()[0, 0]], f] = g[t, f]
    rows = torch.cat([(dst == t).nonzero().view(-1) for t in range(0, T, 17)])
    assert torch.equal(v1.grad[rows], want[rows])


def test_pointnet_plus_plus_matches_reference_golden(cuda_device, golden_dir):
    from garment_pattern_estimation_b200 import net_blocks as nb
    gold = torch.load(os.path.join(golden_dir, 'pointnet.pt'))
    dev = cuda_device
    model = nb.PointNetPlusPlus(gold['out_size'], dict(gold['config']))
    assert set(model.state_dict()) == set(gold['state'])                 # the reference's keys: sa1_module.conv.local_nn.*, ...
    model.load_state_dict(gold['state'], strict=True)
    model.to(dev).train()
    x = gold['x'].to(dev)
    y = model(x)
    assert_close(y, gold['train']['y'], what='PointNet++ train forward')
    y.backward(gold['train']['gout'].to(dev))
    for name, p in model.named_parameters():
        assert_grad_close(p.grad, gold['train']['grads'][name], what='PointNet++ grad ' + name)
    sd = model.state_dict()
    for k_, want in gold['train']['buffers'].items():
        if want.dtype.is_floating_point:
            assert_close(sd[k_], want, what='BN buffer ' + k_)
        else:
            assert int(sd[k_]) == int(want), k_
    model.load_state_dict(gold['state'])
    model.eval()
    with torch.no_grad():
        assert_close(model(x), gold['eval_y'], what='PointNet++ eval forward')
    src, dst = model.sa1_module.conv.last_edges
    assert src.numel() == gold['radius_row_col'].shape[1] - int((gold['radius_row_col'][0] == gold['radius_row_col'][1]).sum()) + 240


def test_all_edge_pairs_matches_the_reference_loop(cuda_device):
    """the edge-pair batching in front of StitchOnEdge3DPairs.  The reference builds the pairs panel pair
    by panel pair with numpy (nn/data/pattern_converter.py:458-499); restated here as that loop on the same per-panel edge arrays."""
    import numpy as np
    from garment_pattern_estimation_b200 import ops
    dev = cuda_device
    rng = np.random.default_rng(3)
    counts = [4, 0, 7, 3, 0, 14, 5]
    P, Lmax, F = len(counts), 14, 8
    edges = np.zeros((P, Lmax, F), dtype=np.float32)
    for p_, n in enumerate(counts):
        edges[p_, :n] = rng.standard_normal((n, F)).astype(np.float32)
    want_pairs, want_map = [], []
    for i in range(P):                                   # pattern_converter.py:471-490 (panels without edges contribute nothing)
        ei = edges[i, :counts[i]]
        for j in range(i + 1, P):
            ej = edges[j, :counts[j]]
            if len(ei) == 0 or len(ej) == 0:
                continue
            rows, cols = np.indices((len(ei), len(ej)))
            want_pairs.append(np.concatenate([ei[rows], ej[cols]], axis=-1).reshape(-1, 2 * F))
            want_map += [(i, r, j, c) for r in range(len(ei)) for c in range(len(ej))]
    want_pairs = np.concatenate(want_pairs)
    pairs, mapping = ops.all_edge_pairs(torch.from_numpy(edges).to(dev), counts)
    assert pairs.shape == want_pairs.shape == (len(want_map), 2 * F)
    assert np.array_equal(pairs.cpu().numpy(), want_pairs)
    assert mapping.cpu().tolist() == [list(m) for m in want_map]
    with pytest.raises(ValueError):
        ops.all_edge_pairs(torch.zeros(3, Lmax, F, device=dev), [0, 5, 0])
