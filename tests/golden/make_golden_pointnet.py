"""Golden fixture for SURVEY.md section 8 row a14 (PointNet++ extractor; run in the build container only):

    python tests/golden/make_golden_pointnet.py

tests/golden/pointnet.pt holds what the UNMODIFIED reference ``nn/net_blocks.py::PointNetPlusPlus`` (nn/net_blocks.py:10-88)
returns on top of the restated torch_geometric operators of oracle/thirdparty.py (fps with a deterministic start, radius,
PointConv): seeded weights, a train-mode forward + gradient digests + BatchNorm buffers, an eval-mode forward, the sampled centre
indices and the PointConv edge list, plus small known-answer cases for fps / radius.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import knn as oknn  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from oracle import thirdparty as tp  # noqa: E402


def main():
    _, nb = ref_stubs.import_reference()
    cfg = {'EConv_hidden': 64, 'EConv_feature': 48}
    torch.manual_seed(4321)
    model = nb.PointNetPlusPlus(40, dict(cfg))
    with torch.no_grad():       # non-trivial BatchNorm affine / running statistics
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(torch.randn_like(m.weight))
                m.bias.copy_(0.3 * torch.randn_like(m.bias))
                m.running_mean.copy_(0.2 * torch.rand_like(m.running_mean))
                m.running_var.copy_(0.5 + torch.rand_like(m.running_var))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    B, N = 3, 400
    x = 0.35 * torch.randn(B, N, 3, generator=torch.Generator().manual_seed(9))          # ~10-25 points inside r1 = 0.3
    out = {'config': cfg, 'out_size': 40, 'state': state, 'x': x}
    model.train()
    y = model(x)
    gout = torch.randn(y.shape, generator=torch.Generator().manual_seed(10))
    y.backward(gout)
    out['train'] = {'y': y.detach().clone(), 'gout': gout,
                    'grads': {n: p.grad.detach().clone() for n, p in model.named_parameters()},
                    'buffers': {k: v.clone() for k, v in model.state_dict().items() if 'running' in k or 'num_batches' in k}}
    model.load_state_dict(state)
    model.eval()
    with torch.no_grad():
        out['eval_y'] = model(x).clone()
    # intermediate integer results
    flat = x.view(-1, 3)
    batch = torch.arange(B).repeat_interleave(N)
    idx = tp.fps(flat, batch, ratio=0.2)
    row, col = tp.radius(flat, flat[idx], 0.3, batch, batch[idx], max_num_neighbors=25)
    out['fps_idx'] = idx
    out['radius_row_col'] = torch.stack([row, col])
    # known-answer cases: collinear equispaced points (ties), and a cloud with > max_nbr points inside the radius
    line = torch.arange(12, dtype=torch.float32).view(1, 12, 1) * torch.tensor([1., 0., 0.]).view(1, 1, 3)
    out['kat_fps_line'] = oknn.fps_indices(line, 5)
    dense = 0.05 * torch.randn(1, 60, 3, generator=torch.Generator().manual_seed(2))
    nbr, cnt = oknn.radius_neighbours(dense, torch.tensor([[0, 7]], dtype=torch.int32), 0.3, 25)
    out['kat_dense'] = {'pos': dense, 'nbr': nbr, 'cnt': cnt}
    torch.save(out, os.path.join(HERE, 'pointnet.pt'))
    print('fps', idx[:8].tolist(), 'edges', row.numel(), 'line', out['kat_fps_line'].tolist(), 'dense cnt', cnt.tolist())
    print('y', float(y.abs().mean()), 'eval', float(out['eval_y'].abs().mean()))


if __name__ == '__main__':
    main()
