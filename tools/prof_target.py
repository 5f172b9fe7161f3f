"""Developer tool for ncu captures: runs a few launches of one hot kernel at the C2 shape (B=32, N=2048)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from garment_pattern_estimation_b200 import ops  # noqa: E402
from garment_pattern_estimation_b200 import net_blocks as nb  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'knn'
B, N, k = 32, 2048, 5
dev = torch.device('cuda:0')
torch.manual_seed(0)
if what == 'knn':
    x = torch.randn(B * N, 150, device=dev)
    for _ in range(3):
        ops.knn_graph(x, B, N, k)
elif what == 'edgeconv':
    conv = nb.DynamicEdgeConv(nb.MLP([300, 200, 200, 150]), k=k).to(dev).train()
    x = torch.randn(B * N, 150, device=dev, requires_grad=True)
    for _ in range(2):
        out = conv(x, cloud_shape=(B, N))
        out.sum().backward()
elif what == 'model':
    # two eager training steps of the attention model at the C2 shape (every kernel of the step, in launch order)
    import bench
    import garment_pattern_estimation_b200 as g
    dc, nc, lc = bench.att_configs(k)
    torch.manual_seed(bench.SEED_INIT)
    from garment_pattern_estimation_b200.parallel import FlatAdam, FlatDataParallel
    model = g.GarmentSegmentPattern3D(dc, nc, lc).to(dev).train()
    wrapper = FlatDataParallel(model, device_ids=[dev], auto_reduce=False)
    opt = FlatAdam(wrapper, lr=2e-3)
    x, gt = bench.synthetic_batch(B, N, seed=1234)
    x, gt = x.to(dev), {kk: v.to(dev) for kk, v in gt.items()}
    for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
        loss, _, _ = model.loss(wrapper(x), gt)
        loss.backward()
        opt.step(zero_grad=True)
elif what == 'infer':
    # one C5 inference batch (B=128, N=2048, eval mode): the fused EdgeConv kernel, the kNN kernels, the LSTM forward
    import bench
    import garment_pattern_estimation_b200 as g
    dc, nc, lc = bench.att_configs(k)
    torch.manual_seed(bench.SEED_INIT)
    model = g.GarmentSegmentPattern3D(dc, nc, lc).to(dev).eval()
    x = torch.randn(128, N, 3, device=dev)
    with torch.no_grad():
        for _ in range(2):
            model(x)
torch.cuda.synchronize()
