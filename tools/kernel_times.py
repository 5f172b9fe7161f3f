"""Developer tool: CUDA-event timings of the hot-path pieces at a BASELINE config (default C2: B=32, N=2048)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import garment_pattern_estimation_b200 as g  # noqa: E402
from garment_pattern_estimation_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=32)
    ap.add_argument('--N', type=int, default=2048)
    ap.add_argument('--k', type=int, default=5)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    B, N, k = args.B, args.N, args.k
    res = {}
    x3 = torch.randn(B * N, 3, device=dev)
    x150 = torch.randn(B * N, 150, device=dev)
    res['knn_d3_ms'] = timeit(lambda: ops.knn_graph(x3, B, N, k))
    res['knn_d150_ms'] = timeit(lambda: ops.knn_graph(x150, B, N, k))
    pair_dims = B * N * N * 150
    res['knn_d150_Tpairdims_per_s'] = pair_dims / (res['knn_d150_ms'] * 1e-3) / 1e12

    from garment_pattern_estimation_b200 import configs as om
    nc = dict(om.ATT_NN_CONFIG)
    nc['k_neighbors'] = k
    lc = {'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [],
          'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}
    torch.manual_seed(0)
    model = g.GarmentSegmentPattern3D(dict(om.ATT_DATA_CONFIG), nc, lc).to(dev).train()
    pos = torch.randn(B, N, 3, device=dev)
    import bench
    gt = {kk: v.to(dev) for kk, v in bench.synthetic_ground_truth(B, seed=11).items()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def fwd():
        with torch.no_grad():
            model(pos)

    def step():
        out = model(pos)
        loss, _, _ = model.loss(out, gt)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    def enc_only():
        model.feature_extractor(pos, False)[1].sum().backward()

    res['forward_nograd_ms'] = timeit(fwd)
    res['encoder_fwd_bwd_ms'] = timeit(enc_only)
    l0 = _lib.launch_count()
    res['train_step_ms'] = timeit(step, iters=5, warmup=2)
    res['launches_per_step'] = (_lib.launch_count() - l0) / 7
    res['clouds_per_s'] = B / (res['train_step_ms'] * 1e-3)
    res['mem_GB'] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
