"""Developer tool: the BASELINE.json configurations that are not the bench headline.
  C4: EdgeConv encoder fwd+bwd, B=16, N=10000, k=16 (HBM / kNN stress)
  C5: full-model inference (eval mode, BN running statistics), B=128, N in {1024, 2048, 4096, 8192}
Prints one JSON object; CUDA-event timings, 3 warm-up + 5 timed iterations each."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import garment_pattern_estimation_b200 as g  # noqa: E402
from garment_pattern_estimation_b200 import net_blocks as nb  # noqa: E402
from garment_pattern_estimation_b200 import configs as om  # noqa: E402  (config values)

dev = torch.device('cuda:0')


def timeit(fn, iters=5, warmup=3):
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


res = {}
# ---- C4
cfg = dict(om.ATT_NN_CONFIG)
cfg['k_neighbors'] = 16
torch.manual_seed(0)
enc = nb.EdgeConvFeatures(250, cfg).to(dev).train()
B, N = 16, 10000
pos = torch.randn(B, N, 3, device=dev)


def c4():
    for p in enc.parameters():
        p.grad = None
    _, feats, _ = enc(pos, False)
    feats.sum().backward()


ms = timeit(c4, iters=3, warmup=2)
res['C4_encoder_fwd_bwd'] = {'B': B, 'N': N, 'k': 16, 'ms': ms, 'clouds_per_s': B / (ms * 1e-3),
                             'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9}
del enc, pos
torch.cuda.empty_cache()

# ---- C5
lc = {'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [],
      'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}
torch.manual_seed(0)
model = g.GarmentSegmentPattern3D(dict(om.ATT_DATA_CONFIG), dict(om.ATT_NN_CONFIG), lc)
ck = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', '_ckpt', 'att_state.pt')
weights = 'random init'
if os.path.exists(ck):
    model.load_state_dict(torch.load(ck))
    weights = 'shipped neural_tailor_panels.pth'
model.to(dev).eval()
res['C5_inference'] = {'B': 128, 'weights': weights, 'sweep': {}}
for N in (1024, 2048, 4096, 8192):
    pos = torch.randn(128, N, 3, device=dev)

    def c5():
        with torch.no_grad():
            model(pos)

    ms = timeit(c5, iters=3, warmup=2)
    res['C5_inference']['sweep'][str(N)] = {'ms': ms, 'clouds_per_s': 128 / (ms * 1e-3)}
    del pos
print(json.dumps(res, indent=1))
