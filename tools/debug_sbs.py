import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
torch.backends.cudnn.allow_tf32 = False
from oracle import model as om
from oracle import thirdparty as tp
import garment_pattern_estimation_b200 as g
from helpers import rel_err
from test_gpu_model import _configs, _build
dev = torch.device('cuda:0')
dc, nc, lc = _configs()
torch.manual_seed(21)
oracle = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc)).to(dev)
mine = _build(21, dev)
mine.load_state_dict(oracle.state_dict())
B, N = 4, 512
x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(2)).to(dev)
gt = om.synthetic_ground_truth(B, seed=5, device=dev)
torch.manual_seed(7)
h0, c0 = om.init_state(3, B * 23, 250).to(dev), om.init_state(3, B * 23, 250).to(dev)
oracle.train(); mine.train()
# capture oracle graphs
graphs = []
orig = tp.knn_graph
def spy(x_, b_, k_):
    r = orig(x_, b_, k_); graphs.append(r); return r
tp.knn_graph = spy
o1 = oracle(x, lstm_state=(h0, c0))
o2 = mine(x, lstm_state=(h0, c0))
for li, conv in enumerate(mine.feature_extractor.conv_layers):
    mi = conv.last_index.long() + (torch.arange(B*N, device=dev)//N*N).unsqueeze(1)
    print('layer', li, 'graph mismatches', int((mi != graphs[li]).sum()))
l1, _ = om.main_losses(o1, gt); l2, _, _ = mine.loss(o2, gt)
l1.backward(); l2.backward()
g1, g2 = dict(oracle.named_parameters()), dict(mine.named_parameters())
for n in g1:
    if g1[n].grad is None: continue
    a, b = g2[n].grad, g1[n].grad
    e = rel_err(a, b)
    d = (a-b).abs()
    flat = d.reshape(-1)
    top = flat.topk(min(3, flat.numel()))
    print('%-55s rel %.2e  l2rel %.2e  maxabs %.2e refmax %.2e' % (n, e, float((a-b).norm()/b.norm()), float(d.max()), float(b.abs().max())))
    if e > 3e-3 and a.dim()==2:
        rows = d.max(dim=1).values; cols = d.max(dim=0).values
        print('    worst rows', rows.topk(3), 'worst cols', cols.topk(3))

# ---- fp64 ground truth: oracle in double, SAME graphs (replay captured fp32 graphs)
import copy
o64 = copy.deepcopy(oracle).double()
for p in o64.parameters(): p.grad = None
it = iter(graphs)
tp.knn_graph = lambda x_, b_, k_: next(it)
o3 = o64(x.double(), lstm_state=(h0.double(), c0.double()))
l3, _ = om.main_losses(o3, {k: (v.double() if v.dtype.is_floating_point else v) for k, v in gt.items()})
l3.backward()
g3 = dict(o64.named_parameters())
print('---- vs fp64 truth: (mine, oracle32)')
for n in g1:
    if g1[n].grad is None or 'conv_layers' not in n: continue
    t = g3[n].grad
    em = float((g2[n].grad.double()-t).abs().max()/t.abs().max()); eo = float((g1[n].grad.double()-t).abs().max()/t.abs().max())
    print('%-55s mine %.2e  oracle32 %.2e' % (n, em, eo))
