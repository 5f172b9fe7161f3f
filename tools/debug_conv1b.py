import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
torch.backends.cudnn.allow_tf32 = False
from oracle import model as om
from oracle import thirdparty as tp
import garment_pattern_estimation_b200 as g
from helpers import rel_err, ref_edgeconv, global_index
from test_gpu_model import _configs, _build
dev = torch.device('cuda:0')
dc, nc, lc = _configs()
torch.manual_seed(21)
oracle = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc)).to(dev)
mine = _build(21, dev); mine.load_state_dict(oracle.state_dict())
B, N = 4, 512
x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(2)).to(dev)
gt = om.synthetic_ground_truth(B, seed=5, device=dev)
torch.manual_seed(7)
h0, c0 = om.init_state(3, B * 23, 250).to(dev), om.init_state(3, B * 23, 250).to(dev)
oracle.train(); mine.train()
capo, capm = {}, {}
def mk(cap):
    def fhook(mod, inp, out):
        cap['x1'] = inp[0].detach().clone()
        out.register_hook(lambda gr: cap.__setitem__('gout', gr.detach().clone()))
    return fhook
oracle.feature_extractor.conv_layers[1].register_forward_hook(mk(capo))
mine.feature_extractor.conv_layers[1].register_forward_hook(mk(capm))
o1 = oracle(x, lstm_state=(h0, c0)); l1, _ = om.main_losses(o1, gt); l1.backward()
o2 = mine(x, lstm_state=(h0, c0)); l2, _, _ = mine.loss(o2, gt); l2.backward()
print('x1 rel', rel_err(capm['x1'], capo['x1']))
go, gm = capo['gout'], capm['gout']
print('gout shapes', go.shape, gm.shape)
gm150 = gm[:, :150]
# oracle's gout is wrt conv output [M,150] BEFORE cat; compare
print('gout rel (first 150 cols)', rel_err(gm150, go))
d = (gm150 - go).abs()
print('worst cols', d.max(dim=0).values.topk(5)); print('worst rows', d.max(dim=1).values.topk(5))
print('gout abs max', float(go.abs().max()), 'l2rel', float((gm150-go).norm()/go.norm()))
# torch reference on MY inputs
import copy
refmlp = copy.deepcopy(oracle.feature_extractor.conv_layers[1].nn)
for p in refmlp.parameters(): p.grad = None
xin = capm['x1'].clone().requires_grad_(True)
idx = global_index(mine.feature_extractor.conv_layers[1].last_index, N)
want = ref_edgeconv(xin, idx, refmlp)
want.backward(gm150.contiguous())
gmine = dict(mine.feature_extractor.conv_layers[1].nn.named_parameters()); gref = dict(refmlp.named_parameters())
for n in gref:
    print('self-consistency %-12s rel %.2e' % (n, rel_err(gmine[n].grad, gref[n].grad)))

print('---- standalone re-runs of MY conv1 on MY captured inputs')
mconv = mine.feature_extractor.conv_layers[1]
def run(tail):
    for p in mconv.nn.parameters(): p.grad = None
    xin2 = capm['x1'].clone().requires_grad_(True)
    pos = x.reshape(-1, 3)
    out = mconv(xin2, cloud_shape=(B, N), tail_src=pos if tail else None)
    out.backward(gm.contiguous() if tail else gm150.contiguous())
    return {n: p.grad.clone() for n, p in mconv.nn.named_parameters()}
for tail in (False, True, True):
    gs = run(tail)
    print('tail', tail, ' '.join('%s:%.1e' % (n, rel_err(gs[n], gref[n].grad)) for n in gref if n.endswith('weight')))
# determinism of the torch reference itself under a 1e-6 perturbation of the input
gref_saved = {n: p.grad.clone() for n, p in refmlp.named_parameters()}
for p in refmlp.parameters(): p.grad = None
xin3 = (capo['x1']).clone().requires_grad_(True)
want3 = ref_edgeconv(xin3, idx, refmlp); want3.backward(gm150.contiguous())
print('torch-ref(oracle x1) vs torch-ref(my x1):', ' '.join('%s:%.1e' % (n, rel_err(p.grad, gref_saved[n])) for n, p in refmlp.named_parameters() if n.endswith('weight')))
