import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
torch.backends.cudnn.allow_tf32 = False
from oracle import model as om
from oracle import thirdparty as tp
import garment_pattern_estimation_b200 as g
from helpers import rel_err, ref_edgeconv
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
cap = {}
conv1 = oracle.feature_extractor.conv_layers[1]
def fhook(mod, inp, out):
    cap['x1'] = inp[0].detach().clone()
    out.register_hook(lambda gr: cap.__setitem__('gout', gr.detach().clone()))
conv1.register_forward_hook(fhook)
zs = {}
def zhook(name):
    def h(mod, inp, out):
        out.register_hook(lambda gr: zs.__setitem__(name, gr.detach().clone()))
        zs[name + '_val'] = out.detach()
    return h
for li in range(3):
    conv1.nn[li][0].register_forward_hook(zhook('z%d' % (li + 1)))
graphs = []
orig = tp.knn_graph
def spy(x_, b_, k_):
    r = orig(x_, b_, k_); graphs.append(r); return r
tp.knn_graph = spy
o1 = oracle(x, lstm_state=(h0, c0))
l1, _ = om.main_losses(o1, gt); l1.backward()
x1, gout = cap['x1'], cap['gout']
print('x1', x1.shape, 'gout', gout.shape)
# standalone mine conv1 with same weights
mconv = mine.feature_extractor.conv_layers[1]
xin = x1.clone().requires_grad_(True)
out = mconv(xin, cloud_shape=(B, N))
out.backward(gout.contiguous())
gm = dict(mconv.nn.named_parameters()); go = dict(conv1.nn.named_parameters())
for n in go:
    d = (gm[n].grad - go[n].grad).abs()
    print('%-12s rel %.2e' % (n, float(d.max() / go[n].grad.abs().max())))
# per-channel db error for last linear
d = (gm['2.0.bias'].grad - go['2.0.bias'].grad).abs()
print('db3 worst channels', d.topk(5))
bn3 = conv1.nn[2][2]
c = int(d.argmax())
z3 = zs['z3_val']; a3 = z3.clamp_min(0)
print('channel', c, 'gamma', float(bn3.weight[c]), 'frac a3>0', float((a3[:, c] > 0).float().mean()), 'mean', float(a3[:, c].mean()), 'std', float(a3[:, c].std()))
print('oracle dz3 colsum', float(zs['z3'][:, c].sum()), 'mine', float(gm['2.0.bias'].grad[c]), 'ref', float(go['2.0.bias'].grad[c]))
# count of rows where |z3| tiny
print('rows with |z3|<1e-6 in channel', int((z3[:, c].abs() < 1e-6).sum()), 'exact zeros', int((z3[:, c] == 0).sum()))
# how many nodes have ties in max over k for this channel (post-BN values equal)
y3 = bn3(a3) if False else None
a3n = a3.view(-1, 5, a3.shape[1])
mx = a3n.max(dim=1).values
ties = (a3n == mx.unsqueeze(1)).sum(dim=1)
print('nodes with tie at max in channel', int((ties[:, c] > 1).sum()), 'of', ties.shape[0], ' (ties where max>0:', int(((ties[:, c] > 1) & (mx[:, c] > 0)).sum()), ')')
mn = a3n.min(dim=1).values
tmin = (a3n == mn.unsqueeze(1)).sum(dim=1)
print('gamma<0 channels', int((bn3.weight < 0).sum()), 'nodes with tie at min in channel', int((tmin[:, c] > 1).sum()))
# grad of oracle wrt z3 at tie nodes: where does torch route gradient?
gz = zs['z3'].view(-1, 5, a3.shape[1])[:, :, c]
print('oracle: nonzero dz3 per node histogram', torch.bincount((gz != 0).sum(dim=1), minlength=6))
