"""GPU parity AT THE SHAPES THE NUMBERS ARE QUOTED ON (BASELINE.json configs C2 / C4 / C5; VERDICT r1 "pin parity where the
numbers are quoted"), plus the pieces of the boundary that round 1 only exercised on the CPU:

  * C2  -- the full attention model at B=32 x N=2048 (327 680 edge rows per GEMM: several tiles per persistent CTA, the
           knn_tc split plan of the benchmark) side by side with the oracle model on the same device: forward, loss, gradients;
  * C4  -- EdgeConvFeatures forward + backward at N=10 000, k=16 (the 2.56 M-edge stress shape, two clouds of it);
  * C5  -- the shipped checkpoint in eval mode at N=8192 (largest point of the inference sweep);
  * gradients at 1e-5 when the ReLU masks of the two implementations are identical (justifies the looser relative-L2 bounds
    used elsewhere: those come from mask flips at kinks, not from the arithmetic);
  * the pattern loss (GT order / origin matching, stitch terms) with every tensor on the device;
  * INTEGRATION.md swap A + C for real: the UNMODIFIED reference nn/nets.py model class driven by the UNMODIFIED
    nn/trainer.py::Trainer._fit_loop on top of the B200 blocks, on a B200.

Tolerances: kNN bit-exact; activations max|a-b| / max|b| <= 1e-3 (BASELINE.json north_star); gradients in relative L2 (helpers).
"""
import os
import subprocess
import sys

import pytest
import torch

from helpers import REL_TOL, assert_close, assert_grad_close, global_index, ref_edgeconv, rel_err, torch_mlp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _configs():
    from oracle import model as om
    dc = dict(om.ATT_DATA_CONFIG)
    dc['standardize'] = {'gt_shift': {'outlines': [0, 0, 0.14890235662460327, 0.05642016604542732]},
                         'gt_scale': {'outlines': [25.267892837524418, 31.298505783081055, 0.2677369713783264,
                                                   0.2352069765329361]}}
    nc = dict(om.ATT_NN_CONFIG)
    lc = {'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [],
          'loop_loss_weight': 1., 'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}
    return dc, nc, lc


# ------------------------------------------------------------------------------------------------------------
# C2: B=32, N=2048 training step, oracle model on the same GPU
# ------------------------------------------------------------------------------------------------------------
def test_c2_full_shape_train_step_side_by_side_with_oracle(cuda_device):
    import garment_pattern_estimation_b200 as g
    from oracle import model as om
    dev = cuda_device
    dc, nc, lc = _configs()
    torch.manual_seed(916143406)                     # models/att/att.yaml:147
    oracle = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc)).to(dev).train()
    mine = g.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc)).to(dev).train()
    mine.load_state_dict(oracle.state_dict())
    B, N = 32, 2048
    x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(1234)).to(dev)
    gt = om.synthetic_ground_truth(B, seed=1235, device=dev)
    torch.manual_seed(7)
    h0, c0 = om.init_state(3, B * 23, 250).to(dev), om.init_state(3, B * 23, 250).to(dev)
    o2 = mine(x, lstm_state=(h0, c0))
    l2, _, _ = mine.loss(o2, gt)
    l2.backward()
    torch.cuda.synchronize()
    o1 = oracle(x, lstm_state=(h0, c0))
    l1, _ = om.main_losses(o1, gt)
    l1.backward()
    # layer-1 graph (raw positions) must be bit-exact at the benchmark shape; layer 2 is built on features that agree to 1e-6
    from oracle import knn as oknn
    idx1 = mine.feature_extractor.conv_layers[0].last_index.view(B, N, -1).cpu()
    assert torch.equal(idx1, oknn.knn_indices(x.cpu(), 5, nthreads=os.cpu_count() or 1))
    for key in o1:
        assert_close(o2[key], o1[key], what='C2 forward ' + key)
    assert abs(float(l1) - float(l2)) <= REL_TOL * abs(float(l1))
    g1, g2 = dict(oracle.named_parameters()), dict(mine.named_parameters())
    worst = 0.0
    for name in g1:
        if g1[name].grad is None:
            assert g2[name].grad is None, name
            continue
        a, b = g2[name].grad.double(), g1[name].grad.double()
        err = float((a - b).norm() / b.norm().clamp_min(1e-30))
        worst = max(worst, err)
        # 327 680 edge rows: ~4x more pre-activations sit within 1e-6 of a ReLU kink than in the 81 920-row case (helpers.py)
        assert err <= 2e-2, 'C2 grad {} relative L2 error {:.2e}'.format(name, err)
    print('C2 worst parameter-gradient relative L2 error: {:.2e}'.format(worst))


# ------------------------------------------------------------------------------------------------------------
# C4: N=10 000, k=16 EdgeConv encoder forward + backward
# ------------------------------------------------------------------------------------------------------------
def test_c4_edgeconv_encoder_n10000_k16(cuda_device):
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    B, N, k = 2, 10000, 16
    cfg = {'conv_depth': 2, 'k_neighbors': k, 'EConv_hidden': 200, 'EConv_hidden_depth': 2, 'EConv_feature': 150,
           'EConv_aggr': 'max', 'global_pool': 'mean', 'skip_connections': True, 'graph_pooling': False}
    torch.manual_seed(4)
    enc = nb.EdgeConvFeatures(250, cfg).to(dev).train()
    refs = [torch_mlp([6, 200, 200, 150]).to(dev).train(), torch_mlp([300, 200, 200, 150]).to(dev).train()]
    for conv, ref in zip(enc.conv_layers, refs):
        ref.load_state_dict(conv.nn.state_dict())
    pos = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(40)).to(dev)
    p1 = pos.clone().requires_grad_(True)
    _, feats, batch = enc(p1, global_pool=False)
    gout = torch.randn_like(feats)
    feats.backward(gout)
    torch.cuda.synchronize()
    # bit-exact xyz graph against the C oracle on one of the clouds (the full C4 kNN shape is covered in test_gpu_ops.py)
    from oracle import knn as oknn
    idx1 = enc.conv_layers[0].last_index.view(B, N, k)
    assert torch.equal(idx1[1].cpu(), oknn.knn_indices(pos[1:2].cpu(), k, nthreads=os.cpu_count() or 1)[0])
    # plain-torch reference layer by layer on the SAME neighbour tables
    p2 = pos.clone().requires_grad_(True)
    flat = p2.reshape(B * N, 3)
    f1 = ref_edgeconv(flat, global_index(enc.conv_layers[0].last_index, N), refs[0])
    f2 = ref_edgeconv(f1, global_index(enc.conv_layers[1].last_index, N), refs[1])
    want = torch.cat([f2, flat], dim=-1)
    want.backward(gout)
    assert feats.shape == (B * N, 153) and batch.shape == (B * N,)
    assert_close(feats, want, what='C4 encoder forward')
    # 20 000 x 3 small entries after two layers of k=16 max-aggregation: one mask / argmax flip moves a single entry by several
    # per cent of the largest one (measured 4.0e-2 worst entry at 4.4e-3 relative L2)
    assert_grad_close(p1.grad, p2.grad, what='C4 grad wrt positions', l2_tol=1e-2, max_tol=1e-1)
    for conv, ref in zip(enc.conv_layers, refs):
        for (n1, a), (n2, b) in zip(conv.nn.named_parameters(), ref.named_parameters()):
            assert_grad_close(a.grad, b.grad, what='C4 grad ' + n1, l2_tol=1e-2)
        for (n1, a), (n2, b) in zip(conv.nn.named_buffers(), ref.named_buffers()):
            if a.dtype.is_floating_point:
                assert_close(a, b, what='C4 BN buffer ' + n1)


# ------------------------------------------------------------------------------------------------------------
# C5: shipped checkpoint, eval mode, N=8192
# ------------------------------------------------------------------------------------------------------------
def test_c5_shipped_checkpoint_eval_n8192(cuda_device, golden_dir):
    ck = os.path.join(golden_dir, '_ckpt', 'att_state.pt')
    if not os.path.exists(ck):
        pytest.skip('tests/golden/_ckpt/att_state.pt absent (generated by __graft_entry__.build() from the reference tree)')
    import garment_pattern_estimation_b200 as g
    from oracle import model as om
    dev = cuda_device
    dc, nc, lc = _configs()
    sd = torch.load(ck)
    oracle = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc))
    oracle.load_state_dict(sd, strict=True)
    oracle.to(dev).eval()
    mine = g.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc))
    mine.load_state_dict(sd, strict=True)
    mine.to(dev).eval()
    mine.save_att_weights = oracle.save_att_weights = True
    B, N = 2, 8192
    x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(55)).to(dev)
    torch.manual_seed(7)
    h0, c0 = om.init_state(3, B * 23, 250).to(dev), om.init_state(3, B * 23, 250).to(dev)
    with torch.no_grad():
        o2 = mine(x, lstm_state=(h0, c0))
        o1 = oracle(x, lstm_state=(h0, c0))
    from oracle import knn as oknn
    idx1 = mine.feature_extractor.conv_layers[0].last_index.view(B, N, -1).cpu()
    assert torch.equal(idx1, oknn.knn_indices(x.cpu(), 5, nthreads=os.cpu_count() or 1))
    for key in o1:
        if key == 'att_weights':      # sparsemax support can flip where a score sits within 1e-6 of the threshold
            assert float((o2[key] - o1[key]).abs().max()) <= 2e-3, key
            continue
        assert_close(o2[key], o1[key], what='C5 N=8192 ' + key)


# ------------------------------------------------------------------------------------------------------------
# gradients at 1e-5 when both implementations see the SAME ReLU masks
# ------------------------------------------------------------------------------------------------------------
def _kink_free_mlp(widths, gen):
    """Weights with one dominant input per unit (|w| = 4 on a signed pseudo-permutation + N(0, 0.02) elsewhere) and small biases.
    Fed with inputs whose coordinates are +-2.5 + N(0, 0.05), every pre-activation of every layer sits near +-10 (the BN output of
    a {0, ~10}-valued ReLU is again two well separated clusters), i.e. no pre-activation comes within ~1 of the ReLU kink: the mask
    is the same for any two fp32-class implementations, while it still differs from row to row."""
    mlp = torch_mlp(widths)
    with torch.no_grad():
        for blk in mlp:
            lin, bn = blk[0], blk[2]
            w = 0.02 * torch.randn(lin.weight.shape, generator=gen)
            cols = torch.randint(0, lin.weight.shape[1], (lin.weight.shape[0],), generator=gen)
            sign = torch.where(torch.rand(lin.weight.shape[0], generator=gen) < 0.5, -1., 1.)
            w[torch.arange(lin.weight.shape[0]), cols] = 4. * sign
            lin.weight.copy_(w)
            lin.bias.copy_(0.1 * torch.randn(lin.bias.shape, generator=gen))
            bn.weight.copy_(2.5 * torch.where(torch.rand(bn.weight.shape, generator=gen) < 0.2, -1., 1.))   # some negative gammas
            bn.bias.copy_(0.05 * torch.randn(bn.bias.shape, generator=gen))
    return mlp


@pytest.mark.parametrize('mode', ['plain', 'edge'])
def test_gradients_agree_to_1e5_when_relu_masks_are_identical(cuda_device, mode):
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    gen = torch.Generator().manual_seed(77)
    C, widths = 24, [64, 64, 40]
    if mode == 'plain':
        rows = 20000
        ref = _kink_free_mlp([C] + widths, gen).to(dev).double().train()
        mine = nb.MLP([C] + widths).to(dev).train()
        mine.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        x = (2.5 * torch.where(torch.rand(rows, C, generator=gen) < 0.5, -1., 1.) + 0.05 * torch.randn(rows, C, generator=gen)).to(dev)
        x1, x2 = x.clone().requires_grad_(True), x.double().clone().requires_grad_(True)
        out, want = mine(x1), ref(x2)
    else:
        B, N, k = 4, 2048, 5
        ref = _kink_free_mlp([2 * C] + widths, gen)
        with torch.no_grad():      # dominant inputs on the x_i half only: x_j - x_i is {0, +-5} + noise and stays a small term
            lin = ref[0][0]
            w = lin.weight.clone()
            big = w.abs() > 1
            w[:, C:] = torch.where(big[:, C:], torch.zeros_like(w[:, C:]), w[:, C:])
            need = ~(w.abs() > 1).any(dim=1)
            w[need, torch.randint(0, C, (int(need.sum()),), generator=gen)] = 4.
            lin.weight.copy_(w)
        ref = ref.to(dev).double().train()
        conv = nb.DynamicEdgeConv(nb.MLP([2 * C] + widths), k=k).to(dev).train()
        conv.nn.load_state_dict({k_: v.float() for k_, v in ref.state_dict().items()})
        mine = conv.nn
        x = (2.5 * torch.where(torch.rand(B * N, C, generator=gen) < 0.5, -1., 1.) + 0.05 * torch.randn(B * N, C, generator=gen)).to(dev)
        x1, x2 = x.clone().requires_grad_(True), x.double().clone().requires_grad_(True)
        out = conv(x1, cloud_shape=(B, N))
        gi = global_index(conv.last_index, N)
        want = ref_edgeconv(x2, gi, ref)
    # the construction must hold: no reference pre-activation near the kink in any layer
    with torch.no_grad():
        if mode == 'plain':
            h = x2.detach()
        else:
            xi = x2.detach().unsqueeze(1).expand(B * N, k, C)
            h = torch.cat([xi, x2.detach()[gi] - xi], dim=-1).reshape(B * N * k, 2 * C)
        for blk in ref:
            z = blk[0](h)
            assert float(z.abs().min()) > 0.5, 'construction broke: a pre-activation sits near the ReLU kink'
            h = blk[2](torch.relu(z))
    assert rel_err(out, want) <= 2e-5, 'forward {:.2e}'.format(rel_err(out, want))
    gout = torch.randn(out.shape, generator=gen).to(dev)
    if mode == 'edge':
        # the max over the k messages is a second source of discontinuity (which edge receives the gradient): take the
        # upstream gradient only where the winner leads the runner-up by a clear margin, so both implementations route it
        # to the same edge
        top2 = h.view(B * N, k, -1).topk(2, dim=1).values
        gout = gout * ((top2[:, 0] - top2[:, 1]) > 1e-3 * float(h.abs().max())).to(gout.dtype)
    out.backward(gout)
    want.backward(gout.double())

    report = []

    def check(a, b, what, l2_tol):
        a, b = a.detach().double(), b.detach().double()
        l2 = float((a - b).norm() / b.norm().clamp_min(1e-30))
        mx = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        report.append('{}: relative L2 {:.2e} (tol {:.0e}), worst entry {:.2e}'.format(what, l2, l2_tol, mx))
        return l2 <= l2_tol and mx <= 10 * l2_tol

    # 2-D gradients (inputs, Linear weights): 1e-5.  1-D gradients (Linear bias, BN gamma / beta) are sums over all rows whose
    # terms cancel almost completely behind a BatchNorm (sum_r dz ~ 0): the fp32 rounding of the terms is amplified by the
    # cancellation factor in ANY fp32 implementation (measured 3e-5 .. 9e-5 against the float64 reference) -> 3e-4.
    ok = check(x1.grad, x2.grad, mode + ' grad wrt input', 1e-5)
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        ok = check(p1.grad, p2.grad, mode + ' grad ' + n1, 1e-5 if p1.dim() > 1 else 3e-4) and ok
    print('\n'.join(report))
    assert ok, '\n'.join(report)


# ------------------------------------------------------------------------------------------------------------
# pattern loss with every tensor on the device: GT order / origin matching, stitch terms (goldens of the unmodified reference)
# ------------------------------------------------------------------------------------------------------------
def _to_dev(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


def _check_parts(want_parts, want_total, total, parts, tol=5e-5):
    import math
    assert set(parts) == set(want_parts)
    assert abs(float(total) - float(want_total)) <= tol * abs(float(want_total))
    for k, want in want_parts.items():
        got = parts[k]
        if want is None:
            assert got is None, k
            continue
        w, gv = float(want), float(got)
        assert (math.isnan(w) and math.isnan(gv)) or abs(gv - w) <= tol * max(abs(w), 1e-3), (k, gv, w)


def test_gt_matching_and_stitch_terms_on_device_match_reference_golden(cuda_device, golden_dir):
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    from oracle import model as om
    dev = cuda_device
    dc = dict(om.ATT_DATA_CONFIG)
    gold = torch.load(os.path.join(golden_dir, 'n1_matching.pt'))
    dc['standardize'] = gold['standardize']
    for name, case in gold['cases'].items():
        loss_obj = ComposedPatternLoss(dc, dict(case['loss_config']))
        total, parts, flag = loss_obj(_to_dev(case['preds'], dev), _to_dev(case['gt'], dev), epoch=3)
        assert total.is_cuda and bool(flag) == case['flag'], name
        _check_parts(case['parts'], case['loss'], total, parts)
    gold = torch.load(os.path.join(golden_dir, 'n1_stitch_terms.pt'))
    dc['standardize'] = gold['standardize']
    for name, case in gold['cases'].items():
        loss_obj = ComposedPatternLoss(dc, dict(case['loss_config']))
        for epoch, want in case['runs'].items():
            total, parts, flag = loss_obj(_to_dev(case['preds'], dev), _to_dev(case['gt'], dev), epoch=epoch)
            assert total.is_cuda and bool(flag) == want['flag'], (name, epoch)
            _check_parts(want['parts'], want['loss'], total, parts)


# ------------------------------------------------------------------------------------------------------------
# INTEGRATION.md swap A + C on a B200: unmodified reference nets.py + trainer.py on top of the B200 blocks
# ------------------------------------------------------------------------------------------------------------
_SWAP_A = r'''
import os, sys, types, torch
sys.path.insert(0, %(root)r)
from oracle import ref_stubs
from oracle import model as om
ref_stubs.install()                                   # stand-ins for entmax / data / torch_geometric / wandb
DRY = os.environ.get('NT_SWAP_A_DRY') == '1'          # harness check on the CPU: oracle blocks instead of the B200 blocks
import garment_pattern_estimation_b200.net_blocks as b200_blocks
import garment_pattern_estimation_b200.nets as b200_nets
from garment_pattern_estimation_b200.parallel import FlatDataParallel
if not DRY:                                           # (DRY: the reference's own net_blocks on the restated operators)
    sys.modules['net_blocks'] = b200_blocks               # INTEGRATION.md swap A
    sys.modules['sparsemax'] = types.SimpleNamespace(Sparsemax=b200_nets.Sparsemax)
import nets                                           # the UNMODIFIED reference module (nn/nets.py)
import trainer as ref_trainer                         # the UNMODIFIED reference module (nn/trainer.py)
import wandb as wb
assert (DRY or nets.blocks is b200_blocks) and nets.__file__.startswith(ref_stubs.REFERENCE_ROOT)
torch.backends.cudnn.allow_tf32 = False
DEVNAME = 'cpu' if DRY else 'cuda:0'
dev = torch.device(DEVNAME)
dc, nc, lc = ref_stubs.att_configs()
torch.manual_seed(916143406)
model = nets.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc))
assert type(model).__module__ == 'nets'
assert DRY or type(model.feature_extractor).__module__ == 'garment_pattern_estimation_b200.net_blocks'
oracle = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc))
oracle.load_state_dict(model.state_dict(), strict=True)
oracle.to(dev)
model.to(dev)

# ---- forward parity of the reference's own forward() (its 23-iteration pooling loop included) on the B200 blocks
B, N = (2, 64) if DRY else (4, 512)
x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(5)).to(dev)
torch.manual_seed(7)
state = (om.init_state(3, B * 23, 250).to(dev), om.init_state(3, B * 23, 250).to(dev))
if not DRY:
    model.panel_decoder.state_provider = lambda *a: state      # the reference forward has no lstm_state argument
for mode in (() if DRY else ('train', 'eval')):
    getattr(model, mode)(); getattr(oracle, mode)()
    out = model(x)
    want = oracle(x, lstm_state=state)
    for key in want:
        err = float((out[key] - want[key]).abs().max() / want[key].abs().max())
        assert err <= 1e-3, (mode, key, err)
print('reference forward on B200 blocks matches the oracle')
if not DRY:
    model.panel_decoder.state_provider = None

# ---- the reference Trainer's own fit loop (nn/trainer.py:83-136), unmodified, on the wrapped model
class Experiment:                                      # duck-typed ExperimentWrappper (SURVEY A.6): no W&B, no files
    checkpoints, stats, configs = [], {}, {}
    def init_run(self, cfg): wb.config.trainer.update(cfg['trainer'])
    def last_best_validation_loss(self): return None
    def add_config(self, tag, cfg): self.configs[tag] = cfg
    def add_statistic(self, tag, info, log=''): self.stats[tag] = info
    def save_checkpoint(self, state, aliases=[], wait_for_upload=False): self.checkpoints.append((state, list(aliases)))
    def cloud_path(self): return 'local'

def batches(n, seed):
    out = []
    for i in range(n):
        gt = om.synthetic_ground_truth(B, seed=seed + i)
        out.append({'features': torch.randn(B, N, 3, generator=torch.Generator().manual_seed(seed + 50 + i)),
                    'ground_truth': gt, 'name': ['s%%d' %% j for j in range(B)], 'data_folder': ['synthetic'] * B})
    return out

setup = {'batch_size': B, 'devices': [DEVNAME], 'epochs': 2, 'random_seed': 916143406, 'learning_rate': 0.002,
         'optimizer': 'Adam', 'weight_decay': 0, 'lr_scheduling': {'mode': '1cyclic'},
         'early_stopping': {'window': 0.0001, 'patience': 50}}
exp = Experiment()
tr = ref_trainer.Trainer(setup, exp, dataset=None, with_visualization=False)
tr.init_randomizer()
train_loader, valid_loader = batches(3, 100), batches(1, 900)
tr.datawraper = types.SimpleNamespace(loaders=types.SimpleNamespace(train=train_loader, validation=valid_loader),
                                      save_to_wandb=lambda experiment: None)
wrapped = FlatDataParallel(model, device_ids=[DEVNAME])          # INTEGRATION.md swap C (was nn.DataParallel, nn/train.py:124)
before = {k: v.detach().clone() for k, v in wrapped.state_dict().items()}
tr.fit(wrapped)                                                    # -> _add_optimizer, _add_scheduler(OneCycleLR), _fit_loop
steps = [d for _, d in wb.logged if 'batch' in d]
assert len(steps) == 6 and all(torch.isfinite(d['loss']) for d in steps), len(steps)
assert {'pattern_loss', 'loop_loss', 'rotation_loss', 'translation_loss', 'learning_rate'} <= set(steps[0])
assert steps[1]['learning_rate'] != steps[0]['learning_rate']     # OneCycleLR stepped (nn/trainer.py:100-101)
epochs = [d for _, d in wb.logged if 'valid_loss' in d]
assert len(epochs) == 2 and all(torch.isfinite(d['valid_loss']) for d in epochs)
assert len(exp.checkpoints) == 2 and 'best' in exp.checkpoints[0][1]
ck = exp.checkpoints[-1][0]
assert set(ck) == {'epoch', 'model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict'}
assert all(k.startswith('module.') for k in ck['model_state_dict'])
after = wrapped.state_dict()
moved = sum(int(not torch.equal(before[k], after[k])) for k in before if before[k].dtype.is_floating_point)
n_float = sum(int(v.dtype.is_floating_point) for v in before.values())
assert moved >= n_float - 2, (moved, n_float)      # everything but the unused feature_extractor.lin.* (SURVEY F6)
assert int(after['module.feature_extractor.conv_layers.0.nn.0.2.num_batches_tracked']) >= 6
first, last = float(steps[0]['loss']), float(steps[-1]['loss'])
print('trainer ran: loss %%.4f -> %%.4f over %%d steps, %%d checkpoints' %% (first, last, len(steps), len(exp.checkpoints)))
'''


def test_unmodified_reference_nets_and_trainer_run_on_the_b200_blocks(cuda_device):
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('neither /root/reference nor the staged copy tests/golden/_ref_nn/ (written by __graft_entry__.build()) exists')
    out = subprocess.run([sys.executable, '-c', _SWAP_A % {'root': ROOT}], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert 'reference forward on B200 blocks matches the oracle' in out.stdout and 'trainer ran' in out.stdout
