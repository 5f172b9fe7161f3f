"""CPU tests of the host side: the reference-facing module surface (names, config merging, state_dict contract, error
behaviour), the vectorised loss against the reference's loop form, and the data-parallel wrapper over gloo (world size 2)."""
import os
import socket
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _att():
    from oracle import model as om
    return dict(om.ATT_DATA_CONFIG), dict(om.ATT_NN_CONFIG), {
        'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [],
        'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}


def test_state_dict_contract_matches_survey_a2():
    import garment_pattern_estimation_b200 as g
    dc, nc, lc = _att()
    model = g.GarmentSegmentPattern3D(dc, nc, lc)
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert shapes['feature_extractor.conv_layers.0.nn.0.0.weight'] == (200, 6)
    assert shapes['feature_extractor.conv_layers.1.nn.0.0.weight'] == (200, 300)
    assert shapes['feature_extractor.conv_layers.1.nn.2.0.weight'] == (150, 200)
    assert shapes['feature_extractor.conv_layers.0.nn.1.2.running_var'] == (200,)
    assert sd['feature_extractor.conv_layers.0.nn.1.2.num_batches_tracked'].dtype == torch.int64
    assert shapes['feature_extractor.lin.weight'] == (250, 153)
    assert shapes['panel_decoder.lstm.weight_ih_l2'] == (1000, 250) and shapes['panel_decoder.lin.weight'] == (8, 250)
    assert shapes['placement_decoder.weight'] == (7, 250)
    assert shapes['point_segment_mlp.0.2.0.weight'] == (23, 153) and shapes['panel_dec_lin.weight'] == (250, 153)
    assert not any(k.startswith('pattern_decoder') for k in sd)
    assert sum(p.numel() for p in model.parameters()) == 1842589
    # wrapped checkpoints carry the 'module.' prefix (nn/trainer.py:275-291)
    from garment_pattern_estimation_b200.parallel import FlatDataParallel
    wrapped = FlatDataParallel(model)
    assert all(k.startswith('module.') for k in wrapped.state_dict())
    assert wrapped.module is model and len(wrapped.device_ids) == 1


def test_same_seed_gives_same_init_as_oracle_and_baseline_model_builds():
    import garment_pattern_estimation_b200 as g
    from oracle import model as om
    dc, nc, lc = _att()
    torch.manual_seed(123)
    a = g.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc)).state_dict()
    torch.manual_seed(123)
    b = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc)).state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    nc2 = dict(nc)
    nc2.update(model='GarmentFullPattern3D', skip_connections=False)
    base = g.GarmentFullPattern3D(dict(dc), nc2, dict(lc))
    keys = base.state_dict().keys()
    assert 'pattern_decoder.lstm.weight_ih_l1' in keys and 'feature_extractor.lin.weight' in keys
    assert base.state_dict()['feature_extractor.lin.weight'].shape == (250, 150)


def test_config_merging_follows_the_reference():
    import garment_pattern_estimation_b200 as g
    from garment_pattern_estimation_b200 import net_blocks as nb
    dc, nc, lc = _att()
    model = g.GarmentSegmentPattern3D(dc, nc, lc)
    # block defaults .update()-ed with the whole NN config, then merged back (nn/net_blocks.py:98-110, nn/nets.py:102-103)
    assert model.feature_extractor.config['panel_encoding_size'] == 250
    assert model.config['EConv_feature'] == 150 and model.config['k_neighbors'] == 5
    assert model.config['loss']['loss_components'] == ['shape', 'loop', 'rotation', 'translation']
    enc = nb.EdgeConvFeatures(10)
    assert enc.config['EConv_feature'] == 112 and enc.config['skip_connections'] is False
    assert enc.lin.in_features == 112
    assert callable(enc.global_pool)


def test_error_behaviour_matches_reference_conventions():
    from garment_pattern_estimation_b200 import net_blocks as nb
    with pytest.raises(ValueError):                                   # nn/net_blocks.py:152
        nb.EdgeConvFeatures(10, {'global_pool': 'median'})
    with pytest.raises(NotImplementedError):                          # nn/net_blocks.py:313
        nb.initial_state(3, 4, 5, 'cpu', init_type='xavier')
    with pytest.raises(NotImplementedError):                          # nn/net_blocks.py:332
        nb.LSTMDecoderModule(8, 8, 4, 1, custom_init='xavier')
    with pytest.raises(NotImplementedError):
        nb.EdgeConvFeatures(10, {'graph_pooling': True})
    enc = nb.EdgeConvFeatures(10)
    with pytest.raises(RuntimeError):                                 # no CPU fallback
        enc(torch.randn(2, 16, 3))


def test_global_pools_have_no_cpu_fallback():
    """global_{mean,max,add}_pool run in libnt_b200 (nt_global_pool_*; numerics in tests/test_gpu_ops.py): CPU tensors are refused
    with the documented RuntimeError, a bad pool name with the reference's ValueError (nn/net_blocks.py:152)."""
    import pytest
    from garment_pattern_estimation_b200 import net_blocks as nb
    from garment_pattern_estimation_b200 import ops
    x = torch.randn(3 * 7, 5)
    batch = torch.arange(3).repeat_interleave(7)
    for pool in (nb.global_mean_pool, nb.global_max_pool, nb.global_add_pool):
        with pytest.raises(RuntimeError):
            pool(x, batch, 3)
    with pytest.raises(ValueError):
        ops.global_pool(x, 3, 7, 'median')


def test_initial_state_distribution_matches_reference_draw():
    from garment_pattern_estimation_b200 import net_blocks as nb
    from oracle import model as om
    torch.manual_seed(0)
    mine = nb.initial_state(3, 736, 250, 'cpu', 'kaiming_normal_')
    ref = om.init_state(3, 736, 250)
    assert mine.shape == ref.shape
    assert abs(float(mine.std()) / float(ref.std()) - 1) < 0.02 and abs(float(mine.mean())) < 1e-4
    assert float(nb.initial_state(2, 3, 4, 'cpu', '').abs().sum()) == 0


def test_vectorised_loss_equals_reference_loop_form():
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss, panel_loop_loss
    from oracle import model as om
    dc, _, lc = _att()
    B = 5
    gt = om.synthetic_ground_truth(B, seed=3)
    gt['num_edges'][0, :3] = torch.tensor([0, 2, 14])                 # < 3 edges contribute 0 but count in the mean
    preds = {'outlines': torch.randn(B, 23, 14, 4, requires_grad=True), 'rotations': torch.randn(B, 23, 4),
             'translations': torch.randn(B, 23, 3)}
    loop_ref = om.panel_loop_loss(preds['outlines'], gt['num_edges'].view(-1), fast=False)
    loop_new = panel_loop_loss(preds['outlines'], gt['num_edges'].view(-1))
    assert torch.allclose(loop_ref, loop_new, rtol=1e-6, atol=1e-8)
    loss_obj = ComposedPatternLoss(dc, lc)
    total, parts, flag = loss_obj(preds, gt, epoch=3)
    want, want_parts = om.main_losses(preds, gt)
    assert flag is False and set(parts) == set(want_parts)
    assert torch.allclose(total, want, rtol=1e-6)
    g1, = torch.autograd.grad(total, preds['outlines'], retain_graph=True)
    g2, = torch.autograd.grad(want, preds['outlines'])
    assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-8)
    assert panel_loop_loss(torch.randn(4, 14, 4)).dim() == 0          # no num_edges => no padding assumed
    with pytest.raises(NotImplementedError):                              # accepted at construction, refused when evaluated
        ComposedPatternLoss(dc, {'loss_components': ['shape', 'segmentation'], 'panel_origin_invariant_loss': False,
                                 'panel_order_inariant_loss': False})(preds, gt)
    loss_obj.train(True)
    assert loss_obj.training is True
    loss_obj.eval()
    assert loss_obj.training is False


# ------------------------------------------------------------------------------------------------------------
# data-parallel wrapper over gloo, world size 2 (the N>1 path of SURVEY.md section 8e on CPU)
# ------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, results):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from garment_pattern_estimation_b200.parallel import FlatDataParallel
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # different init per rank: the wrapper must broadcast rank 0's
        net = nn.Sequential(nn.Linear(6, 8), nn.BatchNorm1d(8), nn.ReLU(), nn.Linear(8, 3))
        unused = nn.Linear(4, 4)                           # never used in forward: gradient stays zero
        net.add_module('unused', unused)
        fwd = lambda m, x: m[3](m[2](m[1](m[0](x))))       # noqa: E731
        dp = FlatDataParallel(net, auto_reduce=True)
        torch.manual_seed(7)
        X, Y = torch.randn(8, 6), torch.randn(8, 3)        # the same global batch on both ranks
        xs, ys = dp.shard(X), dp.shard(Y)
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        for it in range(2):
            loss = ((fwd(dp.module, xs) - ys) ** 2).mean()
            loss.backward()                                # auto all-reduce at the end of backward
            opt.step()
            opt.zero_grad(set_to_none=(it == 0))           # the reference Trainer's default drops the flat views once
        dp.sync_buffers()
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        results[rank] = (flat.clone(), net[1].running_mean.clone(), float(unused.weight.grad.abs().sum())
                         if unused.weight.grad is not None else 0.0)
    finally:
        dist.destroy_process_group()


def test_flat_data_parallel_gloo_world2_matches_single_process():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, results), nprocs=world, join=True)
    p0, rm0, unused0 = results[0]
    p1, rm1, unused1 = results[1]
    assert torch.equal(p0, p1), 'replicas diverged'
    assert torch.equal(rm0, rm1), 'BN buffers must follow rank 0'
    assert unused0 == 0.0 and unused1 == 0.0
    # single-process replay: per-rank BN statistics, averaged gradients (DataParallel semantics)
    torch.manual_seed(100)
    net = nn.Sequential(nn.Linear(6, 8), nn.BatchNorm1d(8), nn.ReLU(), nn.Linear(8, 3))
    net.add_module('unused', nn.Linear(4, 4))
    replicas = [net, None]
    import copy
    replicas[1] = copy.deepcopy(net)
    torch.manual_seed(7)
    X, Y = torch.randn(8, 6), torch.randn(8, 3)
    for it in range(2):
        grads = []
        for r, m in enumerate(replicas):
            for p in m.parameters():
                p.grad = None
            xs, ys = X[r * 4:(r + 1) * 4], Y[r * 4:(r + 1) * 4]
            loss = ((m[3](m[2](m[1](m[0](xs)))) - ys) ** 2).mean()
            loss.backward()
            grads.append([p.grad if p.grad is not None else torch.zeros_like(p) for p in m.parameters()])
        for m in replicas:
            with torch.no_grad():
                for p, g0, g1 in zip(m.parameters(), grads[0], grads[1]):
                    p -= 0.1 * (g0 + g1) / 2
    want = torch.cat([p.detach().reshape(-1) for p in replicas[0].parameters()])
    assert torch.allclose(p0, want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rm0, replicas[0][1].running_mean, rtol=1e-5, atol=1e-7)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the agreed keys (tiny run: it times the CPU oracle)."""
    import json
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'clouds/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['higher_is_better'] is True


def test_blocks_drop_into_unmodified_reference_nets_when_available():
    """INTEGRATION.md swap A: the reference's own nn/nets.py, with this package installed as `net_blocks`, builds the
    attention model, loads the shipped checkpoint strict=True and routes forward() into the B200 blocks (which refuse CPU
    tensors).  Needs /root/reference (build container only); runs in a subprocess to keep sys.modules clean."""
    import subprocess
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('reference tree not present on this machine')
    code = r'''
import sys, types, torch
sys.path.insert(0, %r)
from oracle import ref_stubs
ref_stubs.install()                                   # stubs for entmax / data / torch_geometric (unused below)
import garment_pattern_estimation_b200.net_blocks as b200_blocks
import garment_pattern_estimation_b200.nets as b200_nets
sys.modules['net_blocks'] = b200_blocks
sys.modules['sparsemax'] = types.SimpleNamespace(Sparsemax=b200_nets.Sparsemax)
import nets                                           # the UNMODIFIED reference module
assert nets.blocks is b200_blocks
dc, nc, lc = ref_stubs.att_configs()
model = nets.GarmentSegmentPattern3D(dc, nc, lc)
assert type(model.feature_extractor).__module__ == 'garment_pattern_estimation_b200.net_blocks'
print(model.load_state_dict(ref_stubs.att_checkpoint_state(), strict=True))
model.eval()
try:
    model(torch.randn(2, 32, 3))
except RuntimeError as e:
    assert 'CUDA device' in str(e), e
    print('forward reached the B200 blocks')
''' % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'All keys matched' in out.stdout and 'forward reached the B200 blocks' in out.stdout


# ------------------------------------------------------------------------------------------------------------
# quality metrics of the pattern loss (vectorised, device-side) vs the unmodified reference
# ------------------------------------------------------------------------------------------------------------
def _check_loss_against(parts_ref, total_ref, total, parts, tol=2e-5):
    import math
    assert set(parts) == set(parts_ref)
    assert abs(float(total) - float(total_ref)) <= tol * abs(float(total_ref))
    for k, want in parts_ref.items():
        got = parts[k]
        if want is None:
            assert got is None, k
            continue
        assert got is not None, k
        w, gval = float(want), float(got)
        if math.isnan(w):
            assert math.isnan(gval), k
        else:
            assert abs(gval - w) <= tol * max(abs(w), 1e-3), (k, gval, w)


def test_pattern_loss_with_quality_metrics_matches_reference_golden():
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    gold = torch.load(os.path.join(ROOT, 'tests', 'golden', 'n1_quality.pt'))
    dc, _, _ = _att()
    dc['standardize'] = gold['standardize']
    loss_obj = ComposedPatternLoss(dc, dict(gold['loss_config']))
    assert loss_obj.with_quality_eval is True
    for name, case in gold['cases'].items():
        total, parts, flag = loss_obj(case['preds'], {k: v.clone() for k, v in case['gt'].items()}, epoch=3)
        assert flag is False
        _check_loss_against(case['parts'], case['loss'], total, parts)
    loss_obj.with_quality_eval = False
    _, parts, _ = loss_obj(gold['cases']['mixed']['preds'], gold['cases']['mixed']['gt'])
    assert set(parts) == {'pattern_loss', 'loop_loss', 'rotation_loss', 'translation_loss'}
    ComposedPatternLoss(dc, dict(gold['loss_config'], quality_components=['stitch']))      # builds; raises when evaluated


def test_quality_metrics_match_unmodified_reference_on_random_batches():
    """Runs only where /root/reference exists: random prediction batches (padding rows, open loops, random panel counts)
    through the reference's ComposedPatternLoss and through the vectorised one."""
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('reference tree not present on this machine')
    from oracle import model as om
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    ref_stubs.import_reference()
    import metrics.composed_loss as cl
    dc, _, lc = ref_stubs.att_configs()
    lc = dict(lc, loss_components=['shape', 'loop', 'rotation', 'translation'],
              quality_components=['shape', 'discrete', 'rotation', 'translation'])
    ref_loss = cl.ComposedPatternLoss(dict(dc), dict(lc))
    mine = ComposedPatternLoss(dict(dc), dict(lc))
    st = dc['standardize']
    pad = -torch.tensor(st['gt_shift']['outlines']) / torch.tensor(st['gt_scale']['outlines'])
    for seed in range(4):
        g = torch.Generator().manual_seed(100 + seed)
        B = 3 + seed
        gt = om.synthetic_ground_truth(B, seed=50 + seed)
        live = torch.arange(14)[None, None, :] < gt['num_edges'][..., None]
        pred = torch.where(live[..., None], gt['outlines'], pad.expand_as(gt['outlines']).clone())
        noise = [0.003, 0.02, 0.06, 0.2][seed]
        pred = pred + noise * torch.randn(pred.shape, generator=g)
        drop = torch.rand(B, 23, generator=g) < 0.1                               # some panels vanish, some appear
        pred = torch.where(drop[..., None, None], pad.expand_as(pred), pred)
        preds = {'outlines': pred, 'rotations': torch.randn(B, 23, 4, generator=g),
                 'translations': torch.randn(B, 23, 3, generator=g)}
        t1, p1, _ = ref_loss({k: v.clone() for k, v in preds.items()}, {k: v.clone() for k, v in gt.items()}, epoch=1)
        t2, p2, _ = mine(preds, gt, epoch=1)
        _check_loss_against(p1, t1, t2, p2)


@pytest.mark.parametrize('loss_cfg', [
    dict(panel_origin_invariant_loss=True, panel_order_inariant_loss=False),
    dict(panel_origin_invariant_loss=True, panel_order_inariant_loss=True, order_by='placement'),
    dict(panel_origin_invariant_loss=False, panel_order_inariant_loss=True, order_by='shape_translation'),
    dict(panel_origin_invariant_loss=True, panel_order_inariant_loss=True, order_by='translation', epoch_with_order_matching=5),
])
def test_gt_order_and_origin_matching_match_unmodified_reference(loss_cfg):
    """SURVEY.md section 8f row N1, second half (runs only where /root/reference exists): the vectorised panel-order and
    edge-loop-origin matching against the reference's per-panel Python loops -- loss terms, quality metrics and the
    structure-update flag, at epochs before / at / after the start of order matching."""
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('reference tree not present on this machine')
    from oracle import model as om
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    ref_stubs.import_reference()
    import metrics.composed_loss as cl
    dc, _, lc = ref_stubs.att_configs()
    lc = dict(lc, loss_components=['shape', 'loop', 'rotation', 'translation'],
              quality_components=['shape', 'discrete', 'rotation', 'translation'], **loss_cfg)
    ref_loss = cl.ComposedPatternLoss(dict(dc), dict(lc))
    mine = ComposedPatternLoss(dict(dc), dict(lc))
    st = dc['standardize']
    pad = -torch.tensor(st['gt_shift']['outlines']) / torch.tensor(st['gt_scale']['outlines'])
    for seed, epoch in ((0, 0), (1, 5), (2, 9)):
        g = torch.Generator().manual_seed(200 + seed)
        B = 4
        gt = om.synthetic_ground_truth(B, seed=70 + seed)
        gt['empty_panels_mask'] = gt['num_edges'] == 0
        live = torch.arange(14)[None, None, :] < gt['num_edges'][..., None]
        outl = torch.where(live[..., None], gt['outlines'], pad.expand_as(gt['outlines']).clone())
        # predictions = GT with panels shuffled per pattern and every edge loop started at a random edge, plus noise
        perm = torch.stack([torch.randperm(23, generator=g) for _ in range(B)])
        pred = torch.gather(outl, 1, perm[..., None, None].expand_as(outl)).clone()
        ne = torch.gather(gt['num_edges'], 1, perm)
        for b in range(B):
            for p in range(23):
                n = int(ne[b, p])
                if n >= 3:
                    s0 = int(torch.randint(0, n, (1,), generator=g))
                    pred[b, p, :n] = torch.roll(pred[b, p, :n], -s0, dims=0)
        preds = {'outlines': pred + 0.01 * torch.randn(pred.shape, generator=g),
                 'rotations': torch.gather(gt['rotations'], 1, perm[..., None].expand(B, 23, 4)) + 0.01 * torch.randn(B, 23, 4, generator=g),
                 'translations': torch.gather(gt['translations'], 1, perm[..., None].expand(B, 23, 3)) + 0.01 * torch.randn(B, 23, 3, generator=g)}
        torch.manual_seed(11)
        t1, p1, f1 = ref_loss({k: v.clone() for k, v in preds.items()}, {k: v.clone() for k, v in gt.items()}, epoch=epoch)
        torch.manual_seed(11)
        t2, p2, f2 = mine(preds, {k: v.clone() for k, v in gt.items()}, epoch=epoch)
        assert bool(f1) == bool(f2), (loss_cfg, epoch)
        _check_loss_against(p1, t1, t2, p2)


def test_gt_matching_matches_reference_golden():
    """Same check as above against the committed fixture (runs everywhere, including the GPU box without the reference)."""
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    gold = torch.load(os.path.join(ROOT, 'tests', 'golden', 'n1_matching.pt'))
    dc, _, _ = _att()
    dc['standardize'] = gold['standardize']
    for name, case in gold['cases'].items():
        loss_obj = ComposedPatternLoss(dc, dict(case['loss_config']))
        total, parts, flag = loss_obj(case['preds'], {k: v.clone() for k, v in case['gt'].items()}, epoch=3)
        assert bool(flag) == case['flag'], name
        _check_loss_against(case['parts'], case['loss'], total, parts)


def test_bench_encoder_roofline_accounting():
    """bench.encoder_roofline: compulsory bytes of the fused encoder over the encoder kernels' time."""
    sys.path.insert(0, ROOT)
    import bench
    groups = {'nt_knn[D=3]': 0.1, 'nt_gemm_nt[relu_stats,plain]': 0.5, 'nt_edge_activation': 0.4, 'nt_attn_pool_fwd': 9.0,
              'nt_sparsemax_fwd': 9.0}
    r = bench.encoder_roofline(groups, 2.0, 32, 2048, 6500.0)
    assert abs(r['ms_per_step'] - 1.0) < 1e-12                      # attention-head kernels are not part of the encoder
    assert r['compulsory_bytes_per_step'] == 3.0 * 32 * 2048 * 1812
    assert abs(r['achieved'] - r['compulsory_bytes_per_step'] / 1e-3 / 1e9) < 1e-6
    assert abs(r['frac'] - r['achieved'] / 6500.0) < 1e-12 and abs(r['share_of_step'] - 0.5) < 1e-12


def test_package_config_values_and_bench_inputs_equal_the_oracle_copies():
    """bench.py's measured arm and the tools take the att.yaml values and the synthetic ground truth from the package / from
    bench.py itself (the oracle is checker-only); both copies must stay identical to the oracle's."""
    sys.path.insert(0, ROOT)
    import bench
    from garment_pattern_estimation_b200 import configs
    from oracle import model as om
    assert configs.ATT_NN_CONFIG == om.ATT_NN_CONFIG and configs.ATT_DATA_CONFIG == om.ATT_DATA_CONFIG
    a, b = bench.synthetic_ground_truth(3, seed=5), om.synthetic_ground_truth(3, seed=5)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    gold = torch.load(os.path.join(ROOT, 'tests', 'golden', 'n1_quality.pt'))
    for key in ('outlines', 'rotations', 'translations'):          # att.yaml:61-73
        assert configs.ATT_STANDARDIZE['gt_shift'][key] == gold['standardize']['gt_shift'][key]
        assert configs.ATT_STANDARDIZE['gt_scale'][key] == gold['standardize']['gt_scale'][key]
    src = open(os.path.join(ROOT, 'bench.py')).read()
    measured_arm = src[src.index('def main():'):]
    assert not [ln for ln in measured_arm.splitlines() if 'import' in ln and 'oracle' in ln], 'main() must not import oracle/'


# ------------------------------------------------------------------------------------------------------------
# stitch-related loss terms of the shipped baseline config (the model must build from the shipped yaml)
# ------------------------------------------------------------------------------------------------------------
def test_stitch_loss_terms_match_reference_golden():
    """nn/metrics/losses.py:PatternStitchLoss (both negative terms), the free-edge BCE, stitch_supervised and the re-numbering
    of the stitch ground truth under panel-order / edge-origin matching, against what the unmodified reference returned
    (tests/golden/make_golden_n1_stitch.py) -- before, at and after config['epoch_with_stitches']."""
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    gold = torch.load(os.path.join(ROOT, 'tests', 'golden', 'n1_stitch_terms.pt'))
    dc, _, _ = _att()
    dc['standardize'] = gold['standardize']
    for name, case in gold['cases'].items():
        loss_obj = ComposedPatternLoss(dc, dict(case['loss_config']))
        for epoch, want in case['runs'].items():
            total, parts, flag = loss_obj(case['preds'], {k: v.clone() for k, v in case['gt'].items()}, epoch=epoch)
            assert bool(flag) == want['flag'], (name, epoch)
            _check_loss_against(want['parts'], want['loss'], total, parts)
            assert set(parts) == set(want['parts']), (name, epoch)


def test_models_build_from_the_shipped_yaml_loss_sections():
    """models/baseline/lstm_stitch_tags.yaml lists `stitch, free_class`; the reference builds
    model_class(data_config, NN, NN['loss']) (nn/experiment.py:233), so construction must not raise."""
    import garment_pattern_estimation_b200 as g
    dc, nc, _ = _att()
    dc['standardize'] = torch.load(os.path.join(ROOT, 'tests', 'golden', 'n1_quality.pt'))['standardize']
    baseline_loss = {'loss_components': ['shape', 'loop', 'rotation', 'translation', 'stitch', 'free_class'],
                     'quality_components': ['shape', 'discrete', 'rotation', 'translation', 'stitch', 'free_class'],
                     'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False, 'epoch_with_stitches': 40,
                     'stitch_tags_margin': 0.3, 'stitch_hardnet_version': False, 'loop_loss_weight': 1.}
    nc = dict(nc, pattern_decoder='LSTMDecoderModule', pattern_encoding_size=250, pattern_n_layers=2)
    model = g.GarmentFullPattern3D(dict(dc), dict(nc), dict(baseline_loss))
    assert model.loss.config['loss_components'] == baseline_loss['loss_components']
    assert model.config['loss']['quality_components'] == baseline_loss['quality_components']
    B = 2
    gt = {'outlines': torch.randn(B, 23, 14, 4), 'rotations': torch.randn(B, 23, 4), 'translations': torch.randn(B, 23, 3),
          'num_edges': torch.full((B, 23), 4), 'num_panels': torch.full((B,), 23)}
    preds = {k: gt[k] + 0.1 for k in ('outlines', 'rotations', 'translations')}
    total, parts, _ = model.loss(preds, gt, epoch=3)            # before the stitch stage: the 4 regression terms only
    assert 'pattern_loss' in parts and 'stitch_similarity_loss' not in parts and torch.isfinite(total)
    with pytest.raises(NotImplementedError):                    # the stitch QUALITY metric is the one piece not built
        gt2 = dict(gt, stitches=torch.zeros(B, 2, 4, dtype=torch.long), num_stitches=torch.full((B,), 2),
                   free_edges_mask=torch.ones(B, 23, 14))
        preds2 = dict(preds, stitch_tags=torch.randn(B, 23, 14, 3), free_edges_mask=torch.randn(B, 23, 14))
        model.loss(preds2, gt2, epoch=50)
    with pytest.raises(ValueError):
        g.GarmentFullPattern3D(dict(dc), dict(nc), dict(baseline_loss, loss_components=['shape', 'no_such_term']))


def test_stitch_loss_terms_match_unmodified_reference_on_random_batches():
    """Runs only where /root/reference exists: more random batches (other seeds / batch sizes than the fixture) through the
    reference's ComposedPatternLoss and the vectorised one, all four loss-section variants of the fixture script."""
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('reference tree not present on this machine')
    import importlib.util
    from garment_pattern_estimation_b200.losses import ComposedPatternLoss
    spec = importlib.util.spec_from_file_location('mk_n1_stitch', os.path.join(ROOT, 'tests', 'golden', 'make_golden_n1_stitch.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    ref_stubs.import_reference()
    import metrics.composed_loss as cl
    dc, _, lc = ref_stubs.att_configs()
    st = dc['standardize']
    pad = -torch.tensor(st['gt_shift']['outlines']) / torch.tensor(st['gt_scale']['outlines'])
    for i, (name, cfg) in enumerate(mk.CASES.items()):
        cfg = dict(lc, **cfg)
        ref_loss, mine = cl.ComposedPatternLoss(dict(dc), dict(cfg)), ComposedPatternLoss(dict(dc), dict(cfg))
        for seed in (1, 2):
            preds, gt = mk.stitch_batch(2 + seed, 900 + 17 * i + seed, pad, permute='order' in name, rotate='origin' in name)
            t1, p1, f1 = ref_loss({k: v.clone() for k, v in preds.items()}, {k: v.clone() for k, v in gt.items()}, epoch=60)
            t2, p2, f2 = mine(preds, {k: v.clone() for k, v in gt.items()}, epoch=60)
            assert bool(f1) == bool(f2)
            _check_loss_against(p1, t1, t2, p2)


def test_unmodified_reference_trainer_drives_flat_data_parallel_on_cpu():
    """Host-side half of the swap-A/C GPU test (tests/test_gpu_baseline_shapes.py): the UNMODIFIED nn/trainer.py::Trainer.fit
    (Adam + OneCycleLR + validation + checkpoint dict) runs against parallel.FlatDataParallel and a duck-typed experiment;
    here with the reference's own blocks on the restated CPU operators (the B200 blocks need a GPU)."""
    import subprocess
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('reference tree not present on this machine')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import test_gpu_baseline_shapes as t
    env = dict(os.environ, NT_SWAP_A_DRY='1')
    out = subprocess.run([sys.executable, '-c', t._SWAP_A % {'root': ROOT}], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'trainer ran' in out.stdout
