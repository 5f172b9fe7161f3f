"""CPU tests of the ORACLE: it must reproduce the committed golden vectors (generated from the unmodified reference code,
tests/golden/make_golden.py) and, where the reference tree is present, the reference itself bit-for-bit."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_close


def test_knn_oracle_reproduces_known_answers(golden_dir):
    from oracle import knn as oknn
    cases = torch.load(os.path.join(golden_dir, 'knn_kat.pt'))
    assert 'reference_smoke_collinear_k5' in cases
    for name, c in cases.items():
        idx, dist = oknn.knn_indices(c['x'], c['k'], return_dist=True)
        assert torch.equal(idx, c['idx']), name
        assert torch.equal(dist, c['dist']), name          # same fmaf chain -> bit-identical distances
        idx_mt = oknn.knn_indices(c['x'], c['k'], nthreads=4)
        assert torch.equal(idx_mt, c['idx']), name + ' (threaded)'


def test_knn_oracle_collinear_tie_rule():
    """the reference's own smoke input (nn/net_blocks.py:505-506): equispaced collinear points => exact distance ties;
    the lower index must come first and the query itself is its own nearest neighbour."""
    from oracle import knn as oknn
    x = torch.arange(1, 601, dtype=torch.float32).view(2, -1, 3)
    idx = oknn.knn_indices(x, 5)
    assert idx.shape == (2, 100, 5)
    assert idx[0, 50].tolist() == [50, 49, 51, 48, 52]
    assert idx[0, 0].tolist() == [0, 1, 2, 3, 4] and idx[1, 99].tolist() == [99, 98, 97, 96, 95]


def test_knn_oracle_matches_float64_brute_force():
    """independent check of the neighbour SETS with float64 distances on inputs without near-ties."""
    from oracle import knn as oknn
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 120, 6)).astype(np.float32)
    idx = oknn.knn_indices(x, 7).numpy()
    d = ((x[:, :, None, :].astype(np.float64) - x[:, None, :, :].astype(np.float64)) ** 2).sum(-1)
    want = np.argsort(d, axis=-1, kind='stable')[:, :, :7]
    assert (np.sort(idx, -1) == np.sort(want, -1)).all()


def test_knn_oracle_argument_errors():
    from oracle import knn as oknn
    with pytest.raises(RuntimeError):
        oknn.knn_indices(torch.randn(1, 10, 3), 0)


def test_sparsemax_oracle_properties():
    from oracle.thirdparty import Sparsemax
    z = torch.randn(64, 23, dtype=torch.float64, requires_grad=True)
    p = Sparsemax(dim=1)(z)
    assert torch.allclose(p.sum(-1), torch.ones(64, dtype=torch.float64))
    assert bool((p >= 0).all()) and bool((p == 0).any())
    # projection property: adding a constant to a row does not change the output
    assert torch.allclose(Sparsemax(dim=1)(z + 3.0), p)
    assert torch.autograd.gradcheck(lambda t: Sparsemax(dim=1)(t), (z[:4],), atol=1e-6)


def _golden_model(gold):
    from oracle import model as om
    dc = dict(om.ATT_DATA_CONFIG)
    torch.manual_seed(gold['seed_init'])
    return om.OracleSegmentPattern3D(dc, dict(om.ATT_NN_CONFIG), {})


def test_oracle_reproduces_reference_golden_random_init(golden_dir):
    from oracle import model as om
    gold = torch.load(os.path.join(golden_dir, 'att_random_init.pt'))
    model = _golden_model(gold)
    chk = float(sum(v.double().abs().sum() for v in model.state_dict().values() if v.dtype.is_floating_point))
    assert abs(chk - gold['state_checksum']) <= 1e-9 * gold['state_checksum']
    model.train()
    out = model(gold['x'], lstm_state=(gold['h0'], gold['c0']))
    for key, want in gold['out_train'].items():
        assert_close(out[key], want, tol=1e-5, what='oracle train ' + key)
    loss, parts = om.main_losses(out, gold['gt'])
    assert abs(float(loss) - float(gold['loss'])) <= 1e-5 * abs(float(gold['loss']))
    loss.backward()
    named = dict(model.named_parameters())
    for name, dig in gold['grads'].items():
        g = named[name].grad.reshape(-1)
        assert abs(float(g.double().norm()) - dig['norm']) <= 1e-4 * max(dig['norm'], 1e-12), name
    assert named['feature_extractor.lin.weight'].grad is None
    # the vectorised loss and pooling variants used by the CPU baseline agree with the loop forms
    loss_fast, _ = om.main_losses(out, gold['gt'], fast=True)
    assert abs(float(loss_fast) - float(loss)) <= 1e-6 * abs(float(loss))


def test_oracle_eval_and_knn_golden(golden_dir):
    from oracle import knn as oknn
    gold = torch.load(os.path.join(golden_dir, 'att_random_init.pt'))
    model = _golden_model(gold).eval()
    with torch.no_grad():
        out = model(gold['x'], lstm_state=(gold['h0'], gold['c0']))
        fast = model(gold['x'], lstm_state=(gold['h0'], gold['c0']), fast=True)
    for key, want in gold['out_eval'].items():
        assert_close(out[key], want, tol=1e-5, what='oracle eval ' + key)
        assert_close(fast[key], want, tol=1e-5, what='oracle eval (batched pooling) ' + key)
    assert torch.equal(oknn.knn_indices(gold['x'], 5), gold['knn_idx_eval'][0])


def test_oracle_loads_shipped_checkpoint_if_present(golden_dir):
    ck = os.path.join(golden_dir, '_ckpt', 'att_state.pt')
    if not os.path.exists(ck):
        pytest.skip('extracted checkpoint fixture absent')
    from oracle import model as om
    gold = torch.load(os.path.join(golden_dir, 'att_shipped_ckpt.pt'))
    model = om.OracleSegmentPattern3D(dict(om.ATT_DATA_CONFIG), dict(om.ATT_NN_CONFIG), {})
    model.load_state_dict(torch.load(ck), strict=True)
    model.eval()
    model.save_att_weights = True
    with torch.no_grad():
        out = model(gold['x'], lstm_state=(gold['h0'], gold['c0']))
    for key, want in gold['out_eval'].items():
        assert_close(out[key], want, tol=1e-5, what='oracle shipped ' + key)


def test_oracle_equals_unmodified_reference_when_available():
    """Runs only where /root/reference exists (the build container): the reference's own nets.py, executed through
    oracle.ref_stubs, must agree with the oracle restatement bit-for-bit on CPU."""
    from oracle import ref_stubs
    if not ref_stubs.reference_available():
        pytest.skip('reference tree not present on this machine')
    from oracle import model as om
    nets, _ = ref_stubs.import_reference()
    dc, nc, lc = ref_stubs.att_configs()
    torch.manual_seed(5)
    ref = nets.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc))
    torch.manual_seed(5)
    mine = om.OracleSegmentPattern3D(dict(dc), dict(nc), dict(lc))
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr.keys()) == list(sm.keys()) and all(torch.equal(sr[k], sm[k]) for k in sr)
    x = torch.randn(2, 150, 3, generator=torch.Generator().manual_seed(1))
    gt = om.synthetic_ground_truth(2)
    ref.loss.with_quality_eval = False
    for mode in (True, False):
        ref.train(mode)
        mine.train(mode)
        torch.manual_seed(7)
        o1 = ref(x)
        torch.manual_seed(7)
        o2 = mine(x)
        assert all(torch.equal(o1[k], o2[k]) for k in o1)
        l1, d1, _ = ref.loss(o1, dict(gt), epoch=0)
        l2, d2, _ = mine.loss(o2, dict(gt))
        assert float(l1) == float(l2)
    mine.load_state_dict(ref_stubs.att_checkpoint_state(), strict=True)


# ------------------------------------------------------------------------------------------------------------
# baseline model (global pool + pattern LSTM) and the non-local attention branch
# ------------------------------------------------------------------------------------------------------------
BASELINE_NN = {'feature_extractor': 'EdgeConvFeatures', 'conv_depth': 2, 'k_neighbors': 5, 'EConv_hidden': 200,
               'EConv_hidden_depth': 2, 'EConv_feature': 150, 'EConv_aggr': 'max', 'global_pool': 'mean',
               'skip_connections': False, 'graph_pooling': False, 'pool_ratio': 0.1, 'local_attention': True,
               'panel_decoder': 'LSTMDecoderModule', 'panel_encoding_size': 250, 'panel_hidden_size': 250,
               'panel_n_layers': 3, 'lstm_init': 'kaiming_normal_', 'pattern_decoder': 'LSTMDecoderModule',
               'pattern_encoding_size': 250, 'pattern_hidden_size': 250, 'pattern_n_layers': 2}       # lstm_stitch_tags.yaml:96-122


def test_oracle_reproduces_reference_golden_baseline_and_nonlocal(golden_dir):
    from oracle import model as om
    gold = torch.load(os.path.join(golden_dir, 'n2_variants.pt'))
    x, gt = gold['x'], gold['gt']
    for name, build in (('baseline', lambda: om.OracleFullPattern3D(dict(om.ATT_DATA_CONFIG), dict(BASELINE_NN), {})),
                        ('att_nonlocal', lambda: om.OracleSegmentPattern3D(
                            dict(om.ATT_DATA_CONFIG), dict(om.ATT_NN_CONFIG, local_attention=False), {}))):
        v = gold['variants'][name]
        torch.manual_seed(gold['seed_init'])
        model = build().train()
        chk = float(sum(t.double().abs().sum() for t in model.state_dict().values() if t.dtype.is_floating_point))
        assert abs(chk - v['state_checksum']) <= 1e-9 * v['state_checksum'], name
        kw = {'lstm_state': v['states']['panel']}
        if 'pattern' in v['states']:
            kw['pattern_lstm_state'] = v['states']['pattern']
        out = model(x, **kw)
        for key, want in v['out_train'].items():
            assert_close(out[key], want, tol=1e-5, what='oracle {} {}'.format(name, key))
        loss, _ = om.main_losses(out, gt)
        assert abs(float(loss) - float(v['loss'])) <= 1e-5 * abs(float(v['loss'])), name
        loss.backward()
        named = dict(model.named_parameters())
        for pname, dig in v['grads'].items():
            g = named[pname].grad.reshape(-1)
            assert abs(float(g.double().norm()) - dig['norm']) <= 1e-4 * max(dig['norm'], 1e-12), (name, pname)


def test_oracle_encoder_global_pools_golden(golden_dir):
    from oracle import model as om
    gold = torch.load(os.path.join(golden_dir, 'n2_variants.pt'))
    for pool, want in gold['encoder_pools'].items():
        torch.manual_seed(gold['seed_init'])
        enc = om.OracleEdgeConvFeatures(250, dict(BASELINE_NN, global_pool=pool)).eval()
        with torch.no_grad():
            got = enc(gold['x'])[0]
        assert_close(got, want, tol=1e-5, what='oracle encoder global_pool=' + pool)


def test_oracle_loads_shipped_baseline_checkpoint_if_present(golden_dir):
    ck = os.path.join(golden_dir, '_ckpt', 'baseline_state.pt')
    if not os.path.exists(ck):
        pytest.skip('extracted checkpoint fixture absent')
    from oracle import model as om
    gold = torch.load(os.path.join(golden_dir, 'n2_baseline_ckpt.pt'))
    model = om.OracleFullPattern3D(dict(om.ATT_DATA_CONFIG), dict(BASELINE_NN), {})
    model.load_state_dict(torch.load(ck), strict=True)
    model.eval()
    with torch.no_grad():
        out = model(gold['x'], lstm_state=gold['states']['panel'], pattern_lstm_state=gold['states']['pattern'])
    for key, want in gold['out_eval'].items():
        assert_close(out[key], want, tol=1e-5, what='oracle baseline ckpt ' + key)


# ------------------------------------------------------------------------------------------------------------
# stage-2 stitch model
# ------------------------------------------------------------------------------------------------------------
def test_oracle_reproduces_reference_golden_stitch_model(golden_dir):
    from oracle import model as om
    gold = torch.load(os.path.join(golden_dir, 'n3_stitch.pt'))
    model = om.OracleStitchOnEdge3DPairs()
    model.load_state_dict(gold['ckpt']['state'], strict=True)
    model.eval()
    with torch.no_grad():
        logits = model(gold['pairs'])
    assert_close(logits, gold['ckpt']['logits'], tol=1e-5, what='oracle stitch logits (shipped weights)')
    assert abs(float(model.loss(logits, gold['gt'])) - float(gold['ckpt']['loss'])) <= 1e-5 * float(gold['ckpt']['loss'])
    torch.manual_seed(gold['seed_init'])
    model = om.OracleStitchOnEdge3DPairs().train()
    logits = model(gold['pairs'])
    assert_close(logits, gold['train']['logits'], tol=1e-5, what='oracle stitch logits (train)')
    loss = model.loss(logits, gold['gt'])
    loss.backward()
    for name, dig in gold['train']['grads'].items():
        g = dict(model.named_parameters())[name].grad.reshape(-1)
        assert abs(float(g.double().norm()) - dig['norm']) <= 1e-4 * max(dig['norm'], 1e-12), name


def test_stitch_loss_and_quality_metrics_match_reference_golden(golden_dir):
    """The product's ComposedLoss (device-side metrics, no host syncs) on the reference's own logits."""
    from garment_pattern_estimation_b200.losses import ComposedLoss
    gold = torch.load(os.path.join(golden_dir, 'n3_stitch.pt'))
    loss_obj = ComposedLoss({'element_size': 16}, {'loss_components': ['edge_pair_class'],
                                                   'quality_components': ['edge_pair_class', 'edge_pair_stitch_recall']})
    for case in ('ckpt', 'train'):
        total, parts, flag = loss_obj(gold[case]['logits'], gold['gt'])
        assert flag is False and set(parts) == set(gold[case]['parts'])
        assert abs(float(total) - float(gold[case]['loss'])) <= 1e-6 * abs(float(gold[case]['loss']))
        for k, want in gold[case]['parts'].items():
            assert abs(float(parts[k]) - float(want)) <= 1e-6 * max(abs(float(want)), 1e-6), (case, k)
    with pytest.raises(NotImplementedError):
        ComposedLoss({}, {'loss_components': ['shape']})


# ------------------------------------------------------------------------------------------------------------
# specification of the LSTM decoder as GEMMs + cell updates (oracle/lstm_decomposed.py) vs torch.nn.LSTM
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('layers,R,T,H', [(3, 23, 14, 50), (2, 5, 23, 32), (1, 7, 3, 16)])
def test_decomposed_lstm_matches_torch_lstm_forward_and_backward(layers, R, T, H):
    from oracle import lstm_decomposed as ld
    torch.manual_seed(layers * 100 + R)
    dt = torch.float64
    lstm = torch.nn.LSTM(H, H, layers, batch_first=True).to(dt)
    x = torch.randn(R, H, dtype=dt, requires_grad=True)
    h0, c0 = torch.randn(layers, R, H, dtype=dt) * 0.1, torch.randn(layers, R, H, dtype=dt) * 0.1
    out, _ = lstm(x.unsqueeze(1).repeat(1, T, 1), (h0, c0))          # the reference's decoder input (net_blocks.py:388)
    g = torch.randn_like(out)
    out.backward(g)
    weights = [tuple(getattr(lstm, '{}_l{}'.format(n, l)).detach() for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'))
               for l in range(layers)]
    mine, saved = ld.lstm_forward(x.detach(), weights, h0, c0, T)     # time-major [T, R, H]
    assert torch.allclose(mine.transpose(0, 1), out.detach(), rtol=1e-10, atol=1e-12)
    dx, grads = ld.lstm_backward(g.transpose(0, 1).contiguous(), x.detach(), weights, saved, T)
    assert torch.allclose(dx, x.grad, rtol=1e-9, atol=1e-12)
    for l in range(layers):
        for got, name in zip(grads[l], ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')):
            want = getattr(lstm, '{}_l{}'.format(name, l)).grad
            assert torch.allclose(got, want, rtol=1e-9, atol=1e-12), (l, name)


# ------------------------------------------------------------------------------------------------------------
# PointNet++ operators (fps / radius / PointConv restated in oracle/)
# ------------------------------------------------------------------------------------------------------------
def test_pointnet_oracle_known_answers_and_reference_golden(golden_dir):
    from oracle import knn as oknn
    from oracle import thirdparty as tp
    gold = torch.load(os.path.join(golden_dir, 'pointnet.pt'))
    # farthest point sampling on 12 collinear equispaced points from point 0: the far end, then the ties resolve to the lower index
    line = torch.arange(12, dtype=torch.float32).view(1, 12, 1) * torch.tensor([1., 0., 0.]).view(1, 1, 3)
    assert oknn.fps_indices(line, 5).tolist() == [[0, 11, 5, 8, 2]] == gold['kat_fps_line'].tolist()
    # radius: more than max_num_neighbors points inside the ball -> the first 25 in index order, strict '<' on the squared distance
    kd = gold['kat_dense']
    nbr, cnt = oknn.radius_neighbours(kd['pos'], torch.tensor([[0, 7]], dtype=torch.int32), 0.3, 25)
    assert torch.equal(nbr, kd['nbr']) and cnt.tolist() == [[25, 25]] and nbr[0, 0].tolist() == list(range(25))
    two = torch.tensor([[[0., 0., 0.], [0.3, 0., 0.], [0.29, 0., 0.]]])
    nbr, cnt = oknn.radius_neighbours(two, torch.tensor([[0]], dtype=torch.int32), 0.3, 4)
    assert cnt.tolist() == [[2]] and nbr[0, 0].tolist() == [0, 2, -1, -1]            # the point AT distance r is excluded
    # the integer results stored with the reference golden
    x = gold['x']
    B, N = x.shape[0], x.shape[1]
    flat, batch = x.view(-1, 3), torch.arange(B).repeat_interleave(N)
    idx = tp.fps(flat, batch, ratio=0.2)
    assert torch.equal(idx, gold['fps_idx']) and idx.numel() == B * 80
    row, col = tp.radius(flat, flat[idx], 0.3, batch, batch[idx], max_num_neighbors=25)
    assert torch.equal(torch.stack([row, col]), gold['radius_row_col'])
    with pytest.raises(NotImplementedError):
        tp.fps(flat, batch, ratio=0.2, random_start=True)


def test_unmodified_reference_pointnet_reproduces_its_golden_when_available(golden_dir):
    from oracle import ref_stubs
    if not ref_stubs.reference_available() or not os.path.isfile(os.path.join(ref_stubs.REFERENCE_ROOT, 'nn', 'net_blocks.py')):
        pytest.skip('reference tree not present on this machine')
    _, nb = ref_stubs.import_reference()
    gold = torch.load(os.path.join(golden_dir, 'pointnet.pt'))
    model = nb.PointNetPlusPlus(gold['out_size'], dict(gold['config']))
    model.load_state_dict(gold['state'])
    model.eval()
    with torch.no_grad():
        y = model(gold['x'])
    assert torch.allclose(y, gold['eval_y'], rtol=1e-5, atol=1e-6)
