"""Generate the committed golden fixtures from the UNMODIFIED reference code (run in the build container only).

    python tests/golden/make_golden.py

The reference's nn/nets.py + nn/net_blocks.py + nn/metrics/* are imported from /root/reference through
``oracle.ref_stubs`` (third-party operators restated in oracle/thirdparty.py) and executed on seeded inputs:

  tests/golden/att_random_init.pt   attention model (models/att/att.yaml NN section), default torch init under
                                    manual_seed(916143406) (att.yaml:147), B=2 x N=192 cloud: inputs, h0/c0, GT, train- and
                                    eval-mode outputs, the four loss terms, gradients, kNN indices of both EdgeConv layers.
                                    The weights are NOT stored: they are re-created from the seed (same torch build on the
                                    GPU box) and checked against a stored checksum.
  tests/golden/att_shipped_ckpt.pt  same model with the shipped weights models/att/neural_tailor_panels.pth, eval mode,
                                    B=2 x N=256: inputs, h0/c0, outputs, per-point features, attention weights.
  tests/golden/_ckpt/att_state.pt   the shipped model_state_dict itself (7.4 MB; git-ignored, travels to the GPU box
                                    with the snapshot) so the GPU tests can load it where /root/reference does not exist.
  tests/golden/knn_kat.pt           kNN known-answer cases from the oracle C restatement, incl. the tie-heavy input of the
                                    reference's own smoke block (nn/net_blocks.py:503-511: arange(1, 601).view(2, -1, 3)).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import knn as oknn  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import ref_stubs  # noqa: E402

SEED_INIT = 916143406      # models/att/att.yaml:147
SEED_STATE = 7             # LSTM h0/c0 draw


def state_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))


def grad_digest(g, samples=512):
    """Compact fingerprint of a gradient tensor: L2 norm, sum, and a strided sample of its entries."""
    flat = g.detach().reshape(-1)
    pick = torch.linspace(0, flat.numel() - 1, min(samples, flat.numel())).long()
    return {'norm': flat.double().norm().item(), 'sum': flat.double().sum().item(), 'index': pick,
            'values': flat[pick].clone()}


def run_reference(model, x, train, seed_state=SEED_STATE):
    """Forward of the unmodified reference model; returns outputs, the h0/c0 it drew, and kNN indices per layer."""
    captured = {}
    model.train(train)
    torch.manual_seed(seed_state)
    # replay the two draws the reference makes inside LSTMDecoderModule.forward (nn/net_blocks.py:391-392)
    rows = x.shape[0] * 23
    h0 = om.init_state(3, rows, 250)
    c0 = om.init_state(3, rows, 250)
    torch.manual_seed(seed_state)
    out = model(x)
    captured['h0'], captured['c0'] = h0, c0
    return out, captured


def knn_per_layer(model, x):
    """kNN indices of both DynamicEdgeConv layers for the given (eval-mode) model, via the oracle."""
    B, N = x.shape[:2]
    feats = x.reshape(B * N, 3)
    batch = torch.arange(B).repeat_interleave(N)
    idxs = []
    with torch.no_grad():
        for conv in model.feature_extractor.conv_layers:
            idxs.append(oknn.knn_indices(feats.reshape(B, N, -1), conv.k))
            feats = conv(feats, batch)
    return idxs


def main():
    nets, _ = ref_stubs.import_reference()
    dc, nc, lc = ref_stubs.att_configs()
    os.makedirs(os.path.join(HERE, '_ckpt'), exist_ok=True)

    # ---------------- random init, train + eval, with loss and grads
    torch.manual_seed(SEED_INIT)
    ref = nets.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc))
    ref.loss.with_quality_eval = False
    B, N = 2, 192
    x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(1234))
    gt = om.synthetic_ground_truth(B, seed=11)
    init_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    out_train, cap = run_reference(ref, x, train=True)
    loss, parts, _ = ref.loss(out_train, {k: v.clone() for k, v in gt.items()}, epoch=0)
    loss.backward()
    grads = {n: grad_digest(p.grad) for n, p in ref.named_parameters() if p.grad is not None}
    after_sd = {k: v.clone() for k, v in ref.state_dict().items() if 'running' in k or 'num_batches' in k}
    # eval pass on a FRESH copy of the initial weights (the train pass moved the BN running statistics)
    ref.load_state_dict(init_sd)
    out_eval, _ = run_reference(ref, x, train=False)
    idxs = knn_per_layer(ref, x)
    torch.save({
        'seed_init': SEED_INIT, 'state_checksum': state_checksum(init_sd), 'x': x, 'gt': gt,
        'h0': cap['h0'], 'c0': cap['c0'],
        'out_train': {k: v.detach().clone() for k, v in out_train.items()},
        'loss': loss.detach(), 'loss_parts': {k: v.detach() for k, v in parts.items()},
        'grads': grads, 'bn_buffers_after_train': after_sd,
        'out_eval': {k: v.detach().clone() for k, v in out_eval.items()},
        'knn_idx_eval': idxs,
    }, os.path.join(HERE, 'att_random_init.pt'))

    # ---------------- shipped checkpoint, eval
    sd = ref_stubs.att_checkpoint_state()
    torch.save(sd, os.path.join(HERE, '_ckpt', 'att_state.pt'))
    ref.load_state_dict(sd, strict=True)
    ref.save_att_weights = True
    B, N = 2, 256
    x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(4321))
    out_eval, cap = run_reference(ref, x, train=False)
    with torch.no_grad():
        _, feats, _ = ref.feature_extractor(x, False)
    idxs = knn_per_layer(ref, x)
    torch.save({
        'state_checksum': state_checksum(sd), 'x': x, 'h0': cap['h0'], 'c0': cap['c0'],
        'out_eval': {k: v.detach().clone() for k, v in out_eval.items()},
        'point_features': feats.clone(), 'knn_idx': idxs,
    }, os.path.join(HERE, 'att_shipped_ckpt.pt'))

    # ---------------- kNN known-answer cases
    cases = {}
    kat = torch.arange(1, 601, dtype=torch.float32).view(2, -1, 3)            # nn/net_blocks.py:505-506
    cases['reference_smoke_collinear_k5'] = dict(x=kat, k=5)
    g = torch.Generator().manual_seed(99)
    cases['randn_3d_k5'] = dict(x=torch.randn(3, 257, 3, generator=g), k=5)
    cases['randn_150d_k5'] = dict(x=torch.randn(2, 300, 150, generator=g), k=5)
    cases['randn_7d_k16'] = dict(x=torch.randn(2, 500, 7, generator=g), k=16)
    dup = torch.randn(1, 64, 3, generator=g).repeat(1, 4, 1)                  # every point 4 times -> exact ties
    cases['duplicates_k8'] = dict(x=dup, k=8)
    grid = torch.stack(torch.meshgrid(torch.arange(8.), torch.arange(8.), torch.arange(4.), indexing='ij'), -1)
    cases['integer_grid_k7'] = dict(x=grid.view(1, -1, 3), k=7)
    cases['n_equals_k'] = dict(x=torch.randn(4, 5, 3, generator=g), k=5)
    for name, c in cases.items():
        idx, dist = oknn.knn_indices(c['x'], c['k'], return_dist=True)
        c['idx'], c['dist'] = idx, dist
    torch.save(cases, os.path.join(HERE, 'knn_kat.pt'))
    for f in ('att_random_init.pt', 'att_shipped_ckpt.pt', 'knn_kat.pt', '_ckpt/att_state.pt'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
