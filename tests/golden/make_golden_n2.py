"""Golden fixtures for SURVEY.md section 8f row N2 (run in the build container only, like make_golden.py):

    python tests/golden/make_golden_n2.py

All values come from the UNMODIFIED reference classes (nn/nets.py, nn/net_blocks.py, nn/metrics/*) imported through
``oracle.ref_stubs``:

  tests/golden/n2_baseline_ckpt.pt      nets.GarmentFullPattern3D (global mean pool -> feature_extractor.lin -> pattern LSTM ->
                                        panel LSTM) with the shipped weights models/baseline/lstm_stitch_tags.pth, eval mode,
                                        B=2 x N=256: inputs, the two pairs of LSTM initial states the reference drew, outputs.
  tests/golden/_ckpt/baseline_state.pt  the shipped model_state_dict (git-ignored; travels to the GPU box).
  tests/golden/n2_variants.pt           random-init (seed 916143406) train-mode runs with the 4-term loss and gradient digests:
                                          'baseline'      GarmentFullPattern3D, global_pool = mean
                                          'att_nonlocal'  GarmentSegmentPattern3D with local_attention = False (global encoding
                                                          concatenated to every point before the attention MLP, nn/nets.py:257-260)
                                        plus eval-mode EdgeConvFeatures encodings for global_pool in {mean, max, add}.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import model as om  # noqa: E402
from oracle import ref_stubs  # noqa: E402
from make_golden import SEED_INIT, SEED_STATE, grad_digest, state_checksum  # noqa: E402


def draw_states(B, with_pattern):
    """The draws the reference makes inside its forward, in order (nn/net_blocks.py:391-392 per decoder)."""
    torch.manual_seed(SEED_STATE)
    states = {}
    if with_pattern:
        states['pattern'] = (om.init_state(2, B, 250), om.init_state(2, B, 250))
    states['panel'] = (om.init_state(3, B * 23, 250), om.init_state(3, B * 23, 250))
    return states


def run(model, x, train, with_pattern):
    model.train(train)
    states = draw_states(x.shape[0], with_pattern)
    torch.manual_seed(SEED_STATE)
    return model(x), states


def main():
    nets, blocks = ref_stubs.import_reference()
    os.makedirs(os.path.join(HERE, '_ckpt'), exist_ok=True)

    # ---------------- shipped baseline checkpoint, eval
    dc, nc, lc = ref_stubs.baseline_configs()
    torch.manual_seed(SEED_INIT)
    ref = nets.GarmentFullPattern3D(dict(dc), dict(nc), dict(lc))
    sd = ref_stubs.baseline_checkpoint_state()
    torch.save(sd, os.path.join(HERE, '_ckpt', 'baseline_state.pt'))
    ref.load_state_dict(sd, strict=True)
    B, N = 2, 256
    x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(2468))
    with torch.no_grad():
        out, states = run(ref, x, train=False, with_pattern=True)
        enc = ref.forward_encode(x)
    torch.save({'state_checksum': state_checksum(sd), 'x': x, 'states': states, 'encoding': enc.clone(),
                'out_eval': {k: v.detach().clone() for k, v in out.items()}},
               os.path.join(HERE, 'n2_baseline_ckpt.pt'))

    # ---------------- random-init variants, train mode with loss and gradients
    variants = {}
    B, N = 2, 160
    x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(1357))
    gt = om.synthetic_ground_truth(B, seed=13)
    torch.manual_seed(SEED_INIT)
    ref = nets.GarmentFullPattern3D(dict(dc), dict(nc), dict(lc))
    ref.loss.with_quality_eval = False
    init_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    out, states = run(ref, x, train=True, with_pattern=True)
    loss, parts, _ = ref.loss(out, {k: v.clone() for k, v in gt.items()}, epoch=0)
    loss.backward()
    variants['baseline'] = {
        'state_checksum': state_checksum(init_sd), 'states': states,
        'out_train': {k: v.detach().clone() for k, v in out.items()}, 'loss': loss.detach(),
        'grads': {n: grad_digest(p.grad) for n, p in ref.named_parameters() if p.grad is not None}}

    adc, anc, alc = ref_stubs.att_configs()
    anc = dict(anc)
    anc['local_attention'] = False
    torch.manual_seed(SEED_INIT)
    ref = nets.GarmentSegmentPattern3D(dict(adc), dict(anc), dict(alc))
    ref.loss.with_quality_eval = False
    init_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    out, states = run(ref, x, train=True, with_pattern=False)
    loss, parts, _ = ref.loss(out, {k: v.clone() for k, v in gt.items()}, epoch=0)
    loss.backward()
    variants['att_nonlocal'] = {
        'state_checksum': state_checksum(init_sd), 'states': states,
        'out_train': {k: v.detach().clone() for k, v in out.items()}, 'loss': loss.detach(),
        'grads': {n: grad_digest(p.grad) for n, p in ref.named_parameters() if p.grad is not None}}

    pools = {}
    for pool in ('mean', 'max', 'add'):
        cfg = dict(nc)
        cfg['global_pool'] = pool
        torch.manual_seed(SEED_INIT)
        enc = blocks.EdgeConvFeatures(250, cfg).eval()
        with torch.no_grad():
            pools[pool] = enc(x)[0].clone()
    torch.save({'seed_init': SEED_INIT, 'x': x, 'gt': gt, 'variants': variants, 'encoder_pools': pools},
               os.path.join(HERE, 'n2_variants.pt'))
    for f in ('n2_baseline_ckpt.pt', 'n2_variants.pt', '_ckpt/baseline_state.pt'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
