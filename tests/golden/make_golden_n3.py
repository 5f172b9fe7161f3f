"""Golden fixtures for SURVEY.md section 8f row N3 -- the stage-2 stitch model (run in the build container only):

    python tests/golden/make_golden_n3.py

Values come from the UNMODIFIED reference class nets.StitchOnEdge3DPairs (nn/nets.py:303-353) and its ComposedLoss
(nn/metrics/composed_loss.py:10-127), imported through ``oracle.ref_stubs``:

  tests/golden/n3_stitch.pt   'ckpt':  shipped weights models/att/neural_tailor_stitch_model.pth, eval mode, pairs [3, 400, 16]:
                                        logits, BCE loss and the quality metrics (accuracy, stitch precision / recall);
                              'train': random init (seed 916143406), train mode: logits, loss, gradient digests, BN buffers.
  The shipped state_dict is small (330 KB) and stored inside the fixture.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_stubs  # noqa: E402
from make_golden import SEED_INIT, grad_digest, state_checksum  # noqa: E402


def main():
    nets, _ = ref_stubs.import_reference()
    data_config = {'element_size': 16}
    g = torch.Generator().manual_seed(97531)
    pairs = torch.randn(3, 400, 16, generator=g)
    gt = (torch.rand(3, 400, generator=g) < 0.4)

    ref = nets.StitchOnEdge3DPairs(dict(data_config), {}, {})
    sd = ref_stubs.stitch_checkpoint_state()
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    with torch.no_grad():
        logits = ref(pairs)
        loss, parts, _ = ref.loss(logits, gt)
    ckpt = {'state': sd, 'state_checksum': state_checksum(sd), 'logits': logits.clone(), 'loss': loss.clone(),
            'parts': {k: torch.as_tensor(v).clone() for k, v in parts.items()}}

    torch.manual_seed(SEED_INIT)
    ref = nets.StitchOnEdge3DPairs(dict(data_config), {}, {})
    init_sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref.train()
    logits = ref(pairs)
    loss, parts, _ = ref.loss(logits, gt)
    loss.backward()
    train = {'state_checksum': state_checksum(init_sd), 'logits': logits.detach().clone(), 'loss': loss.detach().clone(),
             'parts': {k: torch.as_tensor(v).detach().clone() for k, v in parts.items()},
             'grads': {n: grad_digest(p.grad) for n, p in ref.named_parameters()},
             'bn_buffers_after_train': {k: v.clone() for k, v in ref.state_dict().items()
                                        if 'running' in k or 'num_batches' in k}}
    out = os.path.join(HERE, 'n3_stitch.pt')
    torch.save({'seed_init': SEED_INIT, 'pairs': pairs, 'gt': gt, 'ckpt': ckpt, 'train': train}, out)
    print('n3_stitch.pt', os.path.getsize(out) // 1024, 'KiB', 'ckpt loss', float(ckpt['loss']), ckpt['parts'])


if __name__ == '__main__':
    main()
