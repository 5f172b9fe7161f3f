"""Golden fixture for the stitch-related terms of the pattern loss (run in the build container only):

    python tests/golden/make_golden_n1_stitch.py

tests/golden/n1_stitch_terms.pt holds synthetic prediction / ground-truth batches WITH stitch ground truth (stitches,
num_stitches, free_edges_mask, stitch_tags) and what the UNMODIFIED reference ``ComposedPatternLoss``
(nn/metrics/composed_loss.py + nn/metrics/losses.py:PatternStitchLoss) returns for them with the loss section of the shipped
BASELINE config (models/baseline/lstm_stitch_tags.yaml:125: shape, loop, rotation, translation, stitch, free_class), with the
HardNet variant, with ``stitch_supervised``, and with panel-order + edge-origin matching switched on (which re-numbers the
stitch ground truth, composed_loss.py:590-618, 726-753).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import model as om  # noqa: E402
from oracle import ref_stubs  # noqa: E402

CASES = {
    'baseline_yaml': dict(loss_components=['shape', 'loop', 'rotation', 'translation', 'stitch', 'free_class'],
                          quality_components=['shape', 'discrete', 'rotation', 'translation', 'free_class'],
                          panel_origin_invariant_loss=False, panel_order_inariant_loss=False),
    'hardnet_supervised': dict(loss_components=['shape', 'stitch', 'stitch_supervised', 'free_class'], quality_components=[],
                               stitch_hardnet_version=True, panel_origin_invariant_loss=False,
                               panel_order_inariant_loss=False),
    'origin_matching': dict(loss_components=['shape', 'loop', 'rotation', 'translation', 'stitch', 'free_class',
                                             'stitch_supervised'],
                            quality_components=['free_class'], panel_origin_invariant_loss=True,
                            panel_order_inariant_loss=False),
    'order_and_origin': dict(loss_components=['shape', 'loop', 'rotation', 'translation', 'stitch', 'free_class'],
                             quality_components=['shape', 'discrete', 'rotation', 'translation', 'free_class'],
                             panel_origin_invariant_loss=True, panel_order_inariant_loss=True, order_by='stitches'),
}


def stitch_batch(B, seed, pad, permute, rotate):
    """GT + predictions.  Stitches connect random pairs of LIVE edges; the free-edge mask marks every edge without a stitch.
    With `permute` the prediction has the GT panels in a random order, with `rotate` every edge loop starts at a random edge."""
    g = torch.Generator().manual_seed(seed)
    gt = om.synthetic_ground_truth(B, seed=seed + 1)
    P, L = gt['num_edges'].shape[1], gt['outlines'].shape[2]
    gt['empty_panels_mask'] = gt['num_edges'] == 0
    S = 28
    stitches = torch.zeros(B, 2, S, dtype=torch.long)
    nums = torch.zeros(B, dtype=torch.long)
    free = torch.ones(B, P, L)
    for b in range(B):
        live = [(p, e) for p in range(P) for e in range(int(gt['num_edges'][b, p]))]
        order = torch.randperm(len(live), generator=g).tolist()
        n = min(S - 3, len(live) // 2)
        n = max(2, int(torch.randint(max(2, n // 2), n + 1, (1,), generator=g)))
        for i in range(n):
            (p0, e0), (p1, e1) = live[order[2 * i]], live[order[2 * i + 1]]
            stitches[b, 0, i], stitches[b, 1, i] = p0 * L + e0, p1 * L + e1
            free[b, p0, e0] = free[b, p1, e1] = 0.
        nums[b] = n
    gt.update(stitches=stitches, num_stitches=nums, free_edges_mask=free,
              stitch_tags=torch.randn(B, P, L, 3, generator=g))
    live = torch.arange(L)[None, None, :] < gt['num_edges'][..., None]
    outl = torch.where(live[..., None], gt['outlines'], pad.expand_as(gt['outlines']).clone())
    perm = torch.stack([torch.randperm(P, generator=g) if permute else torch.arange(P) for _ in range(B)])

    def take(t):
        idx = perm
        while idx.dim() < t.dim():
            idx = idx.unsqueeze(-1)
        return torch.gather(t, 1, idx.expand(t.shape)).clone()

    pred_outl, ne = take(outl), take(gt['num_edges'])
    tags, mask_logits = take(gt['stitch_tags']), take((free - 0.5) * 6.)
    if rotate:
        for b in range(B):
            for p in range(P):
                n = int(ne[b, p])
                if n >= 3:
                    s0 = int(torch.randint(0, n, (1,), generator=g))
                    pred_outl[b, p, :n] = torch.roll(pred_outl[b, p, :n], -s0, dims=0)
                    tags[b, p, :n] = torch.roll(tags[b, p, :n], -s0, dims=0)
                    mask_logits[b, p, :n] = torch.roll(mask_logits[b, p, :n], -s0, dims=0)
    preds = {'outlines': pred_outl + 0.01 * torch.randn(pred_outl.shape, generator=g),
             'rotations': take(gt['rotations']) + 0.01 * torch.randn(B, P, 4, generator=g),
             'translations': take(gt['translations']) + 0.01 * torch.randn(B, P, 3, generator=g),
             'stitch_tags': tags + 0.15 * torch.randn(tags.shape, generator=g),
             'free_edges_mask': mask_logits + 0.5 * torch.randn(mask_logits.shape, generator=g)}
    return preds, gt


def main():
    ref_stubs.import_reference()
    import metrics.composed_loss as cl
    dc, _, lc = ref_stubs.att_configs()
    st = dc['standardize']
    pad = -torch.tensor(st['gt_shift']['outlines']) / torch.tensor(st['gt_scale']['outlines'])
    out = {'standardize': st, 'cases': {}}
    for i, (name, cfg) in enumerate(CASES.items()):
        cfg = dict(lc, **cfg)
        ref_loss = cl.ComposedPatternLoss(dict(dc), dict(cfg))
        preds, gt = stitch_batch(5, 300 + 10 * i, pad, permute='order' in name, rotate='origin' in name)
        runs = {}
        for epoch in (3, cfg['epoch_with_stitches'], 1000):
            total, parts, flag = ref_loss({k: v.clone() for k, v in preds.items()}, {k: v.clone() for k, v in gt.items()},
                                          epoch=epoch)
            runs[epoch] = {'loss': float(total), 'flag': bool(flag),
                           'parts': {k: (None if v is None else float(v)) for k, v in parts.items()}}
        out['cases'][name] = {'loss_config': cfg, 'preds': preds, 'gt': gt, 'runs': runs}
        print(name, {e: round(r['loss'], 5) for e, r in runs.items()})
    torch.save(out, os.path.join(HERE, 'n1_stitch_terms.pt'))


if __name__ == '__main__':
    main()
