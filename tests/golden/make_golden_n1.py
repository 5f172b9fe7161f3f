"""Golden fixture for SURVEY.md section 8f row N1 (quality metrics of the pattern loss; run in the build container only):

    python tests/golden/make_golden_n1.py

tests/golden/n1_quality.pt holds hand-built prediction / ground-truth batches and what the UNMODIFIED reference
``ComposedPatternLoss`` (nn/metrics/composed_loss.py, with nn/metrics/metrics.py and nn/metrics/losses.py) returns for them with
loss_components [shape, loop, rotation, translation] and quality_components [shape, discrete, rotation, translation]
(models/att/att.yaml:124-138).  The cases cover: perfectly predicted patterns, wrong edge counts, spurious / missing panels,
open edge loops, and a batch without a single correct pattern (the reference reports ``None`` for the ``corr_*`` metrics there).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import model as om  # noqa: E402
from oracle import ref_stubs  # noqa: E402


def closed_loop_gt(B, seed):
    """synthetic GT whose live edges form closed loops (last live edge closes the polygon), like real panels."""
    gt = om.synthetic_ground_truth(B, seed=seed)
    outl, ne = gt['outlines'], gt['num_edges']
    for b in range(B):
        for p in range(outl.shape[1]):
            n = int(ne[b, p])
            if n >= 3:
                outl[b, p, n - 1, :2] = -outl[b, p, :n - 1, :2].sum(0)
    return gt


def build_case(kind, seed, pad):
    B = 6
    gt = closed_loop_gt(B, seed)
    g = torch.Generator().manual_seed(seed + 1)
    outl = gt['outlines'].clone()
    ne = gt['num_edges']
    L = outl.shape[2]
    live = torch.arange(L)[None, None, :] < ne[..., None]
    pred = torch.where(live[..., None], outl, pad.expand_as(outl).clone())          # padding rows = standardised pad vector
    pred = pred + 0.004 * torch.randn(pred.shape, generator=g)                       # well inside the 0.07 padding tolerance
    rot = gt['rotations'] + 0.05 * torch.randn(gt['rotations'].shape, generator=g)
    tr = gt['translations'] + 0.05 * torch.randn(gt['translations'].shape, generator=g)
    if kind == 'mixed':
        # pattern 1: one panel gets an extra edge (wrong edge count, panel count still right)
        p = int((ne[1] >= 3).nonzero()[0]) if bool((ne[1] >= 3).any()) and int(ne[1].max()) < L else 0
        n = int(ne[1, p])
        if 3 <= n < L:
            pred[1, p, n] = torch.tensor([0.5, -0.5, 0.1, 0.1])
        # pattern 2: a spurious panel in an empty slot (wrong panel count)
        empty = (ne[2] < 3).nonzero()
        if len(empty):
            q = int(empty[0])
            pred[2, q, :3] = torch.tensor([[0.3, 0.0, 0.0, 0.0], [0.0, 0.3, 0.0, 0.0], [-0.3, -0.3, 0.0, 0.0]])
        # pattern 3: an open loop (first edge stretched) -> one more edge is counted
        p = int((ne[3] >= 3).nonzero()[0]) if bool((ne[3] >= 3).any()) else 0
        pred[3, p, 0, :2] += 1.0
        # pattern 4: a missing panel (all rows padded)
        p = int((ne[4] >= 3).nonzero()[0]) if bool((ne[4] >= 3).any()) else 0
        pred[4, p] = pad + 0.004 * torch.randn(L, 4, generator=g)
    elif kind == 'none_correct':
        for b in range(B):                                                           # every pattern loses its first panel
            p = int((ne[b] >= 3).nonzero()[0])
            pred[b, p] = pad + 0.004 * torch.randn(L, 4, generator=g)
    elif kind == 'noisy':
        pred = pred + 0.05 * torch.randn(pred.shape, generator=g)
    preds = {'outlines': pred, 'rotations': rot, 'translations': tr}
    return preds, gt


def main():
    nets, _ = ref_stubs.import_reference()
    import metrics.composed_loss as cl                                              # the reference's own module
    dc, _, lc = ref_stubs.att_configs()
    lc = dict(lc)
    lc.update(loss_components=['shape', 'loop', 'rotation', 'translation'],
              quality_components=['shape', 'discrete', 'rotation', 'translation'])
    ref_loss = cl.ComposedPatternLoss(dict(dc), dict(lc))
    st = dc['standardize']
    pad = -torch.tensor(st['gt_shift']['outlines']) / torch.tensor(st['gt_scale']['outlines'])
    cases = {}
    for kind, seed in (('perfect', 31), ('mixed', 32), ('none_correct', 33), ('noisy', 34)):
        preds, gt = build_case(kind, seed, pad)
        total, parts, flag = ref_loss({k: v.clone() for k, v in preds.items()}, {k: v.clone() for k, v in gt.items()}, epoch=3)
        cases[kind] = {'preds': preds, 'gt': gt, 'loss': total.clone(),
                       'parts': {k: (None if v is None else torch.as_tensor(v).clone()) for k, v in parts.items()}}
        print(kind, float(total), {k: (None if v is None else round(float(v), 5)) for k, v in parts.items()})
    out = os.path.join(HERE, 'n1_quality.pt')
    torch.save({'standardize': st, 'loss_config': lc, 'cases': cases}, out)
    print('n1_quality.pt', os.path.getsize(out) // 1024, 'KiB')

    # ---- GT panel-order / edge-origin matching (the reference's DEFAULT loss configuration)
    matching = {}
    for name, extra in (('origin', dict(panel_origin_invariant_loss=True, panel_order_inariant_loss=False)),
                        ('order_placement_origin', dict(panel_origin_invariant_loss=True, panel_order_inariant_loss=True,
                                                        order_by='placement')),
                        ('order_shape_translation', dict(panel_origin_invariant_loss=False, panel_order_inariant_loss=True,
                                                         order_by='shape_translation'))):
        cfg = dict(lc, **extra)
        ref_loss = cl.ComposedPatternLoss(dict(dc), dict(cfg))
        g = torch.Generator().manual_seed(4242)
        B = 4
        gt = closed_loop_gt(B, 77)
        gt['empty_panels_mask'] = gt['num_edges'] == 0
        live = torch.arange(14)[None, None, :] < gt['num_edges'][..., None]
        outl = torch.where(live[..., None], gt['outlines'], pad.expand_as(gt['outlines']).clone())
        perm = torch.stack([torch.randperm(23, generator=g) for _ in range(B)])
        pred = torch.gather(outl, 1, perm[..., None, None].expand_as(outl)).clone()
        ne = torch.gather(gt['num_edges'], 1, perm)
        for b in range(B):
            for p_ in range(23):
                n = int(ne[b, p_])
                if n >= 3:
                    s0 = int(torch.randint(0, n, (1,), generator=g))
                    pred[b, p_, :n] = torch.roll(pred[b, p_, :n], -s0, dims=0)
        preds = {'outlines': pred + 0.01 * torch.randn(pred.shape, generator=g),
                 'rotations': torch.gather(gt['rotations'], 1, perm[..., None].expand(B, 23, 4)) + 0.01 * torch.randn(B, 23, 4, generator=g),
                 'translations': torch.gather(gt['translations'], 1, perm[..., None].expand(B, 23, 3)) + 0.01 * torch.randn(B, 23, 3, generator=g)}
        total, parts, flag = ref_loss({k: v.clone() for k, v in preds.items()}, {k: v.clone() for k, v in gt.items()}, epoch=3)
        matching[name] = {'loss_config': cfg, 'preds': preds, 'gt': gt, 'loss': total.clone(), 'flag': bool(flag),
                          'parts': {k: (None if v is None else torch.as_tensor(v).clone()) for k, v in parts.items()}}
        print(name, float(total), {k: (None if v is None else round(float(v), 5)) for k, v in parts.items()})
    out = os.path.join(HERE, 'n1_matching.pt')
    torch.save({'standardize': st, 'cases': matching}, out)
    print('n1_matching.pt', os.path.getsize(out) // 1024, 'KiB')


if __name__ == '__main__':
    main()
