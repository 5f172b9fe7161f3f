"""Loss terms of the training step that are active in the shipped attention config (models/att/att.yaml:124):
shape / loop / rotation / translation.  Reference: nn/metrics/composed_loss.py:222-334, nn/metrics/losses.py:8-51.

The reference's PanelLoopLoss walks all B*23 panels in a Python loop with a device sync per panel (``if seq_len < 3`` on
a device scalar); here it is one masked reduction.  These are [B, 23, 14, 4]-sized tensors -- a few KB --
so they stay in torch (plumbing), off the kernel budget.

The no-grad quality metrics of the shipped configs ('shape', 'discrete', 'rotation', 'translation'; att.yaml:124-138) are
evaluated by ``metrics.PatternQuality`` (vectorised, one host read); the GT pre-processing of the reference's DEFAULT loss
config -- panel order matching and edge-loop origin matching (composed_loss.py:429-570, 621-703) -- is vectorised below.
The stitch-related terms of the shipped BASELINE config (models/baseline/lstm_stitch_tags.yaml:125: ``stitch``,
``free_class``; plus ``stitch_supervised``) are vectorised here as well: PatternStitchLoss (nn/metrics/losses.py:54-180, both
the extended-triplet and the HardNet negative term), the free-edge BCE, and the re-numbering of the stitch ground truth after
panel-order / edge-origin matching (composed_loss.py:590-618, 706-753).  Every component name of the reference is ACCEPTED at
construction (so a model builds from any shipped yaml, nn/experiment.py:233); the two pieces that are not built -- the
``segmentation`` loss (entmax.SparsemaxLoss) and the ``stitch`` QUALITY metric (PatternStitchPrecisionRecall) -- raise
NotImplementedError only when a call would actually evaluate them.
"""
import torch
import torch.nn.functional as F

from . import ops
from .metrics import PatternQuality

_MAIN = ('shape', 'loop', 'rotation', 'translation')
_STITCH = ('stitch', 'stitch_supervised', 'free_class')          # active from config['epoch_with_stitches'] on
_KNOWN = _MAIN + _STITCH + ('segmentation',)


def panel_loop_loss(outlines, num_edges=None, pad_xy=None):
    """mean over all panels x 2 coords of (sum_{e < num_edges} (xy_e - pad))^2; panels with < 3 edges contribute 0 but
    stay in the denominator (nn/metrics/losses.py:34-51)."""
    panels = outlines.reshape(-1, outlines.shape[-2], outlines.shape[-1])
    P, L = panels.shape[0], panels.shape[1]
    xy = panels[..., :2]
    if pad_xy is not None:
        xy = xy - pad_xy.to(xy.device, xy.dtype)
    if num_edges is not None:
        ne = num_edges.reshape(-1).to(panels.device)
        live = (torch.arange(L, device=panels.device)[None, :] < ne[:, None]) & (ne[:, None] >= 3)
        xy = xy * live[..., None].to(xy.dtype)
    elif L < 3:
        xy = xy * 0
    sums = xy.sum(dim=1)
    return (sums ** 2).sum() / (P * 2)



# ----------------------------------------------------------------------------------------------------------
# GT pre-processing of the reference loss (nn/metrics/composed_loss.py:429-570, 621-703), vectorised
# ----------------------------------------------------------------------------------------------------------
def match_panel_order(pred_features, gt_features):
    """Greedy panel assignment of ``_panel_order_match`` (composed_loss.py:530-570): repeatedly take the globally closest
    (predicted slot, GT panel) pair of every pattern and retire its row and column.  Returns permutation [B, P] (long):
    predicted slot p is matched with GT panel permutation[b, p].  The reference loops over the batch inside each of the P
    rounds; here a round is one argmin + two scatters for the whole batch."""
    B, P = pred_features.shape[0], gt_features.shape[1]
    dist = torch.cdist(pred_features.reshape(B, P, -1), gt_features.reshape(B, P, -1))
    perm = torch.full((B, P), -1, dtype=torch.long, device=pred_features.device)
    ar = torch.arange(B, device=pred_features.device)
    for _ in range(P):
        flat = dist.reshape(B, -1).argmin(dim=1)                 # first minimum, like the reference
        rows, cols = flat // P, flat % P
        perm[ar, rows] = cols
        dist[ar, rows, :] = float('inf')
        dist[ar, :, cols] = float('inf')
    return perm


def permute_panels(features, permutation):
    """``_feature_permute`` (composed_loss.py:573-589): gather along the panel dimension."""
    idx = permutation
    while idx.dim() < features.dim():
        idx = idx.unsqueeze(-1)
    return torch.gather(features, 1, idx.expand(features.shape))


def match_edge_origin(pred_outlines, gt_outlines, gt_num_edges):
    """``_batch_edge_order_match`` (composed_loss.py:656-703): for every panel, the rotation of the GT edge loop (first
    num_edges rows; padding rows stay) that is closest to the prediction.  Returns (rotated GT outlines, leading edge per
    panel [B*P]).  The reference tries the rotations one by one per panel in Python; here all L candidate rotations of all
    panels are one gather + one reduction (first minimum wins, as with the reference's strict ``<``)."""
    shape = gt_outlines.shape
    L = shape[-2]
    pred = pred_outlines.reshape(-1, L, shape[-1])
    gt = gt_outlines.reshape(-1, L, shape[-1])
    n = gt_num_edges.reshape(-1).to(gt.device).long()
    Q = gt.shape[0]
    e = torch.arange(L, device=gt.device)
    shift = e.view(1, L, 1)                                                   # candidate leading edge
    pos = e.view(1, 1, L)
    nn_ = n.view(Q, 1, 1).clamp_min(1)
    src = torch.where(pos < n.view(Q, 1, 1), (pos + shift) % nn_, pos)        # [Q, shift, L] source row of every row
    cand = torch.gather(gt.unsqueeze(1).expand(Q, L, L, shape[-1]), 2, src.unsqueeze(-1).expand(Q, L, L, shape[-1]))
    dist = ((pred.unsqueeze(1) - cand) ** 2).sum(dim=(-1, -2))                # [Q, shift]
    valid = (shift.view(1, L) < n.view(Q, 1)) | (shift.view(1, L) == 0)      # rotation 0 is always a candidate
    dist = torch.where(valid, dist, torch.full_like(dist, float('inf')))
    lead = dist.argmin(dim=1)
    chosen = torch.gather(cand, 1, lead.view(Q, 1, 1, 1).expand(Q, 1, L, shape[-1])).squeeze(1)
    return chosen.view(shape), lead


def shift_panel_rows(features, lead, num_edges):
    """``_per_panel_shift`` (composed_loss.py:706-724): per panel, rotate the first num_edges rows so that row ``lead`` comes
    first; padding rows and panels with < 3 edges stay.  features: [B, P, L, ...]; lead, num_edges: [B*P]."""
    shape = features.shape
    L = shape[2]
    f = features.reshape(shape[0] * shape[1], L, -1)
    n = num_edges.reshape(-1).to(f.device).long()
    ld = lead.reshape(-1).to(f.device).long()
    pos = torch.arange(L, device=f.device).view(1, L)
    live = (pos < n.view(-1, 1)) & (n.view(-1, 1) >= 3)
    src = torch.where(live, (pos + ld.view(-1, 1)) % n.view(-1, 1).clamp_min(1), pos)
    out = torch.gather(f, 1, src.unsqueeze(-1).expand_as(f))
    return out.view(shape)


def shift_stitches(stitches, num_stitches, lead, num_edges, max_panels, max_panel_len):
    """``_gt_stitches_shift`` (composed_loss.py:726-753): re-number the edge ids of the GT stitches after the edge loops were
    rotated by ``lead``.  stitches: [B, 2, S] pattern-level edge ids (padded past num_stitches[b]); returns a new tensor."""
    st = stitches.long()
    B, _, S = st.shape
    dev = st.device
    panel = st // max_panel_len
    inner = st - panel * max_panel_len
    gp = (torch.arange(B, device=dev).view(B, 1, 1) * max_panels + panel).clamp(0, lead.numel() - 1)
    ld = lead.reshape(-1).to(dev).long()[gp]
    n = num_edges.reshape(-1).to(dev).long()[gp]
    new_inner = torch.where(inner >= ld, inner - ld, n - (ld - inner))
    valid = torch.arange(S, device=dev).view(1, 1, S) < num_stitches.to(dev).long().view(B, 1, 1)
    return torch.where(valid, panel * max_panel_len + new_inner, st).to(stitches.dtype)


def permute_stitches(stitches, num_stitches, permutation, max_panel_len):
    """``_stitch_after_permute`` (composed_loss.py:590-618): edge ids after the panels were re-ordered (slot i now holds GT
    panel permutation[b, i])."""
    st = stitches.long()
    B, _, S = st.shape
    dev = st.device
    P = permutation.shape[1]
    inv = torch.empty_like(permutation)
    inv.scatter_(1, permutation, torch.arange(P, device=permutation.device).view(1, P).expand(B, P))
    panel = st // max_panel_len
    inner = st - panel * max_panel_len
    new_panel = torch.gather(inv.to(dev), 1, panel.clamp(0, P - 1).reshape(B, -1)).view(B, 2, S)
    valid = torch.arange(S, device=dev).view(1, 1, S) < num_stitches.to(dev).long().view(B, 1, 1)
    return torch.where(valid, new_panel * max_panel_len + inner, st).to(stitches.dtype)


def pattern_stitch_loss(stitch_tags, gt_stitches, gt_num_stitches, margin, hardnet):
    """``PatternStitchLoss`` (nn/metrics/losses.py:54-180) without the per-pattern / per-tag Python loops: tags of stitched
    edges are pulled together (squared distance, averaged per stitch then per pattern), every tag is pushed ``margin`` away
    from all tags that are not itself or its partner (extended triplet form: mean over the other tags; HardNet form: only
    the closest one).  Returns (loss, {'stitch_similarity_loss', 'stitch_neg_loss'})."""
    B = stitch_tags.shape[0]
    D = stitch_tags.shape[-1]
    dev = stitch_tags.device
    st = gt_stitches.long().to(dev)
    S = st.shape[-1]
    nums = gt_num_stitches.to(dev).long().view(B)
    flat = stitch_tags.reshape(B, -1, D)
    left = torch.gather(flat, 1, st[:, 0, :, None].expand(B, S, D))
    right = torch.gather(flat, 1, st[:, 1, :, None].expand(B, S, D))
    valid = torch.arange(S, device=dev).view(1, S) < nums.view(B, 1)                      # [B, S]
    sim = (((left - right) ** 2) * valid[..., None]).sum(dim=(1, 2)) / nums.to(flat.dtype)
    similarity = sim.sum() / B
    tags = torch.cat([left, right], dim=1)                                                # [B, 2S, D]
    valid2 = torch.cat([valid, valid], dim=1)
    dist = ((tags.unsqueeze(2) - tags.unsqueeze(1)) ** 2).sum(-1)                         # [B, 2S, 2S]
    ar = torch.arange(2 * S, device=dev)
    brother = torch.where(ar < S, ar + S, ar - S)
    excluded = (ar.view(-1, 1) == ar.view(1, -1)) | (brother.view(-1, 1) == ar.view(1, -1))
    pair_ok = valid2.unsqueeze(2) & valid2.unsqueeze(1) & ~excluded.unsqueeze(0)
    n_tags = (2 * nums).to(flat.dtype)
    if hardnet:
        closest = torch.where(pair_ok, dist, torch.full_like(dist, float('inf'))).min(dim=2).values
        per_tag = torch.clamp(margin - closest, min=0.)
    else:
        push = torch.clamp(margin - dist, min=0.)
        per_tag = torch.where(pair_ok, push, torch.zeros_like(push)).sum(dim=2) / n_tags.view(B, 1)
    neg = torch.where(valid2, per_tag, torch.zeros_like(per_tag)).sum() / n_tags.sum()
    return similarity + neg, {'stitch_similarity_loss': similarity, 'stitch_neg_loss': neg}


class ComposedPatternLoss:
    """Callable with the reference's interface: ``loss(preds, ground_truth, names=None, epoch=1000)`` ->
    ``(loss, loss_dict, structure_update_flag)``; ``.config``, ``.with_quality_eval``, ``.train()`` / ``.eval()``."""

    def __init__(self, data_config, in_config={}):
        self.config = {
            'loss_components': ['shape'], 'quality_components': [], 'loop_loss_weight': 1., 'segm_loss_weight': 0.05,
            'stitch_tags_margin': 0.3, 'epoch_with_stitches': 40, 'stitch_supervised_weight': 0.1,
            'stitch_hardnet_version': False, 'panel_origin_invariant_loss': True, 'panel_order_inariant_loss': True,
            'order_by': 'placement', 'epoch_with_order_matching': 0,
        }
        self.config.update(in_config)
        self.l_components = self.config['loss_components']
        self.q_components = self.config['quality_components']
        unknown = [c for c in self.l_components if c not in _KNOWN]
        if unknown:
            raise ValueError('unknown loss components {}'.format(unknown))
        if self.config['order_by'] not in ('placement', 'translation', 'shape_translation', 'stitches'):
            raise NotImplementedError('ComposedPatternLoss::Error::Ordering by requested feature <{}> is not implemented'.format(
                self.config['order_by']))
        self.with_quality_eval = True       # reference default (composed_loss.py:159); no-op without quality_components
        # the 'stitch' / 'free_class' quality entries only act from epoch_with_stitches on (composed_loss.py:269-273)
        self.quality = PatternQuality(data_config, [c for c in self.q_components if c not in ('stitch', 'free_class')])
        self.data_config = data_config
        self.training = False
        self.debug_prints = False
        self.max_panel_len = data_config['max_panel_len']
        self.max_pattern_size = data_config['max_pattern_len']
        self.pad_xy = None
        stats = data_config.get('standardize') if hasattr(data_config, 'get') else None
        if stats:     # padding vector -shift/scale (nn/metrics/eval_utils.py:80-87); zero for xy with the shipped stats
            shift, scale = stats['gt_shift']['outlines'], stats['gt_scale']['outlines']
            pad = torch.tensor([-shift[0] / scale[0], -shift[1] / scale[1]], dtype=torch.float32)
            self.pad_xy = pad if bool((pad != 0).any()) else None

    def _stitch_stage(self, epoch):
        return epoch >= self.config['epoch_with_stitches'] and any(c in self.l_components for c in _STITCH)

    def __call__(self, preds, ground_truth, names=None, epoch=1000):
        device = preds['outlines'].device
        gt = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in ground_truth.items()}
        if 'segmentation' in self.l_components:
            if self.config['panel_order_inariant_loss']:
                raise NotImplementedError('Order matching not supported for training with segmentation losses')
            raise NotImplementedError('the segmentation loss (entmax.SparsemaxLoss, composed_loss.py:322-331) is not built '
                                      'on the B200 path')
        stitch_stage = self._stitch_stage(epoch)
        # ---- GT pre-processing (composed_loss.py:239-253): panel order, then edge-loop origin
        if self.config['panel_order_inariant_loss']:
            gt = self._gt_order_match(preds, gt, epoch)
        if self.config['panel_origin_invariant_loss']:
            with torch.no_grad():
                gt = dict(gt)
                ne = gt['num_edges'].int().view(-1)
                gt['outlines'], lead = match_edge_origin(preds['outlines'], gt['outlines'], ne)
                if stitch_stage:                                     # composed_loss.py:631-645
                    gt['stitches'] = shift_stitches(gt['stitches'], gt['num_stitches'], lead, ne, self.max_pattern_size,
                                                    self.max_panel_len)
                    gt['free_edges_mask'] = shift_panel_rows(gt['free_edges_mask'], lead, ne)
                    if 'stitch_supervised' in self.l_components:
                        gt['stitch_tags'] = shift_panel_rows(gt['stitch_tags'], lead, ne)
        loss_dict = {}
        full = 0.
        fused = (preds['outlines'].is_cuda and preds['outlines'].dim() == 4 and 'rotations' in preds and 'translations' in preds
                 and all(preds[k].dtype == torch.float32 and preds[k].stride(-1) == 1 for k in ('outlines', 'rotations', 'translations')))
        if fused:
            # one kernel for the four terms and one for their gradients (csrc/train_step.cu) instead of ~50 element-wise launches
            pad = (0., 0.) if self.pad_xy is None else (float(self.pad_xy[0]), float(self.pad_xy[1]))
            parts = ops.pattern_loss(preds['outlines'], preds['rotations'], preds['translations'], gt['outlines'], gt['rotations'],
                                     gt['translations'], gt['num_edges'], self.l_components, self.config['loop_loss_weight'], pad)
            full = parts[0]
            for i, (name, key) in enumerate((('shape', 'pattern_loss'), ('loop', 'loop_loss'), ('rotation', 'rotation_loss'),
                                             ('translation', 'translation_loss'))):
                if name in self.l_components:
                    loss_dict[key] = parts[i + 1]
        else:
            if 'shape' in self.l_components:
                loss_dict['pattern_loss'] = F.mse_loss(preds['outlines'], gt['outlines'])
                full = full + loss_dict['pattern_loss']
            if 'loop' in self.l_components:
                loss_dict['loop_loss'] = panel_loop_loss(preds['outlines'], gt['num_edges'].int().view(-1), self.pad_xy)
                full = full + self.config['loop_loss_weight'] * loss_dict['loop_loss']
            if 'rotation' in self.l_components:
                loss_dict['rotation_loss'] = F.mse_loss(preds['rotations'], gt['rotations'])
                full = full + loss_dict['rotation_loss']
            if 'translation' in self.l_components:
                loss_dict['translation_loss'] = F.mse_loss(preds['translations'], gt['translations'])
                full = full + loss_dict['translation_loss']
        if stitch_stage:                                             # composed_loss.py:336-363
            if 'stitch' in self.l_components:
                st_loss, parts = pattern_stitch_loss(preds['stitch_tags'], gt['stitches'], gt['num_stitches'],
                                                     self.config['stitch_tags_margin'], self.config['stitch_hardnet_version'])
                loss_dict.update(parts)
                full = full + st_loss
            if 'stitch_supervised' in self.l_components:
                loss_dict['stitch_supervised_loss'] = F.mse_loss(preds['stitch_tags'], gt['stitch_tags'])
                full = full + self.config['stitch_supervised_weight'] * loss_dict['stitch_supervised_loss']
            if 'free_class' in self.l_components:
                loss_dict['free_edges_loss'] = F.binary_cross_entropy_with_logits(
                    preds['free_edges_mask'], gt['free_edges_mask'].to(preds['free_edges_mask'].dtype))
                full = full + loss_dict['free_edges_loss']
        if self.with_quality_eval and self.q_components:
            with torch.no_grad():
                quality, _ = self.quality(preds, gt, gt['num_edges'].int().view(-1), names)
                loss_dict.update(quality)
                if epoch >= self.config['epoch_with_stitches']:      # composed_loss.py:400-425
                    if 'stitch' in self.q_components:
                        raise NotImplementedError('the stitch precision/recall quality metric (nn/metrics/metrics.py:'
                                                  'PatternStitchPrecisionRecall) is not built on the B200 path; drop '
                                                  "'stitch' from quality_components or set with_quality_eval = False")
                    if 'free_class' in self.q_components:
                        free_class = torch.round(torch.sigmoid(preds['free_edges_mask']))
                        gt_mask = gt['free_edges_mask'].to(free_class.device)
                        loss_dict['free_edge_acc'] = (free_class == gt_mask).sum().float() / gt_mask.numel()
        # the loss structure changes at the epochs where the stitch terms / order matching start (composed_loss.py:278-280)
        structure_update = bool((epoch == self.config['epoch_with_stitches'] and any(c in self.l_components for c in _STITCH))
                                or (epoch == self.config['epoch_with_order_matching'] and self.config['panel_order_inariant_loss']))
        return full, loss_dict, structure_update

    def _gt_order_match(self, preds, gt, epoch):
        """composed_loss.py:429-527 without the stitch-related entries."""
        with torch.no_grad():
            order_by = self.config['order_by']
            if 'translations' not in preds or (order_by in ('placement',) and 'rotations' not in preds):
                raise ValueError('ComposedPatternLoss::Error::Ordering by {} requested but it is not predicted'.format(order_by))
            if order_by == 'placement':
                pf = torch.cat([preds['translations'], preds['rotations']], dim=-1)
                gf = torch.cat([gt['translations'], gt['rotations']], dim=-1)
            elif order_by == 'translation':
                pf, gf = preds['translations'], gt['translations']
            elif order_by == 'stitches':        # composed_loss.py:462-485
                if 'free_edges_mask' not in preds or 'rotations' not in preds:
                    raise ValueError('ComposedPatternLoss::Error::Ordering by stitches requested but free edges mask or '
                                     'placement are not predicted')
                pf = torch.cat([preds['translations'], preds['rotations']], dim=-1)
                gf = torch.cat([gt['translations'], gt['rotations']], dim=-1)
                if epoch >= self.config['epoch_with_stitches']:
                    B, P = pf.shape[0], pf.shape[1]
                    pm = torch.round(torch.sigmoid(preds['free_edges_mask'])).reshape(B, P, -1)
                    gm = gt['free_edges_mask'].reshape(B, P, -1).to(pm.dtype)
                    pf, gf = torch.cat([pf, pm], dim=-1), torch.cat([gf, gm], dim=-1)
            else:       # shape_translation
                B, P = preds['outlines'].shape[0], preds['outlines'].shape[1]
                pf = torch.cat([preds['translations'], preds['outlines'].reshape(B, P, -1)], dim=-1)
                gf = torch.cat([gt['translations'], gt['outlines'].reshape(B, P, -1)], dim=-1)
            B, P = pf.shape[0], gf.shape[1]
            if epoch < self.config['epoch_with_order_matching']:      # random order until matching starts (:538-543)
                perm = torch.stack([torch.randperm(P, dtype=torch.long, device=pf.device) for _ in range(B)])
            else:
                perm = match_panel_order(pf, gf)
            out = dict(gt)
            for key in ('outlines', 'num_edges', 'empty_panels_mask'):
                if key in gt:
                    out[key] = permute_panels(gt[key], perm)
            if 'rotation' in self.l_components:
                out['rotations'] = permute_panels(gt['rotations'], perm)
            if 'translation' in self.l_components:
                out['translations'] = permute_panels(gt['translations'], perm)
            if self._stitch_stage(epoch):                            # composed_loss.py:508-520
                out['stitches'] = permute_stitches(gt['stitches'], gt['num_stitches'], perm, self.max_panel_len)
                out['free_edges_mask'] = permute_panels(gt['free_edges_mask'], perm)
                if 'stitch_supervised' in self.l_components:
                    out['stitch_tags'] = permute_panels(gt['stitch_tags'], perm)
        return out

    def eval(self):
        self.training = False

    def train(self, mode=True):
        self.training = mode


class ComposedLoss:
    """Loss of the stage-2 stitch model (reference ``ComposedLoss``, nn/metrics/composed_loss.py:10-127): binary
    cross-entropy with logits on edge-pair scores (``edge_pair_class``) plus the no-grad quality metrics accuracy /
    stitch precision / stitch recall.  Same call interface as above; the metrics are evaluated as device tensors
    (no host synchronisation), with the reference's convention that an empty denominator yields 0."""

    def __init__(self, data_config, in_config={}):
        self.config = {'loss_components': [], 'quality_components': []}
        self.config.update(in_config)
        self.with_quality_eval = True
        self.training = False
        self.l_components = self.config['loss_components']
        self.q_components = self.config['quality_components']
        unsupported = [c for c in list(self.l_components) + list(self.q_components)
                       if c not in ('edge_pair_class', 'edge_pair_stitch_recall')]
        if unsupported:
            raise NotImplementedError('loss / quality components {} are not part of ComposedLoss'.format(unsupported))

    def __call__(self, preds, ground_truth, names=None, epoch=1000):
        gt = ground_truth.to(preds.device)
        loss_dict = {}
        full = 0.
        if 'edge_pair_class' in self.l_components:
            pair_loss = F.binary_cross_entropy_with_logits(preds.reshape(-1), gt.reshape(-1).to(torch.float32))
            loss_dict['edge_pair_class_loss'] = pair_loss
            full = full + pair_loss
        if self.with_quality_eval:
            with torch.no_grad():
                pred_cls = torch.round(torch.sigmoid(preds))
                if 'edge_pair_class' in self.q_components:
                    loss_dict['edge_pair_class_acc'] = (pred_cls == gt).sum().float() / gt.numel()
                if 'edge_pair_stitch_recall' in self.q_components:
                    is_stitch = gt == 1
                    correct = ((pred_cls == 1) & is_stitch).sum().float()
                    n_pred, n_gt = (pred_cls == 1).sum().float(), is_stitch.sum().float()
                    zero = torch.zeros((), device=preds.device)
                    loss_dict['stitch_precision'] = torch.where(n_pred > 0, correct / n_pred.clamp_min(1), zero)
                    loss_dict['stitch_recall'] = torch.where(n_gt > 0, correct / n_gt.clamp_min(1), zero)
        return full, loss_dict, False

    def eval(self):
        self.training = False

    def train(self, mode=True):
        self.training = mode
