"""Loss terms of the training step that are active in the shipped attention config (models/att/att.yaml:124):
shape / loop / rotation / translation.  Reference: nn/metrics/composed_loss.py:222-334, nn/metrics/losses.py:8-51.

The reference's PanelLoopLoss walks all B*23 panels in a Python loop with a device sync per panel (``if seq_len < 3`` on
a device scalar); here it is one masked reduction.  These are [B, 23, 14, 4]-sized tensors -- a few KB --
so they stay in torch (plumbing), off the kernel budget.

The no-grad quality metrics of the shipped configs ('shape', 'discrete', 'rotation', 'translation'; att.yaml:124-138) are
evaluated by ``metrics.PatternQuality`` (vectorised, one host read); the GT pre-processing of the reference's DEFAULT loss
config -- panel order matching and edge-loop origin matching (composed_loss.py:429-570, 621-703) -- is vectorised below.
Stitch, free-class and segmentation losses and the stitch quality metrics need the dataset's stitch ground truth and stay out
of scope: they raise NotImplementedError if requested.
"""
import torch
import torch.nn.functional as F

from .metrics import PatternQuality

_SUPPORTED = ('shape', 'loop', 'rotation', 'translation')


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
        unsupported = [c for c in self.l_components if c not in _SUPPORTED]
        if unsupported:
            raise NotImplementedError('loss components {} are outside the B200 hot path'.format(unsupported))
        if self.config['panel_order_inariant_loss'] and self.config['order_by'] not in ('placement', 'translation',
                                                                                      'shape_translation'):
            raise NotImplementedError("panel order matching by '{}' needs the stitch ground truth (outside the B200 hot "
                                      "path)".format(self.config['order_by']))
        self.with_quality_eval = True       # reference default (composed_loss.py:159); no-op without quality_components
        self.quality = PatternQuality(data_config, self.q_components)
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

    def __call__(self, preds, ground_truth, names=None, epoch=1000):
        device = preds['outlines'].device
        gt = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in ground_truth.items()}
        # ---- GT pre-processing (composed_loss.py:239-253): panel order, then edge-loop origin
        if self.config['panel_order_inariant_loss']:
            gt = self._gt_order_match(preds, gt, epoch)
        if self.config['panel_origin_invariant_loss']:
            with torch.no_grad():
                gt = dict(gt)
                gt['outlines'], _ = match_edge_origin(preds['outlines'], gt['outlines'], gt['num_edges'].int().view(-1))
        loss_dict = {}
        full = 0.
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
        if self.with_quality_eval and self.q_components:
            with torch.no_grad():
                quality, _ = self.quality(preds, gt, gt['num_edges'].int().view(-1), names)
                loss_dict.update(quality)
        # the loss structure changes at the epoch where order matching starts (composed_loss.py:278-280)
        structure_update = bool(self.config['panel_order_inariant_loss'] and epoch == self.config['epoch_with_order_matching'])
        return full, loss_dict, structure_update

    def _gt_order_match(self, preds, gt, epoch):
        """composed_loss.py:429-527 without the stitch-related entries."""
        with torch.no_grad():
            order_by = self.config['order_by']
            if 'translations' not in preds or (order_by == 'placement' and 'rotations' not in preds):
                raise ValueError('ComposedPatternLoss::Error::Ordering by {} requested but it is not predicted'.format(order_by))
            if order_by == 'placement':
                pf = torch.cat([preds['translations'], preds['rotations']], dim=-1)
                gf = torch.cat([gt['translations'], gt['rotations']], dim=-1)
            elif order_by == 'translation':
                pf, gf = preds['translations'], gt['translations']
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
