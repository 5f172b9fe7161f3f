"""Loss terms of the training step that are active in the shipped attention config (models/att/att.yaml:124):
shape / loop / rotation / translation.  Reference: nn/metrics/composed_loss.py:222-334, nn/metrics/losses.py:8-51.

The reference's PanelLoopLoss walks all B*23 panels in a Python loop with a device sync per panel (``if seq_len < 3`` on
a device scalar); here it is one masked reduction.  These are [B, 23, 14, 4]-sized tensors -- a few KB --
so they stay in torch (plumbing), off the kernel budget.

The no-grad quality metrics of the shipped configs ('shape', 'discrete', 'rotation', 'translation'; att.yaml:124-138) are
evaluated by ``metrics.PatternQuality`` (vectorised, one host read).  What the shipped att config does NOT enable (order /
origin matching, stitch, free-class and segmentation losses, stitch quality) is out of scope
and raises NotImplementedError if requested.
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
        if self.config['panel_origin_invariant_loss'] or self.config['panel_order_inariant_loss']:
            raise NotImplementedError('GT origin/order matching is outside the B200 hot path (att.yaml disables both)')
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
        return full, loss_dict, False

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
