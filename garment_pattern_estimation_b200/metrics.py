"""No-grad quality metrics of the pattern loss, vectorised on the device (first half).

Reference: nn/metrics/metrics.py -- ``NumbersInPanelsAccuracies`` (:95-182), ``PanelVertsL2`` (:185-281), ``UniversalL2``
(:284-321), wired by ``ComposedPatternLoss._main_quality_metrics`` (nn/metrics/composed_loss.py:365-398).  The reference walks
all B*23 panels (and every edge of every panel) in Python with a device synchronisation per ``if`` on a device scalar
(736 panels x 14 edges at B = 32); here every metric is a handful of masked tensor reductions over
[B, 23, 14, 4]-sized tensors (KB-sized, plain torch = plumbing) and the whole block needs ONE host read, and only to reproduce the
reference's ``None`` results (``corr_*`` metrics when no pattern of the batch has the right number of panels).
"""
import torch


def _pad_vector(stats):
    """eval_pad_vector (nn/metrics/eval_utils.py:80-87): padding row after standardisation = -shift / scale."""
    return -torch.tensor(stats['shift'], dtype=torch.float32) / torch.tensor(stats['scale'], dtype=torch.float32)


class NumbersInPanelsAccuracies:
    """Fraction of patterns with the right number of panels / of panels with the right number of edges."""

    def __init__(self, max_edges_in_panel, data_stats):
        self.max_panel_len = max_edges_in_panel
        self.pad_vector = _pad_vector(data_stats)
        # 3 cm per coordinate is a tolerable loop-closing error (metrics.py:109)
        self.panel_loop_threshold = torch.tensor([3., 3.]) / torch.tensor(data_stats['scale'], dtype=torch.float32)[:2]

    def __call__(self, predicted_outlines, gt_num_edges, gt_panel_nums, pattern_names=None):
        dev = predicted_outlines.device
        B, P = predicted_outlines.shape[0], predicted_outlines.shape[1]
        pad, thr = self.pad_vector.to(dev), self.panel_loop_threshold.to(dev)
        # rows that are NOT padding (torch.isclose(pred, pad, atol=0.07), default rtol)
        is_pad = (predicted_outlines - pad).abs() <= 0.07 + 1e-5 * pad.abs()
        num_edges = (~is_pad.all(dim=-1)).sum(dim=-1)                                   # [B, P]
        loop = predicted_outlines[..., :2].sum(dim=-2)                                  # [B, P, 2]
        num_edges = num_edges + (loop.abs() > thr).any(dim=-1).to(num_edges.dtype)      # an open loop needs one more edge
        real = num_edges >= 3                                                           # fewer edges = empty panel slot
        num_panels = real.sum(dim=-1)                                                   # [B]
        gt_ne = gt_num_edges.to(dev).view(B, P)
        correct_edges = (real & (num_edges == gt_ne)).sum(dim=-1).to(torch.float32)     # [B]
        gt_np = gt_panel_nums.to(dev).view(B)
        correct_len = num_panels == gt_np
        per_pattern = correct_edges / gt_np
        n_correct = correct_len.sum()
        return (correct_len.sum().to(torch.float32) / B,
                per_pattern.sum() / B,
                correct_len,
                torch.where(correct_len, per_pattern, torch.zeros_like(per_pattern)).sum() / n_correct)


class PanelVertsL2:
    """Mean vertex distance (cm) between predicted and GT panels after converting edge lists to vertices."""

    def __init__(self, max_edges_in_panel, data_stats):
        self.shift = torch.tensor(data_stats['shift'], dtype=torch.float32)
        self.scale = torch.tensor(data_stats['scale'], dtype=torch.float32)
        self.max_panel_len = max_edges_in_panel

    @staticmethod
    def _vertices(edges, live):
        """edges [Q, L, 4] (un-standardised), live [Q, L] bool.  Returns the end vertex and the curvature control point of every
        edge ([Q, L, 2] each; the loop starts at the origin) -- metrics.py:259-281 without the centring."""
        e = edges[..., :2] * live.unsqueeze(-1)
        v = torch.cumsum(e, dim=1)
        prev = v - e
        perp = torch.stack([-e[..., 1], e[..., 0]], dim=-1)
        c = prev + edges[..., 2:3] * e + edges[..., 3:4] * perp
        return v, c

    def __call__(self, predicted_outlines, gt_outlines, gt_num_edges, correct_mask=None):
        dev = predicted_outlines.device
        P = predicted_outlines.shape[1]
        L = predicted_outlines.shape[-2]
        pred = predicted_outlines.reshape(-1, L, predicted_outlines.shape[-1]) * self.scale.to(dev) + self.shift.to(dev)
        gt = gt_outlines.reshape(-1, L, gt_outlines.shape[-1]) * self.scale.to(dev) + self.shift.to(dev)
        n = gt_num_edges.to(dev).view(-1)
        live = torch.arange(L, device=dev)[None, :] < n[:, None]                        # GT edge count un-pads both
        valid = n >= 3
        count = (1 + 2 * n).to(torch.float32).unsqueeze(-1)                             # vertices per panel
        vg, cg = self._vertices(gt, live)
        vp, cp = self._vertices(pred, live)
        m = live.unsqueeze(-1).to(pred.dtype)
        mean_g = ((vg + cg) * m).sum(dim=1) / count                                     # centre of the 1 + 2n vertices
        mean_p = ((vp + cp) * m).sum(dim=1) / count
        off = (mean_g - mean_p).unsqueeze(1)
        d_v = (((vg - vp) - off) ** 2).sum(dim=-1).sqrt() * live
        d_c = (((cg - cp) - off) ** 2).sum(dim=-1).sqrt() * live
        d_0 = (off.squeeze(1) ** 2).sum(dim=-1).sqrt()                                  # the shared origin vertex
        per_panel = (d_0 + d_v.sum(dim=1) + d_c.sum(dim=1)) / count.squeeze(-1)
        zero = torch.zeros_like(per_panel)
        total = torch.where(valid, per_panel, zero).sum() / valid.sum()
        if correct_mask is None:
            return total, None, None
        pm = torch.repeat_interleave(correct_mask.to(dev), P) & valid
        return total, torch.where(pm, per_panel, zero).sum() / pm.sum(), pm.sum()


class UniversalL2:
    """Mean L2 distance of un-standardised [B, P, F] predictions (rotations / translations)."""

    def __init__(self, data_stats):
        self.shift = torch.tensor(data_stats['shift'], dtype=torch.float32)
        self.scale = torch.tensor(data_stats['scale'], dtype=torch.float32)

    def __call__(self, predicted, gt, correct_mask=None):
        dev = predicted.device
        P = predicted.shape[1]
        pred = predicted.reshape(-1, predicted.shape[-1]) * self.scale.to(dev) + self.shift.to(dev)
        g = gt.to(dev).reshape(-1, gt.shape[-1]) * self.scale.to(dev) + self.shift.to(dev)
        norms = ((g - pred) ** 2).sum(dim=1).sqrt()
        if correct_mask is None:
            return norms.mean(), None, None
        pm = torch.repeat_interleave(correct_mask.to(dev), P)
        return norms.mean(), torch.where(pm, norms, torch.zeros_like(norms)).sum() / pm.sum(), pm.sum()


class PatternQuality:
    """``ComposedPatternLoss._main_quality_metrics`` (composed_loss.py:365-398) for the components of the shipped configs:
    'discrete', 'shape', 'rotation', 'translation'.  Returns the reference's dict (same keys; ``corr_*`` entries are None when no
    pattern / panel qualifies, exactly like the reference) and the correct-pattern mask."""

    SUPPORTED = ('shape', 'discrete', 'rotation', 'translation')

    def __init__(self, data_config, components):
        self.components = list(components)
        unsupported = [c for c in self.components if c not in self.SUPPORTED]
        if unsupported:
            raise NotImplementedError('quality components {} are outside the B200 hot path'.format(unsupported))
        if not self.components:
            return
        stats = data_config['standardize']
        outl = {'shift': stats['gt_shift']['outlines'], 'scale': stats['gt_scale']['outlines']}
        L = data_config['max_panel_len']
        if 'shape' in self.components:
            self.shape = PanelVertsL2(L, outl)
        if 'discrete' in self.components:
            self.nums = NumbersInPanelsAccuracies(L, outl)
        if 'rotation' in self.components:
            self.rotation = UniversalL2({'shift': stats['gt_shift']['rotations'], 'scale': stats['gt_scale']['rotations']})
        if 'translation' in self.components:
            self.translation = UniversalL2({'shift': stats['gt_shift']['translations'],
                                            'scale': stats['gt_scale']['translations']})

    def __call__(self, preds, gt, gt_num_edges, names=None):
        out, counts = {}, {}
        mask = None
        if 'discrete' in self.components:
            acc_p, acc_e, mask, acc_e_corr = self.nums(preds['outlines'], gt_num_edges, gt['num_panels'], names)
            out.update(num_panels_accuracy=acc_p, num_edges_accuracy=acc_e, corr_num_edges_accuracy=acc_e_corr)
        if 'shape' in self.components:
            l2, corr, cnt = self.shape(preds['outlines'], gt['outlines'], gt_num_edges, mask)
            out.update(panel_shape_l2=l2, corr_panel_shape_l2=corr)
            counts['corr_panel_shape_l2'] = cnt
        if 'rotation' in self.components:
            l2, corr, cnt = self.rotation(preds['rotations'], gt['rotations'], mask)
            out.update(rotation_l2=l2, corr_rotation_l2=corr)
            counts['corr_rotation_l2'] = cnt
        if 'translation' in self.components:
            l2, corr, cnt = self.translation(preds['translations'], gt['translations'], mask)
            out.update(translation_l2=l2, corr_translation_l2=corr)
            counts['corr_translation_l2'] = cnt
        # the reference reports None (not NaN) for 'corr_*' metrics without any qualifying entry: one host read for all of them
        live = [k for k, c in counts.items() if c is not None]
        if live:
            empty = (torch.stack([counts[k] for k in live]) == 0).tolist()
            for k, is_empty in zip(live, empty):
                if is_empty:
                    out[k] = None
        return out, mask
