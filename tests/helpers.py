"""Shared test helpers: tolerances and plain-torch references of single operators (run on whatever device the inputs
live on).  The per-operator references here restate the reference semantics exactly like oracle/thirdparty.py, but take
the kNN indices as an INPUT so that the float comparison is not polluted by neighbour flips."""
import torch

REL_TOL = 1e-3      # north_star: activations within 1e-3 relative of the reference path


def rel_err(a, b):
    """max |a - b| relative to the scale of b (max |b|); the tolerance stated in BASELINE.json is on this number."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-12)
    return float((a - b).abs().max() / denom)


def assert_close(a, b, tol=REL_TOL, what=''):
    e = rel_err(a, b)
    assert e <= tol, '{}: relative error {:.3e} > {:.1e}'.format(what, e, tol)


def torch_mlp(widths):
    import torch.nn as nn
    return nn.Sequential(*[nn.Sequential(nn.Linear(a, b), nn.ReLU(), nn.BatchNorm1d(b))
                           for a, b in zip(widths[:-1], widths[1:])])


def ref_edgeconv(x, idx_global, mlp):
    """DynamicEdgeConv message + max aggregation with a GIVEN neighbour table idx_global [M, k] (global rows)."""
    M, C = x.shape
    k = idx_global.shape[1]
    x_i = x.unsqueeze(1).expand(M, k, C)
    x_j = x[idx_global]
    msg = mlp(torch.cat([x_i, x_j - x_i], dim=-1).reshape(M * k, 2 * C)).view(M, k, -1)
    return msg.max(dim=1).values


def global_index(idx_local, N):
    M = idx_local.shape[0]
    base = (torch.arange(M, device=idx_local.device) // N * N).unsqueeze(1)
    return idx_local.long() + base


def assert_grad_close(a, b, what='', l2_tol=2e-3, max_tol=3e-2):
    """Gradients of a ReLU network are DISCONTINUOUS in the pre-activations: two correct fp32-class implementations whose
    forward values differ by 1e-6 (TF32x3 tensor-core products vs cuBLAS SGEMM) disagree on the ReLU mask of the handful of
    pre-activations (out of ~1e6 per layer) that sit within 1e-6 of zero, and each such flip moves one row of a gradient
    by its full magnitude.  Gradients are therefore compared in relative L2 (2e-3) with a loose cap on the worst entry.
    (With bit-identical forward values the two paths agree to 1e-6: tools/debug_conv1b.py on the GPU.)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    l2 = float((a - b).norm() / b.norm().clamp_min(1e-30))
    mx = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert l2 <= l2_tol and mx <= max_tol, '{}: relative L2 error {:.2e} (tol {:.0e}), worst entry {:.2e} (tol {:.0e})'.format(
        what, l2, l2_tol, mx, max_tol)
