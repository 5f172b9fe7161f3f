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
