"""ORACLE (test infrastructure): CPU restatement of the third-party operators the reference borrows.

The arithmetic of the reference hot path lives in packages that are not vendored under /root/reference and
cannot be installed here (no network): torch-geometric / torch-cluster / torch-scatter (unpinned,
requirements.txt; docs/Installation.md:46-48 names torch-1.12.0+cu116 wheels) and sparsemax~=0.1.9
(requirements.txt:2).  Each class below restates the PUBLISHED algorithm of the operator and cites the
reference call site it serves.  Parity against the real binaries is UNPINNED (see oracle/__init__.py).
"""
import torch
import torch.nn as nn

import os

from . import knn as _knn

KNN_THREADS = os.cpu_count() or 1     # host threads for the brute-force search (queries are independent)


# ----------------------------------------------------------------------------------------------------
# torch_geometric.nn.DynamicEdgeConv  (call sites: nn/net_blocks.py:127-135, forward at :174)
# ----------------------------------------------------------------------------------------------------
def _equal_cloud_layout(x, batch):
    """The reference always builds `batch` from a dense [B, N, C] tensor (nn/net_blocks.py:164-167), so the
    clouds are contiguous and of equal size.  Returns (B, N)."""
    M = x.shape[0]
    if batch is None:
        return 1, M
    B = int(batch.max().item()) + 1 if batch.numel() else 0   # PyG derives batch_size the same way (host sync)
    if B == 0:
        return 0, 0
    if M % B != 0:
        raise NotImplementedError("oracle DynamicEdgeConv: ragged batches are not produced by the reference path")
    N = M // B
    expect = torch.arange(B, device=batch.device).repeat_interleave(N)
    if not torch.equal(batch.long(), expect):
        raise NotImplementedError("oracle DynamicEdgeConv: batch vector must be contiguous equal-size clouds")
    return B, N


def knn_graph(x, batch, k):
    """torch_cluster.knn(x, x, k, batch, batch) restated (see knn_oracle.c).  x: [M, C].
    Returns neighbour indices [M, k] int64, GLOBAL row ids, ascending by (squared distance, index); the query
    itself is a candidate.  Rows of clouds with fewer than k points would contain -1 (not produced here)."""
    B, N = _equal_cloud_layout(x, batch)
    if N < k:
        raise NotImplementedError("oracle knn_graph: clouds with fewer than k points are outside the hot path")
    idx_local = _knn.knn_indices(x.detach().float().cpu().reshape(B, N, -1), k, nthreads=KNN_THREADS)   # [B,N,k]
    offs = (torch.arange(B, dtype=torch.int64) * N).view(B, 1, 1)
    return (idx_local.long() + offs).reshape(B * N, k).to(x.device)


class DynamicEdgeConv(nn.Module):
    """out_i = aggr_{j in kNN(i)} nn(cat[x_i, x_j - x_i]);  the graph is rebuilt from the CURRENT features on
    every call (dynamic), the centre point is its own nearest neighbour."""

    def __init__(self, nn, k, aggr='max', **kwargs):
        super().__init__()
        self.nn = nn            # attribute name is part of the state_dict contract: conv_layers.N.nn.L.{0,2}.*
        self.k = k
        self.aggr = aggr

    def forward(self, x, batch=None):
        nbr = knn_graph(x, batch, self.k)                               # [M, k]
        M, C = x.shape
        x_i = x.unsqueeze(1).expand(M, self.k, C)
        x_j = x[nbr]                                                    # [M, k, C]
        msg = self.nn(torch.cat([x_i, x_j - x_i], dim=-1).reshape(M * self.k, 2 * C))
        msg = msg.view(M, self.k, -1)
        if self.aggr == 'max':
            return msg.max(dim=1).values
        if self.aggr == 'mean':
            return msg.mean(dim=1)
        if self.aggr == 'add':
            return msg.sum(dim=1)
        raise ValueError('unsupported aggregation {}'.format(self.aggr))


# ----------------------------------------------------------------------------------------------------
# torch_geometric.nn.global_{mean,max,add}_pool  (call sites: nn/net_blocks.py:145-150,184; nn/nets.py:272)
# ----------------------------------------------------------------------------------------------------
def _segments(x, batch, size):
    B = int(size) if size is not None else int(batch.max().item()) + 1
    N = x.shape[0] // B
    return x.view(B, N, x.shape[-1])


def global_mean_pool(x, batch, size=None):
    return _segments(x, batch, size).mean(dim=1)


def global_max_pool(x, batch, size=None):
    return _segments(x, batch, size).max(dim=1).values


def global_add_pool(x, batch, size=None):
    return _segments(x, batch, size).sum(dim=1)


# ----------------------------------------------------------------------------------------------------
# sparsemax.Sparsemax  (call site: nn/nets.py:225; applied at :255/:260 on [B*N, 23] with dim=1)
# Published algorithm (Martins & Astudillo 2016; package sparsemax 0.1.9): translate by max, sort descending,
# k* = max{ j : 1 + j z_(j) > cumsum_j }, tau = (sum_{j<=k*} z_(j) - 1) / k*, out = max(0, z - tau).
# Backward: S = {out != 0}; grad_in = 1_S * (g - mean_S g).
# ----------------------------------------------------------------------------------------------------
class _SparsemaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, dim):
        ctx.dim = dim
        z = inp.transpose(dim, -1)
        z = z - z.max(-1, keepdim=True).values
        zs = z.sort(-1, descending=True).values
        rng = torch.arange(1, z.size(-1) + 1, device=z.device, dtype=z.dtype).expand_as(z)
        bound = 1 + rng * zs
        is_gt = bound.gt(zs.cumsum(-1)).to(z.dtype)
        kk = (is_gt * rng).max(-1, keepdim=True).values
        taus = ((is_gt * zs).sum(-1, keepdim=True) - 1) / kk
        out = torch.max(torch.zeros_like(z), z - taus)
        ctx.save_for_backward(out)
        return out.transpose(dim, -1)

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        g = grad_output.transpose(ctx.dim, -1)
        nz = torch.ne(out, 0)
        cnt = nz.sum(-1, keepdim=True)
        mean = (g * nz).sum(-1, keepdim=True) / cnt
        grad_in = nz * (g - mean)
        return grad_in.transpose(ctx.dim, -1), None


class Sparsemax(nn.Module):
    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        return _SparsemaxFunction.apply(x, self.dim)


# ----------------------------------------------------------------------------------------------------
# Names the reference imports but the shipped configs never execute (SURVEY.md section 2 rows 3,4 -- out of scope)
# ----------------------------------------------------------------------------------------------------
def _out_of_scope(name):
    def fn(*a, **kw):
        raise NotImplementedError('{} is outside the hot path (SURVEY.md section 8)'.format(name))
    return fn


knn = _out_of_scope('torch_geometric.nn.knn')


# ----------------------------------------------------------------------------------------------------
# PointNet++ operators (call sites: nn/net_blocks.py:16,19-21 -- fps, radius, PointConv)
# ----------------------------------------------------------------------------------------------------
def fps(pos, batch, ratio=0.5, random_start=False):
    """torch_geometric.nn.fps / torch_cluster.fps restated (oracle/knn_oracle.c::nt_oracle_fps): ceil(ratio * N) farthest
    points per cloud, returned as GLOBAL row indices in selection order, cloud after cloud.  The library's default is a random
    first point; parity needs the deterministic variant (first point of every cloud), so random_start must stay False."""
    if random_start:
        raise NotImplementedError('oracle fps: a random start cannot be pinned; the restatement starts from the first point')
    import math
    B, N = _equal_cloud_layout(pos, batch)
    n = int(math.ceil(ratio * N))
    idx = _knn.fps_indices(pos.detach().float().cpu().reshape(B, N, -1), n)                # [B, n] local
    return (idx.long() + (torch.arange(B) * N).view(B, 1)).reshape(-1).to(pos.device)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """torch_geometric.nn.radius(x, y, r, batch_x, batch_y, max_num_neighbors) restated (nt_oracle_radius): for every row of
    y, the first max_num_neighbors rows of x of the same cloud (ascending index) with squared distance < r^2.  Returns
    [2, E] = (row: index into y, col: index into x), grouped by y row.  y must be x[idx] for this restatement (as at the only
    call site); it is matched back to its source row by equality."""
    B, N = _equal_cloud_layout(x, batch_x)
    My = y.shape[0] // B
    xc, yc = x.detach().float().cpu().reshape(B, N, -1), y.detach().float().cpu().reshape(B, My, -1)
    centres = torch.empty(B, My, dtype=torch.int32)
    for b in range(B):          # recover the (first) source index of every centre
        eq = (yc[b].unsqueeze(1) == xc[b].unsqueeze(0)).all(-1)
        centres[b] = eq.float().argmax(dim=1).int()
    nbr, cnt = _knn.radius_neighbours(xc, centres, float(r), int(max_num_neighbors))     # [B, My, max], [B, My]
    rows, cols = [], []
    for b in range(B):
        for m in range(My):
            c = int(cnt[b, m])
            rows.append(torch.full((c,), b * My + m, dtype=torch.long))
            cols.append(nbr[b, m, :c].long() + b * N)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0).to(x.device)


class PointConv(nn.Module):
    """torch_geometric.nn.PointConv (= PointNetConv) as published: message = local_nn(cat[x_j, pos_j - pos_i]), max
    aggregation, add_self_loops=True.  In the bipartite call of the reference (pos = (points, centres)) the self-loop handling
    is index-based exactly as in the library: edges whose source INDEX equals their target INDEX are removed, then an edge
    (i -> i) is appended for every i < min(#points, #centres) -- i.e. centre i additionally receives point i, whatever cloud it
    belongs to.  Restated as is."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True, **kw):
        super().__init__()
        self.local_nn, self.global_nn, self.add_self_loops = local_nn, global_nn, add_self_loops

    def forward(self, x, pos, edge_index):
        x_src = x[0] if isinstance(x, tuple) else x
        pos_src, pos_dst = pos if isinstance(pos, tuple) else (pos, pos)
        src, dst = edge_index[0], edge_index[1]
        if self.add_self_loops:
            keep = src != dst
            src, dst = src[keep], dst[keep]
            loops = torch.arange(min(pos_src.shape[0], pos_dst.shape[0]), device=src.device)
            src, dst = torch.cat([src, loops]), torch.cat([dst, loops])
        msg = pos_src[src] - pos_dst[dst]
        if x_src is not None:
            msg = torch.cat([x_src[src], msg], dim=1)
        if self.local_nn is not None:
            msg = self.local_nn(msg)
        M = pos_dst.shape[0]
        out = torch.full((M, msg.shape[1]), float('-inf'), dtype=msg.dtype, device=msg.device)
        out = out.scatter_reduce(0, dst.view(-1, 1).expand_as(msg), msg, reduce='amax', include_self=True)
        out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
        if self.global_nn is not None:
            out = self.global_nn(out)
        return out


class ASAPooling(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, *a, **kw):
        raise NotImplementedError('ASAPooling is outside the hot path')


class SparsemaxLoss(nn.Module):
    """entmax.SparsemaxLoss placeholder (nn/metrics/composed_loss.py:4,196): only built when the
    'segmentation' loss component is requested, which the shipped att config does not (att.yaml:124)."""

    def forward(self, *a, **kw):
        raise NotImplementedError('SparsemaxLoss is outside the hot path')
