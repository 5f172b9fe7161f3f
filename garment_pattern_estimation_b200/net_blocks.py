"""Drop-in building blocks: same names, constructor arguments, ``.config`` dicts, forward signatures and state_dict keys
as the reference's ``nn/net_blocks.py`` -- with the arithmetic running in libnt_b200.so (sm_100a).

The reference resolves these classes BY NAME (``getattr(blocks, config['feature_extractor'])`` nn/nets.py:100-101,
``getattr(blocks, config['panel_decoder'])`` nn/nets.py:106-115, ``blocks.MLP(...)`` nn/nets.py:224), so installing this
module as ``net_blocks`` swaps the hot path under the unmodified ``nets.py`` / ``trainer.py``.

In scope: ``MLP``, ``EdgeConvFeatures`` (+ ``DynamicEdgeConv``), ``LSTMDecoderModule``, ``PointNetPlusPlus``.
Out of scope and therefore absent: graph-pooling variants, GRU / MLP / double-reverse decoders.
"""
import math

import torch
import torch.nn as nn

from . import ops


# ----------------------------------------------------------------------------------------------------------
# MLP  (reference nn/net_blocks.py:43-47)
# ----------------------------------------------------------------------------------------------------------
class FusedMLP(nn.Sequential):
    """``Sequential[Sequential(Linear, ReLU, BatchNorm1d)]*`` as parameter container (state_dict keys ``L.0.*`` /
    ``L.2.*`` are the reference's), evaluated by the fused GEMM kernels: BN after ReLU, training-mode statistics over
    all rows, the BN affine folded into the next Linear."""

    def layer_pairs(self):
        return [(blk[0], blk[2]) for blk in self]

    def forward(self, x):
        return ops.fused_mlp(x, self.layer_pairs(), self.training, mode='plain')


def MLP(channels, batch_norm=True):
    if not batch_norm:
        raise NotImplementedError('the reference ignores batch_norm=False as well (nn/net_blocks.py:43-47)')
    if len(channels) < 3:
        raise NotImplementedError('MLP needs at least two Linear layers on the B200 path')
    return FusedMLP(*[nn.Sequential(nn.Linear(channels[i - 1], channels[i]), nn.ReLU(), nn.BatchNorm1d(channels[i]))
                      for i in range(1, len(channels))])


# ----------------------------------------------------------------------------------------------------------
# DynamicEdgeConv  (torch_geometric.nn.DynamicEdgeConv as used at nn/net_blocks.py:127-135)
# ----------------------------------------------------------------------------------------------------------
class DynamicEdgeConv(nn.Module):
    """out_i = max_{j in kNN(i)} nn([x_i, x_j - x_i]) with the kNN graph rebuilt from the current features.

    The attribute name ``nn`` is part of the state_dict contract (``conv_layers.N.nn.L.{0,2}.*``).  Clouds are the
    dense equal-size layout the reference always produces (nn/net_blocks.py:164-167): pass (B, N)."""

    def __init__(self, nn, k, aggr='max', **kwargs):
        super().__init__()
        if aggr != 'max':
            raise NotImplementedError("only aggr='max' (every shipped config) runs on the B200 path")
        self.nn = nn
        self.k = k
        self.aggr = aggr
        self.last_index = None      # kNN indices of the last forward (int32 [M, k], local) -- for parity tests

    def forward(self, x, batch=None, cloud_shape=None, tail_src=None, pad_out=False):
        if cloud_shape is None:
            if batch is None:
                cloud_shape = (1, x.shape[0])
            else:                      # PyG signature: derive (B, N) from the batch vector (costs a host sync)
                B = int(batch.max().item()) + 1
                cloud_shape = (B, x.shape[0] // B)
        B, N = cloud_shape
        idx = ops.knn_graph(x.detach(), B, N, self.k)
        self.last_index = idx
        return ops.fused_mlp(x, self.nn.layer_pairs(), self.nn.training, mode='edge', idx=idx, k=self.k,
                             n_per_cloud=N, tail_src=tail_src, pad_out=pad_out)


# ----------------------------------------------------------------------------------------------------------
# global pools (torch_geometric.nn.global_*_pool; nn/net_blocks.py:145-150) on the dense equal-size layout
# ----------------------------------------------------------------------------------------------------------
def _cloud_shape(x, batch, size):
    B = int(size) if size is not None else int(batch.max().item()) + 1
    return B, x.shape[0] // B


def global_mean_pool(x, batch, size=None):
    return ops.global_pool(x, *_cloud_shape(x, batch, size), 'mean')


def global_max_pool(x, batch, size=None):
    return ops.global_pool(x, *_cloud_shape(x, batch, size), 'max')


def global_add_pool(x, batch, size=None):
    return ops.global_pool(x, *_cloud_shape(x, batch, size), 'add')


# ----------------------------------------------------------------------------------------------------------
# EdgeConvFeatures  (reference nn/net_blocks.py:93-191)
# ----------------------------------------------------------------------------------------------------------
class EdgeConvFeatures(nn.Module):
    """Point-cloud encoder: conv_depth x DynamicEdgeConv (+ skip-concat of xyz) (+ global pool + Linear)."""

    def __init__(self, out_size, config={}):
        super().__init__()
        self.config = {
            'conv_depth': 2, 'k_neighbors': 5, 'EConv_hidden': 200, 'EConv_hidden_depth': 2, 'EConv_feature': 112,
            'EConv_aggr': 'max', 'global_pool': 'mean', 'skip_connections': False, 'graph_pooling': False,
            'pool_ratio': 0.1,
        }
        self.config.update(config)
        cfg = self.config
        if cfg['graph_pooling']:
            raise NotImplementedError('graph_pooling (ASAPooling) is outside the B200 hot path (SURVEY.md section 2)')
        feat, hid, depth = cfg['EConv_feature'], cfg['EConv_hidden'], cfg['EConv_hidden_depth']
        self.conv_layers = nn.ModuleList()
        c_in = 3
        for _ in range(cfg['conv_depth']):
            self.conv_layers.append(DynamicEdgeConv(MLP([2 * c_in] + [hid] * depth + [feat]),
                                                    k=cfg['k_neighbors'], aggr=cfg['EConv_aggr']))
            c_in = feat
        pools = {'max': global_max_pool, 'mean': global_mean_pool, 'add': global_add_pool}
        if cfg['global_pool'] not in pools:
            raise ValueError('{} pooling is not supported'.format(cfg['global_pool']))
        self.global_pool = pools[cfg['global_pool']]
        self.lin = nn.Linear(feat + 3 if cfg['skip_connections'] else feat, out_size)
        self._batch_cache = {}

    def _batch_vector(self, B, N, device):
        key = (B, N, str(device))
        if key not in self._batch_cache:
            self._batch_cache = {key: torch.arange(B, device=device).repeat_interleave(N)}
        return self._batch_cache[key]

    def forward(self, positions, global_pool=True):
        B, N = positions.shape[0], positions.shape[1]
        pos_flat = positions.reshape(B * N, positions.shape[-1])
        batch = self._batch_vector(B, N, positions.device)
        out = pos_flat
        last = len(self.conv_layers) - 1
        for i, conv in enumerate(self.conv_layers):
            # the skip connection (cat[out, pos], nn/net_blocks.py:178-180) is written by the last layer's epilogue
            tail = pos_flat if (i == last and self.config['skip_connections']) else None
            out = conv(out, cloud_shape=(B, N), tail_src=tail, pad_out=(i != last))       # inner layers: 16-byte aligned rows
        if global_pool:
            pooled = self.global_pool(out, batch, B)
            return ops.linear(pooled, self.lin.weight, self.lin.bias), out, batch
        return None, out, batch


# ----------------------------------------------------------------------------------------------------------
# PointNet++  (reference nn/net_blocks.py:10-88)
# ----------------------------------------------------------------------------------------------------------
class PointConv(nn.Module):
    """torch_geometric.nn.PointConv as the reference calls it (nn/net_blocks.py:16,22): bipartite (points -> sampled centres),
    message = local_nn(pos_j - pos_i), max aggregation, the library's index-based self-loop handling (include/nt_b200.h).  The
    attribute name ``local_nn`` is part of the state_dict contract (``sa1_module.conv.local_nn.L.{0,2}.*``)."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True, **kwargs):
        super().__init__()
        if global_nn is not None or not add_self_loops:
            raise NotImplementedError('only PointConv(local_nn) with the default self loops (the reference call) runs on the B200 path')
        self.local_nn = local_nn
        self.global_nn = None
        self.last_edges = None        # (src, dst) of the last forward -- for parity tests

    def forward(self, pos_flat, cloud_shape, centres, nbr, cnt):
        B, N = cloud_shape
        src, dst, msg = ops.point_edges(pos_flat, B, N, centres, nbr, cnt)
        self.last_edges = (src, dst)
        out = self.local_nn(msg)                                               # [E, F] fused Linear/ReLU/BN stack on the edge rows
        return ops.scatter_max(out, dst, B * centres.shape[1])


class _SetAbstractionModule(nn.Module):
    """fps -> radius(max 25 neighbours) -> PointConv (nn/net_blocks.py:10-26) on the dense equal-size layout."""

    def __init__(self, ratio, conv_radius, per_point_nn):
        super().__init__()
        self.ratio = ratio
        self.radius = conv_radius
        self.conv = PointConv(per_point_nn)

    def forward(self, features, pos, cloud_shape):
        if features is not None:
            raise NotImplementedError('set abstraction on point FEATURES (a second sa module) is commented out in the reference '
                                      '(nn/net_blocks.py:66-68) and not built here')
        B, N = cloud_shape
        idx = ops.fps(pos, B, N, self.ratio)                                   # [B, M] local
        nbr, cnt = ops.radius(pos, B, N, idx, self.radius, 25)
        features = self.conv(pos, cloud_shape, idx, nbr, cnt)
        M = idx.shape[1]
        gidx = (idx.long() + (torch.arange(B, device=pos.device) * N).view(B, 1)).reshape(-1)
        return features, pos[gidx], (B, M)


class _GlobalSetAbstractionModule(nn.Module):
    """cat[features, pos] -> per-point MLP -> global max pool (nn/net_blocks.py:29-40)."""

    def __init__(self, per_point_net):
        super().__init__()
        self.nn = per_point_net

    def forward(self, features, pos, cloud_shape):
        B, M = cloud_shape
        features = torch.cat([features, pos], dim=1) if features is not None else pos
        features = self.nn(features)
        features = ops.global_pool(features, B, M, 'max')
        return features, pos.new_zeros((B, 3)), (B, 1)


class PointNetPlusPlus(nn.Module):
    """The reference's alternative extractor (nn/net_blocks.py:50-88): one set-abstraction level + a global one + Linear.
    Like the reference it returns the bare encoding tensor [B, out_size]."""

    def __init__(self, out_size, config={}):
        super().__init__()
        self.config = {'r1': 0.3, 'r2': 0.4, 'r3': 5, 'r4': 7}
        self.config.update(config)
        hid, feat = self.config['EConv_hidden'], self.config['EConv_feature']
        self.sa1_module = _SetAbstractionModule(0.2, self.config['r1'], MLP([3, hid, hid, feat]))
        self.sa_last_module = _GlobalSetAbstractionModule(MLP([3 + feat, hid, hid, feat]))
        self.lin = nn.Linear(feat, out_size)

    def forward(self, positions):
        B, N = positions.shape[0], positions.shape[1]
        pos_flat = positions.reshape(B * N, positions.shape[-1])
        sa_out = self.sa1_module(None, pos_flat, (B, N))
        out, _, _ = self.sa_last_module(*sa_out)
        return ops.linear(out, self.lin.weight, self.lin.bias)


# ----------------------------------------------------------------------------------------------------------
# LSTM decoder  (reference nn/net_blocks.py:302-333, 363-402)
# ----------------------------------------------------------------------------------------------------------
def _init_weights(module, init_type=''):
    if not init_type:
        return
    if 'kaiming_normal' not in init_type:
        raise NotImplementedError('{} weight initialization is not implemented'.format(init_type))
    for name, param in module.named_parameters():
        if 'weight' in name and param.dim() > 1:
            nn.init.kaiming_normal_(param)


def initial_state(n_layers, rows, hidden, device, init_type='', generator=None):
    """h0 / c0 of the decoders.  The reference draws them with kaiming_normal_ on the CPU on EVERY forward and copies
    them over (nn/net_blocks.py:302-315); here the same distribution -- N(0, 2 / (rows * hidden)) -- is
    drawn directly on the device.  Zeros when init_type is empty, as in the reference."""
    if not init_type:
        return torch.zeros(n_layers, rows, hidden, device=device)
    if 'kaiming_normal' not in init_type:
        raise NotImplementedError('{} tenzor initialization is not implemented'.format(init_type))
    std = math.sqrt(2.0 / float(rows * hidden))
    return torch.randn(n_layers, rows, hidden, device=device, generator=generator) * std


class LSTMDecoderModule(nn.Module):
    """Encoding -> repeated out_len times -> LSTM -> Linear.  ``self.lstm`` is an ``nn.LSTM`` used as the PARAMETER CONTAINER
    (state_dict keys ``lstm.weight_ih_l*`` ... are the reference's); the recurrence itself runs in the persistent tcgen05
    kernels of csrc/lstm.cu (``ops.lstm_decoder``), not in cuDNN.  ``lstm_state=(h0, c0)`` or the ``state_provider`` attribute
    inject the initial states (needed for parity: the reference's are random)."""

    def __init__(self, encoding_size, hidden_size, out_elem_size, n_layers, dropout=0, custom_init='kaiming_normal',
                 **kwargs):
        super().__init__()
        self.custom_init = custom_init
        self.n_layers = n_layers
        self.encoding_size = encoding_size
        self.hidden_size = hidden_size
        self.out_elem_size = out_elem_size
        self.dropout = dropout
        self.lstm = nn.LSTM(encoding_size, hidden_size, n_layers, dropout=dropout, batch_first=True)
        self.lin = nn.Linear(hidden_size, out_elem_size)
        _init_weights(self.lstm, init_type=custom_init)
        self.state_provider = None      # callable(n_layers, rows, hidden, device) -> (h0, c0)
        if not ops.lstm_supported(encoding_size, hidden_size, n_layers):
            raise NotImplementedError('LSTMDecoderModule on the B200 path supports hidden_size <= {}, encoding_size <= {}, '
                                      'n_layers <= {}'.format(ops.LSTM_MAX_HIDDEN, ops.LSTM_MAX_INPUT, ops.LSTM_MAX_LAYERS))

    def _flat_params(self):
        return [getattr(self.lstm, '{}_l{}'.format(name, l)) for l in range(self.n_layers)
                for name in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]

    def forward(self, batch_enc, out_len, lstm_state=None):
        rows = batch_enc.size(0)
        if self.dropout and self.training:
            raise NotImplementedError('inter-layer LSTM dropout is not part of the B200 path (every shipped config uses dropout 0)')
        if lstm_state is None and self.state_provider is not None:
            lstm_state = self.state_provider(self.n_layers, rows, self.hidden_size, batch_enc.device)
        if lstm_state is None:
            lstm_state = (initial_state(self.n_layers, rows, self.hidden_size, batch_enc.device, self.custom_init),
                          initial_state(self.n_layers, rows, self.hidden_size, batch_enc.device, self.custom_init))
        seq = ops.lstm_decoder(batch_enc, lstm_state[0], lstm_state[1], out_len, self._flat_params())      # [T, R, H] time-major
        out = ops.linear(seq.reshape(out_len * rows, self.hidden_size), self.lin.weight, self.lin.bias)
        return out.view(out_len, rows, -1).transpose(0, 1)                                                  # [R, T, out]
