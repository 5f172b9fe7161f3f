"""ORACLE (test infrastructure): import the UNMODIFIED reference modules in this container.

The reference's nn/nets.py, nn/net_blocks.py and nn/metrics/* import third-party packages that are absent here
(torch_geometric, sparsemax, entmax) and a data package that needs libigl + an external pattern library.  This
module registers ``sys.modules`` stand-ins backed by ``oracle.thirdparty`` (the restated operators) and by empty
placeholders (the data package -- never executed on the hot path), puts /root/reference/nn on ``sys.path`` and
returns the reference's own ``nets`` / ``net_blocks`` modules.  It is used ONLY by tests/golden/make_golden.py and
the CPU tests that run when /root/reference exists; nothing on the GPU box imports it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('NT_REFERENCE_ROOT', '/root/reference')
# The GPU box has no /root/reference.  For the one test that must execute the reference's OWN nets.py / trainer.py on a B200
# (INTEGRATION.md swap A), __graft_entry__.build() stages those files -- verbatim, untracked (git-ignored, like
# tests/golden/_ckpt/) -- under tests/golden/_ref_nn/, which travels with the gpurun snapshot.
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', '_ref_nn')
_STAGED_FILES = ('nn/nets.py', 'nn/trainer.py', 'nn/metrics/composed_loss.py', 'nn/metrics/losses.py',
                 'nn/metrics/metrics.py', 'nn/metrics/eval_utils.py', 'models/att/att.yaml')
if not os.path.isfile(os.path.join(REFERENCE_ROOT, 'nn', 'nets.py')) and os.path.isfile(os.path.join(STAGED_ROOT, 'nn', 'nets.py')):
    REFERENCE_ROOT = STAGED_ROOT


def stage_reference_sources():
    """Copy the handful of reference files swap A executes into the untracked staging directory (build container only)."""
    import shutil
    src_root = os.environ.get('NT_REFERENCE_ROOT', '/root/reference')
    if not os.path.isfile(os.path.join(src_root, 'nn', 'nets.py')):
        return False
    for rel in _STAGED_FILES:
        dst = os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
    return True


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'nn', 'nets.py'))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-ins (idempotent)."""
    from . import thirdparty as tp
    if 'torch_geometric' not in sys.modules or not hasattr(sys.modules['torch_geometric'], '_nt_oracle_stub'):
        geo_nn = _module('torch_geometric.nn',
                         DynamicEdgeConv=tp.DynamicEdgeConv, global_mean_pool=tp.global_mean_pool,
                         global_max_pool=tp.global_max_pool, global_add_pool=tp.global_add_pool,
                         fps=tp.fps, radius=tp.radius, knn=tp.knn, PointConv=tp.PointConv,
                         ASAPooling=tp.ASAPooling)
        _module('torch_geometric', nn=geo_nn, _nt_oracle_stub=True)
        _module('sparsemax', Sparsemax=tp.Sparsemax)
        _module('entmax', SparsemaxLoss=tp.SparsemaxLoss)

        class _DatasetPlaceholder:          # metrics/metrics.py:8 only needs the name at import time
            pass

        class InvalidPatternDefError(Exception):   # metrics/eval_utils.py:8
            pass

        class EmptyPanelError(Exception):
            pass
        if 'wandb' not in sys.modules:      # nn/trainer.py:7 -- a local stand-in (no network): records what the Trainer logs
            class _Run:
                step, resumed = 1, False
            wb = _module('wandb', run=_Run(), config=types.SimpleNamespace(trainer={}), logged=[],
                         watch=lambda *a, **k: None, Image=lambda *a, **k: None)
            wb.log = lambda d, step=None: wb.logged.append((step, dict(d)))
        _module('data', Garment3DPatternFullDataset=_DatasetPlaceholder,
                GarmentStitchPairsDataset=_DatasetPlaceholder, DatasetWrapper=_DatasetPlaceholder,
                InvalidPatternDefError=InvalidPatternDefError, EmptyPanelError=EmptyPanelError,
                NNSewingPattern=_DatasetPlaceholder)
    nn_dir = os.path.join(REFERENCE_ROOT, 'nn')
    if nn_dir not in sys.path:
        sys.path.insert(0, nn_dir)


def import_reference():
    """Returns (nets, net_blocks) -- the reference's own modules, unmodified."""
    if not reference_available():
        raise RuntimeError('reference tree not found at {}'.format(REFERENCE_ROOT))
    install()
    net_blocks = importlib.import_module('net_blocks')
    nets = importlib.import_module('nets')
    return nets, net_blocks


def att_configs():
    """(data_config, nn_config, loss_config) of the shipped attention model, read from models/att/att.yaml."""
    import yaml
    with open(os.path.join(REFERENCE_ROOT, 'models', 'att', 'att.yaml')) as f:
        cfg = yaml.safe_load(f)
    data_config = dict(cfg['dataset'])
    data_config['max_pattern_len'] = 23        # panel_classes_condenced.json => 23 classes (nn/data/datasets.py:377-379)
    nn_config = dict(cfg['NN'])
    loss_config = dict(nn_config.pop('loss'))
    nn_config.pop('pre-trained', None)
    return data_config, nn_config, loss_config


def baseline_configs():
    """(data_config, nn_config, loss_config) of the shipped baseline model (models/baseline/lstm_stitch_tags.yaml): global
    mean pool + pattern LSTM + panel LSTM.  The loss section is reduced to the four regression terms (the stitch / free-class
    terms need the stitch ground truth of the dataset, out of the hot path)."""
    import yaml
    with open(os.path.join(REFERENCE_ROOT, 'models', 'baseline', 'lstm_stitch_tags.yaml')) as f:
        cfg = yaml.safe_load(f)
    data_config = dict(cfg['dataset'])
    data_config['max_pattern_len'] = 23
    nn_config = dict(cfg['NN'])
    loss_config = dict(nn_config.pop('loss'))
    loss_config.update(loss_components=['shape', 'loop', 'rotation', 'translation'], quality_components=[])
    nn_config.pop('pre-trained', None)
    return data_config, nn_config, loss_config


def baseline_checkpoint_state():
    """model_state_dict of models/baseline/lstm_stitch_tags.pth without the DataParallel 'module.' prefix."""
    import torch
    ck = torch.load(os.path.join(REFERENCE_ROOT, 'models', 'baseline', 'lstm_stitch_tags.pth'),
                    map_location='cpu', weights_only=False)
    return {k[len('module.'):] if k.startswith('module.') else k: v for k, v in ck['model_state_dict'].items()}


def stitch_checkpoint_state():
    """model_state_dict of models/att/neural_tailor_stitch_model.pth without the DataParallel 'module.' prefix."""
    import torch
    ck = torch.load(os.path.join(REFERENCE_ROOT, 'models', 'att', 'neural_tailor_stitch_model.pth'),
                    map_location='cpu', weights_only=False)
    return {k[len('module.'):] if k.startswith('module.') else k: v for k, v in ck['model_state_dict'].items()}


def att_checkpoint_state():
    """model_state_dict of models/att/neural_tailor_panels.pth without the DataParallel 'module.' prefix."""
    import torch
    ck = torch.load(os.path.join(REFERENCE_ROOT, 'models', 'att', 'neural_tailor_panels.pth'),
                    map_location='cpu', weights_only=False)
    return {k[len('module.'):] if k.startswith('module.') else k: v for k, v in ck['model_state_dict'].items()}
