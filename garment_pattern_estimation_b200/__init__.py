"""B200-native (sm_100a) implementation of the NeuralTailor hot path
(maria-korosteleva/Garment-Pattern-Estimation: nn/net_blocks.py EdgeConv encoder, nn/nets.py attention model,
nn/trainer.py data-parallel step).  See DESIGN.md / INTEGRATION.md at the repository root.

Layout
  csrc/ + lib/libnt_b200.so   hand-written CUDA kernels behind the C ABI of include/nt_b200.h
  _lib.py, build.py           ctypes loader / nvcc build recipe (no CPU fallback: a missing library is an error)
  ops.py                      operator layer (autograd Functions over the C ABI)
  net_blocks.py, nets.py      drop-in mirrors of the reference's plugin surface (same names / state_dict keys)
  losses.py, metrics.py       the reference's pattern loss (4 loss terms, GT order / origin matching, no-grad quality metrics) and
                              the stitch-model loss, vectorised on the device
  parallel.py                 one-process-per-GPU data-parallel wrapper (flat-buffer NCCL allreduce) and GraphedTrainStep
                              (the whole training step as one CUDA graph)
"""
from . import net_blocks, nets  # noqa: F401
from .nets import GarmentFullPattern3D, GarmentSegmentPattern3D, StitchOnEdge3DPairs  # noqa: F401

__all__ = ['net_blocks', 'nets', 'GarmentFullPattern3D', 'GarmentSegmentPattern3D', 'StitchOnEdge3DPairs']
