"""ORACLE (test infrastructure): ctypes binding of ``knn_oracle.c``.

Restates torch_cluster.knn as reached from nn/net_blocks.py:127-135,174 (DynamicEdgeConv.forward).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnt_oracle.so")
_lib = None


def build(force=False):
    """Compile knn_oracle.c -> libnt_oracle.so (gcc, see Makefile)."""
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnt_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.nt_oracle_knn.restype = ctypes.c_int
        lib.nt_oracle_knn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.nt_oracle_fps.restype = ctypes.c_int
        lib.nt_oracle_fps.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        lib.nt_oracle_radius.restype = ctypes.c_int
        lib.nt_oracle_radius.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_float, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        lib.nt_oracle_sqdist.restype = ctypes.c_float
        lib.nt_oracle_sqdist.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _lib = lib
    return _lib


def knn_indices(x, k, return_dist=False, nthreads=1):
    """x: [B, N, D] fp32 (torch CPU tensor or numpy).  Returns idx [B, N, k] int32, local to each cloud,
    sorted ascending by (squared distance, index); -1 where a cloud has fewer than k points."""
    lib = _load()
    xn = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    xn = np.ascontiguousarray(xn, dtype=np.float32)
    assert xn.ndim == 3
    B, N, D = xn.shape
    idx = np.empty((B, N, k), dtype=np.int32)
    dist = np.empty((B, N, k), dtype=np.float32) if return_dist else None
    rc = lib.nt_oracle_knn(xn.ctypes.data, B, N, D, k, idx.ctypes.data,
                           dist.ctypes.data if return_dist else None, int(nthreads))
    if rc != 0:
        raise RuntimeError("nt_oracle_knn failed with status %d" % rc)
    if return_dist:
        return torch.from_numpy(idx), torch.from_numpy(dist)
    return torch.from_numpy(idx)


def sqdist(a, b):
    lib = _load()
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib.nt_oracle_sqdist(a.ctypes.data, b.ctypes.data, a.shape[0]))


def fps_indices(pos, n_samples):
    """pos: [B, N, D] fp32.  Returns [B, n_samples] int32 local indices (farthest point sampling from point 0, see knn_oracle.c)."""
    lib = _load()
    xn = np.ascontiguousarray(pos.detach().cpu().numpy() if isinstance(pos, torch.Tensor) else pos, dtype=np.float32)
    B, N, D = xn.shape
    idx = np.empty((B, n_samples), dtype=np.int32)
    rc = lib.nt_oracle_fps(xn.ctypes.data, B, N, D, n_samples, idx.ctypes.data)
    if rc != 0:
        raise RuntimeError("nt_oracle_fps failed with status %d" % rc)
    return torch.from_numpy(idx)


def radius_neighbours(pos, centres, r, max_nbr):
    """pos: [B, N, D] fp32, centres: [B, M] int32 local indices.  Returns (nbr [B, M, max_nbr] int32 local, -1 padded; count [B, M])."""
    lib = _load()
    xn = np.ascontiguousarray(pos.detach().cpu().numpy() if isinstance(pos, torch.Tensor) else pos, dtype=np.float32)
    cn = np.ascontiguousarray(centres.detach().cpu().numpy() if isinstance(centres, torch.Tensor) else centres, dtype=np.int32)
    B, N, D = xn.shape
    M = cn.shape[1]
    nbr = np.empty((B, M, max_nbr), dtype=np.int32)
    cnt = np.empty((B, M), dtype=np.int32)
    rc = lib.nt_oracle_radius(xn.ctypes.data, B, N, D, cn.ctypes.data, M, float(r), max_nbr, nbr.ctypes.data, cnt.ctypes.data)
    if rc != 0:
        raise RuntimeError("nt_oracle_radius failed with status %d" % rc)
    return torch.from_numpy(nbr), torch.from_numpy(cnt)
