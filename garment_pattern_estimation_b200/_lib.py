"""ctypes binding of libnt_b200.so (C ABI declared in include/nt_b200.h).

There is NO fallback: if the library is missing and cannot be built, or a call returns non-zero, a RuntimeError is
raised (error convention of the reference: RuntimeError, nn/trainer.py:60).
"""
import ctypes
import os

from . import build as _build

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

NT_PROD_PLAIN, NT_PROD_EDGE = 0, 1
NT_EPI_BIAS, NT_EPI_RELU_STATS, NT_EPI_RELU_MAXMIN, NT_EPI_BNRELU_BWD = 0, 1, 2, 3
NT_PREC_BF16X3, NT_PREC_TF32X3 = 0, 1


class GemmArgs(ctypes.Structure):
    """Mirror of `struct nt_gemm_args`."""
    _fields_ = [
        ('rows', c_int64), ('K', c_int), ('n_out', c_int), ('producer', c_int), ('epilogue', c_int),
        ('a', c_void_p), ('lda', c_int),
        ('pq', c_void_p), ('ldpq', c_int), ('qoff', c_int),
        ('idx', c_void_p), ('k', c_int), ('n_per_cloud', c_int),
        ('w', c_void_p), ('ldw', c_int), ('bias', c_void_p), ('w_split', c_void_p), ('precision', c_int),
        ('out', c_void_p), ('ldo', c_int),
        ('stats', c_void_p),
        ('vmax', c_void_p), ('vmin', c_void_p), ('imax', c_void_p), ('imin', c_void_p),
        ('aux', c_void_p), ('ldaux', c_int), ('aux_edge', c_int),
        ('k0', c_void_p), ('k1', c_void_p), ('mu', c_void_p),
        ('colsum', c_void_p),
        ('scatter_dpq', c_void_p), ('ldscatter', c_int), ('engine', c_int),
    ]


class LstmSizes(ctypes.Structure):
    """Mirror of `struct nt_lstm_sizes_t`."""
    _fields_ = [('weights_bytes', c_int64), ('act_bytes', c_int64), ('fwd_workspace_bytes', c_int64), ('y_bytes', c_int64),
                ('cs_bytes', c_int64), ('gates_bytes', c_int64), ('bwd_workspace_bytes', c_int64), ('y_ld', c_int)]


class PatternLossArgs(ctypes.Structure):
    """Mirror of `struct nt_pattern_loss_args`."""
    _fields_ = [('outlines', c_void_p), ('outl_stride_b', c_int64), ('outl_stride_p', c_int64), ('outl_stride_e', c_int64),
                ('rotations', c_void_p), ('rot_stride_b', c_int64), ('rot_stride_p', c_int64),
                ('translations', c_void_p), ('tr_stride_b', c_int64), ('tr_stride_p', c_int64),
                ('gt_outlines', c_void_p), ('gt_rotations', c_void_p), ('gt_translations', c_void_p), ('num_edges', c_void_p),
                ('B', c_int), ('P', c_int), ('Lp', c_int), ('D', c_int), ('Dr', c_int), ('Dt', c_int),
                ('pad_x', c_float), ('pad_y', c_float), ('loop_weight', c_float),
                ('use_shape', c_int), ('use_loop', c_int), ('use_rotation', c_int), ('use_translation', c_int)]


class EdgeConvArgs(ctypes.Structure):
    """Mirror of `struct nt_edgeconv_args`."""
    _fields_ = [('M', c_int64), ('C', c_int), ('H1', c_int), ('H2', c_int), ('H3', c_int), ('k', c_int), ('n_per_cloud', c_int),
                ('x', c_void_p), ('ldx', c_int), ('idx', c_void_p), ('tail_src', c_void_p), ('tail_ld', c_int), ('tail', c_int),
                ('W', c_void_p * 3), ('b', c_void_p * 3), ('gamma', c_void_p * 3), ('beta', c_void_p * 3),
                ('running_mean', c_void_p * 3), ('running_var', c_void_p * 3), ('num_batches_tracked', c_void_p * 3),
                ('momentum', c_float), ('eps', c_float),
                ('out', c_void_p), ('ldo', c_int), ('saved', c_void_p), ('scratch', c_void_p),
                ('gout', c_void_p), ('ldg', c_int), ('gx', c_void_p), ('ldgx', c_int),
                ('gW', c_void_p * 3), ('gb', c_void_p * 3), ('ggamma', c_void_p * 3), ('gbeta', c_void_p * 3)]


_PP = ctypes.POINTER(c_void_p)      # host array of device pointers

_SIGNATURES = {
    'nt_last_error': (ctypes.c_char_p, []),
    'nt_version': (c_int, []),
    'nt_built_arch': (c_int, []),
    'nt_launch_count': (c_int64, []),
    'nt_sizeof': (c_int, [ctypes.c_char_p]),
    'nt_knn_workspace_bytes': (c_int64, [c_int, c_int, c_int]),
    'nt_knn': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'nt_gemm_nt': (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    'nt_gemm_nt_scatter_supported': (c_int, [ctypes.POINTER(GemmArgs)]),
    'nt_gemm_weights_bytes': (c_int64, [c_int, c_int, c_int]),
    'nt_gemm_prepare_weights': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'nt_gemm_tn_workspace_bytes': (c_int64, []),
    'nt_gemm_tn': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64,
                           c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'nt_gemm_tn_centered': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64,
                                    c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p]),
    'nt_bn_fold': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nt_edge_stats': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_void_p]),
    'nt_edge_activation': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p]),
    'nt_maxmin_finish': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'nt_bn_apply': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
    'nt_bn_bwd_reduce': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p, c_void_p]),
    'nt_bn_relu_bwd_last': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p,
                                    c_void_p]),
    'nt_linear_bn_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nt_edge_scatter': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_int,
                                c_void_p]),
    'nt_sparsemax_fwd': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'nt_sparsemax_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'nt_attn_pool_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                 c_void_p]),
    'nt_attn_pool_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'nt_global_pool_fwd': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'nt_global_pool_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'nt_edgeconv_eval_supported': (c_int, [c_int, c_int, c_int, c_int, c_int]),
    'nt_edgeconv_eval_fwd': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p,
                                     c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    'nt_edgeconv_saved_bytes': (c_int64, [ctypes.POINTER(EdgeConvArgs)]),
    'nt_edgeconv_scratch_bytes': (c_int64, [ctypes.POINTER(EdgeConvArgs), c_int]),
    'nt_edgeconv_train_fwd': (c_int, [ctypes.POINTER(EdgeConvArgs), c_void_p]),
    'nt_edgeconv_train_bwd': (c_int, [ctypes.POINTER(EdgeConvArgs), c_void_p]),
    'nt_edge_pairs': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    'nt_fps': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'nt_radius': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    'nt_point_edges_count': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'nt_point_edges_fill': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'nt_scatter_max_fwd': (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nt_scatter_max_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
    'nt_pattern_loss_fwd': (c_int, [ctypes.POINTER(PatternLossArgs), c_void_p, c_void_p, c_void_p]),
    'nt_pattern_loss_bwd': (c_int, [ctypes.POINTER(PatternLossArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nt_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float, c_float,
                             c_float, c_int, c_void_p, c_void_p]),
    'nt_lstm_sizes': (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(LstmSizes)]),
    'nt_lstm_prepare_weights': (c_int, [_PP, _PP, _PP, _PP, c_int, c_int, c_int, c_void_p, c_void_p]),
    'nt_lstm_fwd': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nt_lstm_bwd': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_void_p, c_int, _PP, _PP, _PP, _PP, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))
_lib = None


def library_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent and nvcc is available).  Raises RuntimeError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not _build.is_current():
        # missing, or STALE (sources / header / flags changed since it was linked: a stale library with a drifted argument
        # struct would read garbage instead of failing) -> rebuild, or refuse to load what does not match the sources
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001 -- turn every build problem into the documented error type
            raise RuntimeError('libnt_b200.so is missing or older than its sources and could not be rebuilt ({}); the B200 hot '
                               'path has no CPU or library fallback'.format(e))
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise RuntimeError('cannot load {}: {}'.format(path, e))
    for name, (restype, argtypes) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise RuntimeError('{} does not export {} (stale build? run garment_pattern_estimation_b200/build.py '
                               '--force)'.format(path, name))
        fn.restype, fn.argtypes = restype, argtypes
    for name, struct in (('nt_gemm_args', GemmArgs), ('nt_pattern_loss_args', PatternLossArgs), ('nt_lstm_sizes_t', LstmSizes),
                         ('nt_edgeconv_args', EdgeConvArgs)):
        lib.nt_sizeof.restype, lib.nt_sizeof.argtypes = c_int, [ctypes.c_char_p]
        if lib.nt_sizeof(name.encode()) != ctypes.sizeof(struct):
            raise RuntimeError('{}: struct {} is {} bytes in the library but {} in the ctypes mirror (layout drift)'.format(
                path, name, lib.nt_sizeof(name.encode()), ctypes.sizeof(struct)))
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().nt_last_error()
        raise RuntimeError('libnt_b200 {}: {}'.format(what, msg.decode() if msg else 'status %d' % rc))


def launch_count():
    return int(load().nt_launch_count())
