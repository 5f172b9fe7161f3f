"""The C-ABI shared library: builds for sm_100a, loads without a GPU, exports every symbol include/nt_b200.h declares,
and the Python binding table covers exactly that set.  No compute calls here (CPU box)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'nt_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nt_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for must in ('nt_knn', 'nt_gemm_nt', 'nt_gemm_tn', 'nt_bn_fold', 'nt_maxmin_finish', 'nt_sparsemax_fwd',
                 'nt_sparsemax_bwd', 'nt_attn_pool_fwd', 'nt_attn_pool_bwd', 'nt_last_error'):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    from garment_pattern_estimation_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for sym in _declared_symbols():
        assert hasattr(lib, sym), 'libnt_b200.so does not export ' + sym
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols(), 'ctypes table and header disagree'
    loaded = _lib.load()
    assert loaded.nt_version() >= 1 and loaded.nt_built_arch() == 100
    assert loaded.nt_launch_count() >= 0


def test_library_contains_sm_100a_code_only():
    from garment_pattern_estimation_b200 import build
    path = build.build()
    try:
        out = subprocess.run(['cuobjdump', '-lelf', path], capture_output=True, text=True, timeout=60).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_bad_arguments_return_error_codes_not_crashes():
    """argument validation happens before any CUDA call, so it can be exercised on a CPU box"""
    from garment_pattern_estimation_b200 import _lib
    lib = _lib.load()
    assert lib.nt_knn(None, 1, 10, 3, 3, 5, None, None, None) != 0
    assert b'null' in lib.nt_last_error()
    buf = (ctypes.c_float * 64)()
    ibuf = (ctypes.c_int32 * 64)()
    assert lib.nt_knn(ctypes.cast(buf, ctypes.c_void_p), 1, 4, 3, 3, 64, ctypes.cast(ibuf, ctypes.c_void_p), None, None) != 0
    assert b'k must be' in lib.nt_last_error()
    assert lib.nt_sparsemax_fwd(ctypes.cast(buf, ctypes.c_void_p), 2, 33, ctypes.cast(buf, ctypes.c_void_p), None) != 0
    with pytest.raises(RuntimeError):
        _lib.check(1, 'nt_sparsemax_fwd')


def test_empty_inputs_are_no_ops_and_new_entry_points_validate():
    """Zero clouds / zero rows return 0 before any CUDA call (the reference's blocks accept empty batches); the engine
    selector and the fused-scatter query validate their arguments."""
    from garment_pattern_estimation_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    fp = ctypes.cast(buf, ctypes.c_void_p)
    ibuf = (ctypes.c_int32 * 64)()
    ip = ctypes.cast(ibuf, ctypes.c_void_p)
    assert lib.nt_knn(fp, 0, 16, 3, 3, 5, ip, None, None) == 0                      # no clouds
    assert lib.nt_attn_pool_fwd(fp, fp, 8, 0, 4, 3, 8, 1.0, fp, None) == 0
    assert lib.nt_global_pool_fwd(fp, 8, 0, 4, 8, 0, fp, None, None) == 0
    assert lib.nt_sparsemax_fwd(fp, 0, 23, fp, None) == 0
    g = _lib.GemmArgs()
    g.rows, g.K, g.n_out = 0, 4, 4
    g.w, g.ldw = fp, 4
    assert lib.nt_gemm_nt(ctypes.byref(g), None) == 0                                # no rows: nothing to launch
    assert lib.nt_global_pool_fwd(fp, 8, 1, 4, 8, 7, fp, None, None) != 0           # unknown pooling mode
    assert b'unknown mode' in lib.nt_last_error()
    g.rows = 10
    assert lib.nt_gemm_nt_scatter_supported(ctypes.byref(g)) == 0                    # no scatter target requested


def test_composite_edgeconv_entry_points_size_queries_and_validation():
    """nt_edgeconv_train_fwd / _bwd: the size queries are pure host functions, and argument validation happens before
    any CUDA call; the header's minimum set of section 8b is present by name."""
    from garment_pattern_estimation_b200 import _lib
    lib = _lib.load()
    syms = _declared_symbols()
    for must in ('nt_edgeconv_train_fwd', 'nt_edgeconv_train_bwd', 'nt_edgeconv_saved_bytes', 'nt_edgeconv_scratch_bytes',
                 'nt_edgeconv_eval_fwd', 'nt_edgeconv_eval_supported'):
        assert must in syms
    assert lib.nt_sizeof(b'nt_edgeconv_args') == ctypes.sizeof(_lib.EdgeConvArgs)
    g = _lib.EdgeConvArgs()
    assert lib.nt_edgeconv_saved_bytes(ctypes.byref(g)) == -1                       # C = H = k = 0: bad shape
    g.M, g.C, g.H1, g.H2, g.H3, g.k, g.n_per_cloud = 32 * 2048, 150, 200, 200, 150, 5, 2048
    saved = lib.nt_edgeconv_saved_bytes(ctypes.byref(g))
    fwd, bwd = lib.nt_edgeconv_scratch_bytes(ctypes.byref(g), 0), lib.nt_edgeconv_scratch_bytes(ctypes.byref(g), 1)
    E = g.M * g.k
    # saved = PQ + a1 + a2 + a3 (padded rows) + sel + vsel + small vectors; the backward keeps two edge-sized gradient buffers
    assert saved % 256 == 0 and fwd % 256 == 0 and bwd % 256 == 0
    assert 4 * E * (200 + 200 + 152) + 4 * g.M * 400 <= saved <= 4 * E * (200 + 200 + 152) + 4 * g.M * 400 + 5 * g.M * 150 + (1 << 21)
    assert 4 * E * (200 + 200) <= bwd <= 4 * E * (200 + 200) + 4 * g.M * 400 + lib.nt_gemm_tn_workspace_bytes() + (8 << 20)
    assert fwd < bwd
    g2 = _lib.EdgeConvArgs()
    ctypes.memmove(ctypes.byref(g2), ctypes.byref(g), ctypes.sizeof(g))
    g2.M = 2 * g.M
    assert lib.nt_edgeconv_saved_bytes(ctypes.byref(g2)) > saved
    # nothing to do / missing buffers: status codes, never a crash
    g0 = _lib.EdgeConvArgs()
    ctypes.memmove(ctypes.byref(g0), ctypes.byref(g), ctypes.sizeof(g))
    g0.M = 0
    assert lib.nt_edgeconv_train_fwd(ctypes.byref(g0), None) == 0
    assert lib.nt_edgeconv_train_bwd(ctypes.byref(g0), None) == 0
    assert lib.nt_edgeconv_train_fwd(ctypes.byref(g), None) != 0
    assert b'nt_edgeconv' in lib.nt_last_error()
    g.k = 500
    assert lib.nt_edgeconv_train_bwd(ctypes.byref(g), None) != 0
    assert b'bad shape' in lib.nt_last_error()
