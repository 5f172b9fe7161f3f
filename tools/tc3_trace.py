"""Developer tool: where do the cycles of the streaming row GEMM go?  Builds libnt_b200 with -DNT_TC3_TRACE into tools/_build/
(on the build machine: `python tools/tc3_trace.py --build`), then on a GPU runs the edge-sized GEMMs of the C2 step and prints, per
role, the share of the kernel time one representative thread spends in each phase (mean over CTAs).

    python tools/tc3_trace.py --build          # CPU box
    python tools/tc3_trace.py [engine]         # GPU box; engine = nt_gemm_args.engine (0 auto, 3/4/5 gemm_tc3.cu, 6 gemm_tc4.cu)
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, 'tools', '_build', 'libnt_b200_trace.so')

from garment_pattern_estimation_b200 import build as nt_build  # noqa: E402


def build_trace_lib():
    flags = [f for f in nt_build.NVCC_FLAGS if f != '--shared'] + ['-DNT_TC3_TRACE']
    cmd = [nt_build.nvcc_path()] + flags + ['--shared', '-I', nt_build.INCLUDE, '-o', TRACE_LIB] + nt_build.sources()
    print(' '.join(cmd))
    subprocess.check_call(cmd)


if '--build' in sys.argv:
    build_trace_lib()
    sys.exit(0)

import torch  # noqa: E402

ENGINE = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # nt_gemm_args.engine: 0 / 3 / 4 / 5 = gemm_tc3.cu, 6 = gemm_tc4.cu
nt_build.LIB_PATH = TRACE_LIB
from garment_pattern_estimation_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
read_trace = lib.nt_debug_tc4_trace if ENGINE == 6 else lib.nt_debug_tc3_trace
read_trace.restype = ctypes.c_int
read_trace.argtypes = [ctypes.c_void_p]
ops.NT_ENGINE = ENGINE
dev = torch.device('cuda:0')
rows = 32 * 2048 * 5
NAMES = ['prod: wait cp.async', 'prod: wait empty stage', 'prod: convert+fence+arrive', 'prod: issue cp.async',
         'mma : wait accumulator free', 'mma : wait full stage', 'mma : issue + commit',
         'epi0: wait accumulator', 'epi0: drain', 'epi0: tile prologue', 'epi1: wait accumulator', 'epi1: drain',
         'epi1: tile prologue']


def run(name, epi, K, n_out):
    g = torch.Generator().manual_seed(0)
    a = ops._rowbuf(rows, K, dev)
    a[:, :K] = torch.randn(rows, K, generator=g).to(dev)
    w = (torch.randn(n_out, K, generator=g) / K ** 0.5).to(dev)
    out = ops._rowbuf(rows, n_out, dev)
    kw = dict(a=a, lda=a.stride(0), out=out, ldo=out.stride(0))
    if epi in (_lib.NT_EPI_RELU_STATS, _lib.NT_EPI_RELU_MAXMIN):
        kw['bias'] = torch.randn(n_out, device=dev)
        kw['stats'] = torch.zeros(2 * n_out, dtype=torch.float64, device=dev)
    if epi == _lib.NT_EPI_RELU_MAXMIN:
        M = rows // 5
        kw['agg'] = (torch.empty(M, n_out, device=dev), torch.empty(M, n_out, device=dev),
                     torch.empty(M, n_out, dtype=torch.uint8, device=dev), torch.empty(M, n_out, dtype=torch.uint8, device=dev))
        kw['k_agg'] = 5
    if epi == _lib.NT_EPI_BNRELU_BWD:
        aux = ops._rowbuf(rows, n_out, dev)
        aux[:, :n_out] = torch.relu(torch.randn(rows, n_out, generator=g)).to(dev)
        kw.update(aux=aux, ldaux=aux.stride(0), k0=torch.randn(n_out, device=dev) * 0.1, k1=torch.randn(n_out, device=dev) * 0.1,
                  mu=torch.randn(n_out, device=dev), colsum=torch.zeros(n_out, dtype=torch.float64, device=dev))
    for _ in range(3):
        ops.gemm_nt(rows, K, n_out, w, w.stride(0), epi, **kw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.gemm_nt(rows, K, n_out, w, w.stride(0), epi, **kw)
    e.record()
    torch.cuda.synchronize()
    buf = torch.zeros(256, 16, dtype=torch.int64)
    assert read_trace(buf.data_ptr()) == 0
    t = buf[:148].double()
    ms = s.elapsed_time(e)
    print('== {} K={} n_out={}: {:.3f} ms (engine {})'.format(name, K, n_out, ms, ENGINE))
    total = [t[:, 0:4].sum(1).mean(), t[:, 4:7].sum(1).mean(), t[:, 7:10].sum(1).mean(), t[:, 10:13].sum(1).mean()]
    for i, nm in enumerate(NAMES):
        role = 0 if i < 4 else (1 if i < 7 else (2 if i < 10 else 3))
        print('   {:32s} {:10.0f} cycles  {:5.1f} % of the role'.format(nm, float(t[:, i].mean()), 100 * float(t[:, i].mean() / total[role])))


run('relu_stats', _lib.NT_EPI_RELU_STATS, 200, 200)
run('relu_maxmin', _lib.NT_EPI_RELU_MAXMIN, 200, 150)
run('bnrelu_bwd', _lib.NT_EPI_BNRELU_BWD, 150, 200)
