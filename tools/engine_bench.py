"""Developer tool: time the edge-sized row GEMMs of the C2 step (327 680 rows) per engine with the PRODUCT library.

    python tools/engine_bench.py [engine ...]        # default: 0 6
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from garment_pattern_estimation_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
rows = 32 * 2048 * 5
engines = [int(a) for a in sys.argv[1:]] or [0, 6]


def setup(epi, K, n_out):
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
    return w, kw


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, epi, K, n_out in (('relu_stats', _lib.NT_EPI_RELU_STATS, 200, 200), ('relu_maxmin', _lib.NT_EPI_RELU_MAXMIN, 200, 150),
                            ('bnrelu_bwd', _lib.NT_EPI_BNRELU_BWD, 150, 200), ('bnrelu_bwd', _lib.NT_EPI_BNRELU_BWD, 200, 200)):
    w, kw = setup(epi, K, n_out)
    ref = None
    for eng in engines:
        ops.NT_ENGINE = eng
        for _ in range(3):
            ops.gemm_nt(rows, K, n_out, w, w.stride(0), epi, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.gemm_nt(rows, K, n_out, w, w.stride(0), epi, **kw)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        nbytes = 4 * rows * (K + n_out + (n_out if epi == _lib.NT_EPI_BNRELU_BWD else 0))
        o = kw['out'][:, :n_out].clone()
        same = '' if ref is None else (' bit-identical to engine {}'.format(engines[0]) if torch.equal(o, ref) else ' DIFFERS from engine {}'.format(engines[0]))
        if ref is None:
            ref = o
        print('{:12s} K={} n_out={} engine {}: median {:.3f} ms  min {:.3f} ms  {:.2f} TB/s algorithmic{}'.format(
            name, K, n_out, eng, ts[len(ts) // 2], ts[0], nbytes / ts[len(ts) // 2] / 1e9, same))
    ops.NT_ENGINE = 0
