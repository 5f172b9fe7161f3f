"""Times nt_gemm_tn on the operand shapes of the C2 step (development tool).  NT_TN_PRECISION=tf32x3 selects the transposing engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garment_pattern_estimation_b200 import ops
dev = torch.device('cuda:0')
shapes = [(327680, 200, 200, 0), (327680, 200, 200, 1), (327680, 200, 150, 1), (65536, 400, 3, 0), (65536, 400, 150, 0), (65536, 200, 150, 0),
          (65536, 150, 200, 0), (4096, 200, 200, 0), (32, 512, 512, 0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rows, m, n, centred in shapes:
    a = ops._rowbuf(rows, m, dev).normal_()
    b = ops._rowbuf(rows, n, dev).normal_()
    mu = b[:, :n].mean(0) if centred else None
    out = torch.zeros(m, n, dtype=torch.float64 if centred else torch.float32, device=dev)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_tn(a, a.stride(0), m, rows, out, b=b, ldb=b.stride(0), n=n, mu=mu)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    gb = rows * (m + n) * 4 / 1e9
    print('rows %7d m %4d n %4d centred %d ld (%d, %d): %8.1f us  %6.2f TB/s' % (rows, m, n, centred, a.stride(0), b.stride(0), t, gb / t * 1e3))
