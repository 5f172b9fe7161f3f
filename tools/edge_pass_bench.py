"""Developer tool: time the three edge passes of an EdgeConv layer at the C2 shape (65 536 points, k = 5) with the product library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from garment_pattern_estimation_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
B, N, k, H, C = 32, 2048, 5, 200, 150
M, R = B * N, B * N * k
lib = _lib.load()
g = torch.Generator().manual_seed(0)
pq = ops._rowbuf(M, 2 * H, dev); pq.copy_(torch.randn(M, 2 * H, generator=g))
idx = torch.randint(0, N, (M, k), generator=g, dtype=torch.int32).to(dev)
a1 = ops._rowbuf(R, H, dev)
stats = torch.zeros(2 * H, dtype=torch.float64, device=dev)
a3 = ops._rowbuf(R, C, dev); a3[:, :C] = torch.relu(torch.randn(R, C, generator=g)).to(dev)
gout = torch.randn(M, C, device=dev)
sel = torch.randint(0, k, (M, C), dtype=torch.uint8).to(dev)
vec = torch.randn(4, C, device=dev)
sums = torch.randn(2 * C, dtype=torch.float64, device=dev)
dz3 = ops._rowbuf(R, C, dev)
csum = torch.zeros(C, dtype=torch.float64, device=dev)
dz1 = ops._rowbuf(R, H, dev); dz1.copy_(torch.randn(R, H, generator=g))
dpq = torch.zeros(M, 2 * H, device=dev)
st = ops._stream
p = ops._p
calls = {
    'nt_edge_activation': lambda: lib.nt_edge_activation(p(pq), pq.stride(0), H, p(idx), k, N, R, H, p(a1), a1.stride(0), p(stats), st()),
    'nt_bn_relu_bwd_last': lambda: lib.nt_bn_relu_bwd_last(p(a3), a3.stride(0), p(gout), C, p(sel), k, p(vec[2]), p(vec[0]), p(vec[1]), p(sums), R, R, C,
                                                           p(dz3), dz3.stride(0), p(csum), st()),
    'nt_edge_scatter': lambda: lib.nt_edge_scatter(p(dz1), dz1.stride(0), p(idx), k, N, M, H, p(dpq), 2 * H, st()),
}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in calls.items():
    for _ in range(3):
        assert fn() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    print('{:22s} median {:.3f} ms  min {:.3f} ms'.format(name, ts[5], ts[0]))
