import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from garment_pattern_estimation_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
for rows, m, n in [(16, 128, 16), (64, 8, 16), (256, 150, 200)]:
    a = torch.randn(rows, m, device=dev); b = torch.randn(rows, n, device=dev)
    want = a.double().t() @ b.double()
    out = torch.zeros(m, n, device=dev)
    ops.gemm_tn(a, a.stride(0), m, rows, out, b=b, ldb=b.stride(0), n=n)
    torch.cuda.synchronize()
    err = float((out.double() - want).abs().max() / want.abs().max())
    print('variant', os.environ.get('NT_TN_VARIANT', '0'), rows, m, n, 'rel err', err, 'out absmax', float(out.abs().max()), 'nonzero frac', float((out != 0).float().mean()))
    if rows == 16:
        print(out[:4, :6]); print(want[:4, :6])
