import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from garment_pattern_estimation_b200 import ops, net_blocks as nb
dev = torch.device('cuda:0')
def run(engine, widths, rows, seed=3):
    ops.GEMM_ENGINE = engine
    torch.manual_seed(seed)
    mlp = nb.MLP(widths).to(dev).train()
    x = torch.randn(rows, widths[0], device=dev, requires_grad=True)
    out = mlp(x)
    g = torch.randn(rows, widths[-1], device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    out.backward(g)
    res = {'out': out.detach(), 'gx': x.grad}
    for n, p in mlp.named_parameters(): res[n] = p.grad
    return res
for widths, rows in [([7, 9, 5], 64), ([16, 40, 24], 128), ([16, 40, 24], 300), ([153, 153, 153, 23], 1000)]:
    a, b = run('tc', widths, rows), run('simt', widths, rows)
    print(widths, rows, ' '.join('%s:%.1e' % (k, float((a[k]-b[k]).abs().max()/b[k].abs().max().clamp_min(1e-20))) for k in a))
# direct gemm check of the BNRELU_BWD epilogue
import ctypes
from garment_pattern_estimation_b200._lib import NT_EPI_BNRELU_BWD
R, K, Nn = 300, 24, 40
dz = torch.randn(R, K, device=dev); w = torch.randn(Nn, K, device=dev) / 5
aux = torch.randn(R, Nn, device=dev).clamp_min(0); k0 = torch.randn(Nn, device=dev); k1 = torch.randn(Nn, device=dev); mu = torch.randn(Nn, device=dev)
outs = {}
for eng in ('tc', 'simt'):
    ops.GEMM_ENGINE = eng
    out = torch.zeros(R, Nn, device=dev); cs = torch.zeros(Nn, dtype=torch.float64, device=dev)
    ops.gemm_nt(R, K, Nn, w, K, NT_EPI_BNRELU_BWD, a=dz, lda=K, out=out, ldo=Nn, aux=aux, ldaux=Nn, k0=k0, k1=k1, mu=mu, colsum=cs)
    outs[eng] = (out, cs)
want = torch.where(aux > 0, dz @ w.t() - k0 - (aux - mu) * k1, torch.zeros_like(aux))
print('bwd epi tc vs want', float((outs['tc'][0]-want).abs().max()), 'simt', float((outs['simt'][0]-want).abs().max()), 'colsum', float((outs['tc'][1]-want.double().sum(0)).abs().max()), float((outs['simt'][1]-want.double().sum(0)).abs().max()))
