"""Quick GPU check + timing of the persistent LSTM kernels against cuDNN (development tool; run under `timeout`)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garment_pattern_estimation_b200 import ops

torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda:0')
R, T, L, H, E = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (736, 14, 3, 250, 250))]
g = torch.Generator().manual_seed(1)
lstm = torch.nn.LSTM(E, H, L, batch_first=True).to(dev)
for n, p in lstm.named_parameters():
    if 'weight' in n:
        torch.nn.init.kaiming_normal_(p)
x = torch.randn(R, E, generator=g).to(dev)
h0 = (0.3 * torch.randn(L, R, H, generator=g)).to(dev)
c0 = (0.3 * torch.randn(L, R, H, generator=g)).to(dev)
params = [getattr(lstm, '{}_l{}'.format(n, l)) for l in range(L) for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
mine = [p.detach().clone().requires_grad_(True) for p in params]
x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
print('forward...', flush=True)
got = ops.lstm_decoder(x1, h0, c0, T, mine).transpose(0, 1)
torch.cuda.synchronize()
print('forward done', flush=True)
want, _ = lstm(x2.unsqueeze(1).repeat(1, T, 1), (h0, c0))
err = float((got - want).abs().max() / want.abs().max())
print('forward rel err', err, flush=True)
gout = torch.randn(want.shape, generator=g).to(dev)
got.backward(gout)
torch.cuda.synchronize()
print('backward done', flush=True)
want.backward(gout)
l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print('dx', l2(x1.grad, x2.grad))
for i, (a, b) in enumerate(zip(mine, params)):
    print('grad', i // 4, ('w_ih', 'w_hh', 'b_ih', 'b_hh')[i % 4], l2(a.grad, b.grad))


def bench(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def mine_fb():
    xx = x.clone().requires_grad_(True)
    o = ops.lstm_decoder(xx, h0, c0, T, mine)
    o.backward(gout.transpose(0, 1))


def mine_f():
    with torch.no_grad():
        ops.lstm_decoder(x, h0, c0, T, mine)


def ref_fb():
    xx = x.clone().requires_grad_(True)
    o, _ = lstm(xx.unsqueeze(1).repeat(1, T, 1), (h0, c0))
    o.backward(gout)


def ref_f():
    with torch.no_grad():
        lstm(x.unsqueeze(1).repeat(1, T, 1), (h0, c0))


print('ms fwd+bwd: mine %.3f  cudnn(fp32) %.3f' % (bench(mine_fb), bench(ref_fb)))
print('ms fwd    : mine %.3f  cudnn(fp32) %.3f' % (bench(mine_f), bench(ref_f)))
torch.backends.cudnn.allow_tf32 = True
print('ms fwd+bwd: cudnn(tf32) %.3f   fwd %.3f' % (bench(ref_fb), bench(ref_f)))
ops.EVENT_SINK = {}
for _ in range(5):
    mine_fb()
torch.cuda.synchronize()
print({k: round(sum(s.elapsed_time(e) for s, e in v) / 5, 4) for k, v in ops.EVENT_SINK.items()})
ops.EVENT_SINK = {}
ops._LSTM_SKIP_DW = True
for _ in range(5):
    mine_fb()
torch.cuda.synchronize()
print('without the weight-gradient GEMMs:', {k: round(sum(s.elapsed_time(e) for s, e in v) / 5, 4) for k, v in ops.EVENT_SINK.items()})
