"""Cycle trace of the persistent LSTM kernels (development tool): per CTA and step, clock64 stamps of the gate-math warps, the
operand loader and the MMA issuer (csrc/lstm.cu, nt_debug_lstm_trace).  Prints means over the CTAs of each layer."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garment_pattern_estimation_b200 import _lib, ops

dev = torch.device('cuda:0')
R, T, L, H, E = 736, 14, 3, 250, 250
g = torch.Generator().manual_seed(1)
lstm = torch.nn.LSTM(E, H, L, batch_first=True).to(dev)
x = torch.randn(R, E, generator=g).to(dev)
h0 = (0.3 * torch.randn(L, R, H, generator=g)).to(dev)
c0 = (0.3 * torch.randn(L, R, H, generator=g)).to(dev)
params = [getattr(lstm, '{}_l{}'.format(n, l)).detach().clone().requires_grad_(True) for l in range(L)
          for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
lib = _lib.load()
lib.nt_debug_lstm_trace.restype = ctypes.c_int
lib.nt_debug_lstm_trace.argtypes = [ctypes.c_void_p]
ctas = L * 3 * 16
for _ in range(2):
    xx = x.clone().requires_grad_(True)
    out = ops.lstm_decoder(xx, h0, c0, T, params)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
ops._LSTM_SKIP_DW = True


def run(which):
    buf = torch.zeros(ctas * T * 16, dtype=torch.int64, device=dev)
    xx = x.clone().requires_grad_(True)
    if which == 'fwd':
        lib.nt_debug_lstm_trace(buf.data_ptr())
        out = ops.lstm_decoder(xx, h0, c0, T, params)
        torch.cuda.synchronize()
        lib.nt_debug_lstm_trace(None)
    else:
        out = ops.lstm_decoder(xx, h0, c0, T, params)
        torch.cuda.synchronize()
        lib.nt_debug_lstm_trace(buf.data_ptr())
        out.backward(torch.ones_like(out))
        torch.cuda.synchronize()
        lib.nt_debug_lstm_trace(None)
    return buf.view(ctas, T, 16).cpu().double()


for which, names in (('fwd', {0: 'E tmem_full', 1: 'E loaded', 2: 'E cell done', 3: 'E published', 4: 'P step start', 5: 'P first issue',
                              6: 'P x-part issued', 7: 'P all issued', 8: 'M acc free', 9: 'M first x stage', 10: 'M first h stage',
                              11: 'M committed'}),
                     ('bwd', {0: 'E dh ready', 1: 'E cell done', 2: 'E dG published', 3: 'E tmem_full', 4: 'E step end',
                              8: 'M acc free', 9: 'M committed'})):
    tr = run(which)
    print('====', which, '(cycles relative to the CTA\'s "E" stamp 0 of the same step; mean over CTAs of the layer; steps 4..9)')
    for l in range(L):
        sub = tr[l * 48:(l + 1) * 48, 4:10]                 # [cta, step, slot]
        base = sub[:, :, 0:1]
        rel = (sub - base).mean(dim=(0, 1))
        per_step = (sub[:, 1:, 0] - sub[:, :-1, 0]).mean()
        print('layer', l, 'cycles per step: %.0f' % float(per_step), {names[k]: int(rel[k]) for k in names})
